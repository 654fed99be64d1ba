#include "VQVAECodec.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <stdexcept>

#include "B200Backend.hpp"

namespace {
constexpr size_t kLeafVoxels = 512;
}

VQVAECodec::VQVAECodec(std::unique_ptr<IVQVAECodec> backend) : backend_(std::move(backend)) {
	if (!backend_) throw std::runtime_error("VQVAECodec: Backend cannot be null.");
	// The orchestrator is FloatGrid-only, like the reference's (VQVAECodec.hpp:40,49): every buffer below is sized
	// 512 floats per leaf.  A vec3 weight pack behind the backend would make it read / write 3x that.
	if (const auto* b200 = dynamic_cast<const B200Backend*>(backend_.get()); b200 && b200->channels() != 1)
		throw std::runtime_error("VQVAECodec: the loaded model has " + std::to_string(b200->channels()) +
		                         " channels; compress/decompress handle FloatGrid (1-channel) models only.");
}

Tensor VQVAECodec::encodeBatch(const TensorView& cpuBatch) const { return backend_->encode(cpuBatch); }
Tensor VQVAECodec::decodeBatch(const TensorView& cpuBatch) const { return backend_->decode(cpuBatch); }

void VQVAECodec::compress(const std::vector<LeafGrid>& grids, const std::filesystem::path& outPath, size_t batchSize,
                          const InterruptFn& interrupted) const {
	const auto t0 = std::chrono::steady_clock::now();
	vqvdb::VqvdbWriter writer(outPath.string());
	const auto* fast = dynamic_cast<const B200Backend*>(backend_.get());
	size_t total = 0;
	std::vector<uint8_t> indices;
	for (const LeafGrid& grid : grids) {
		const size_t n = grid.leafCount();
		if (n == 0) {  // the reference skips empty grids (VQVAECodec.cpp:89-92)
			std::printf("Grid '%s' has no active voxels. Skipping.\n", grid.name.c_str());
			continue;
		}
		if (grid.voxels.size() != n * kLeafVoxels) throw std::runtime_error("LeafGrid '" + grid.name + "': voxel buffer size mismatch");
		vqvdb::GridMetadata meta;
		meta.name = grid.name;
		meta.numEmbeddings = 256;
		meta.latentShape = backend_->getLatentShape();
		meta.totalBlocks = n;
		std::memcpy(meta.transform, grid.transform, sizeof(meta.transform));
		writer.startGrid(meta);
		const size_t step = batchSize ? batchSize : n;
		for (size_t lo = 0; lo < n; lo += step) {
			if (interrupted && interrupted()) throw std::runtime_error("Interrupted.");
			const size_t cnt = std::min(step, n - lo);
			const float* src = grid.voxels.data() + lo * kLeafVoxels;  // leaves are already contiguous: no per-batch copy
			if (fast) {
				indices.resize(cnt * 64);
				fast->encodeInto(src, (int64_t)cnt, indices.data());
				writer.writeBatch(indices.data(), grid.origins.data() + lo, cnt);
			} else {
				TensorView view{src, {(int64_t)cnt, 1, 8, 8, 8}, DataType::FLOAT32};
				const Tensor enc = encodeBatch(view);
				writer.writeBatch(enc.getData<uint8_t>(), grid.origins.data() + lo, cnt);
			}
		}
		writer.endGrid();
		total += n;
	}
	writer.close();
	const auto ms = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count();
	std::printf("Grid Compression Complete: %zu leaves in %lld ms.\n", total, (long long)ms);
}

void VQVAECodec::decompress(const std::filesystem::path& inPath, std::vector<LeafGrid>& grids, size_t batchSize,
                            const InterruptFn& interrupted) const {
	const auto t0 = std::chrono::steady_clock::now();
	vqvdb::VqvdbReader reader(inPath.string());
	grids.clear();
	const auto* fast = dynamic_cast<const B200Backend*>(backend_.get());
	size_t total = 0;
	std::vector<uint8_t> indices;
	std::vector<vqvdb::LeafOrigin> origins;
	while (reader.hasNextGrid()) {
		const vqvdb::GridMetadata meta = reader.nextGridMetadata();
		if (meta.latentShape != backend_->getLatentShape())
			throw std::runtime_error("File latent shape does not match the loaded model.");
		LeafGrid grid;
		grid.name = meta.name;
		std::memcpy(grid.transform, meta.transform, sizeof(grid.transform));
		grid.origins.reserve(meta.totalBlocks);
		grid.voxels.resize(meta.totalBlocks * kLeafVoxels);
		const size_t step = batchSize ? batchSize : std::max<size_t>(meta.totalBlocks, 1);
		size_t done = 0;
		while (reader.hasNext()) {
			if (interrupted && interrupted()) throw std::runtime_error("Interrupted.");
			const size_t cnt = reader.nextBatch(step, indices, origins);
			if (cnt == 0) break;
			float* dst = grid.voxels.data() + done * kLeafVoxels;  // decoded straight into the grid's storage
			if (fast) {
				fast->decodeInto(indices.data(), (int64_t)cnt, dst);
			} else {
				TensorView view{indices.data(), {(int64_t)cnt, 4, 4, 4}, DataType::UINT8};
				const Tensor dec = decodeBatch(view);
				std::memcpy(dst, dec.getData<float>(), cnt * kLeafVoxels * sizeof(float));
			}
			grid.origins.insert(grid.origins.end(), origins.begin(), origins.end());
			done += cnt;
		}
		total += done;
		grids.push_back(std::move(grid));
	}
	const auto ms = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count();
	std::printf("Multi-Grid Decompression Complete: %zu leaves in %lld ms.\n", total, (long long)ms);
}
