// C ABI of the host layer (include/vqvdb_b200_host.h).
#include "vqvdb_b200_host.h"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <memory>
#include <string>

#include "B200Backend.hpp"
#include "VQVAECodec.hpp"
#include "vqvdb_file.hpp"

namespace {
thread_local std::string g_err;
int fail(const std::exception& e) {
	g_err = e.what();
	return -1;
}
}  // namespace

struct vqvdb_host_backend {
	std::unique_ptr<IVQVAECodec> codec;
	Tensor last;
};

namespace {
template <class F>
int timed(double* seconds, F&& f) {
	try {
		const auto t0 = std::chrono::steady_clock::now();
		f();
		if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		return 0;
	} catch (const std::exception& e) {
		return fail(e);
	}
}
}  // namespace

struct vqvdb_host_reader {
	std::unique_ptr<vqvdb::VqvdbReader> reader;
	std::vector<uint8_t> idx;
	std::vector<vqvdb::LeafOrigin> org;
};

extern "C" {

const char* vqvdb_host_last_error(void) { return g_err.c_str(); }

int vqvdb_host_write_file(const char* path, int n_grids, const char* const* names, const float* transforms,
                          const int64_t* latent_shape, uint32_t num_embeddings, const int64_t* counts,
                          const int32_t* const* origins, const uint8_t* const* indices) {
	try {
		vqvdb::VqvdbWriter w(path);
		for (int g = 0; g < n_grids; ++g) {
			vqvdb::GridMetadata m;
			m.name = names[g];
			m.numEmbeddings = num_embeddings;
			m.latentShape = {latent_shape[0], latent_shape[1], latent_shape[2]};
			m.totalBlocks = (size_t)counts[g];
			std::memcpy(m.transform, transforms + 16 * g, 64);
			w.startGrid(m);
			w.writeBatch(indices[g], reinterpret_cast<const vqvdb::LeafOrigin*>(origins[g]), (size_t)counts[g]);
			w.endGrid();
		}
		w.close();
		return 0;
	} catch (const std::exception& e) {
		return fail(e);
	}
}

int vqvdb_host_reader_open(const char* path, vqvdb_host_reader** out) {
	try {
		auto r = std::make_unique<vqvdb_host_reader>();
		r->reader = std::make_unique<vqvdb::VqvdbReader>(path);
		*out = r.release();
		return 0;
	} catch (const std::exception& e) {
		*out = nullptr;
		return fail(e);
	}
}

void vqvdb_host_reader_close(vqvdb_host_reader* r) { delete r; }
int vqvdb_host_reader_num_grids(const vqvdb_host_reader* r) { return r ? (int)r->reader->numGrids() : -1; }
uint32_t vqvdb_host_reader_num_embeddings(const vqvdb_host_reader* r) { return r ? r->reader->numEmbeddings() : 0; }

int vqvdb_host_reader_next_grid(vqvdb_host_reader* r, char* name_buf, int name_cap, float transform[16],
                                int64_t latent_shape[3], int64_t* n_blocks) {
	try {
		const vqvdb::GridMetadata m = r->reader->nextGridMetadata();
		if (name_buf && name_cap > 0) {
			std::strncpy(name_buf, m.name.c_str(), (size_t)name_cap - 1);
			name_buf[name_cap - 1] = 0;
		}
		std::memcpy(transform, m.transform, 64);
		for (int i = 0; i < 3; ++i) latent_shape[i] = i < (int)m.latentShape.size() ? m.latentShape[i] : 1;
		*n_blocks = (int64_t)m.totalBlocks;
		return 0;
	} catch (const std::exception& e) {
		return fail(e);
	}
}

int64_t vqvdb_host_reader_next_batch(vqvdb_host_reader* r, int64_t max_blocks, int32_t* origins, uint8_t* indices) {
	try {
		const size_t n = r->reader->nextBatch((size_t)max_blocks, r->idx, r->org);
		if (n) {
			std::memcpy(origins, r->org.data(), n * sizeof(vqvdb::LeafOrigin));
			std::memcpy(indices, r->idx.data(), r->idx.size());
		}
		return (int64_t)n;
	} catch (const std::exception& e) {
		return fail(e);
	}
}

int vqvdb_host_compress(int cuda_device, const char* out_path, int n_grids, const char* const* names, const float* transforms,
                        const int64_t* counts, const int32_t* const* origins, const float* const* voxels, int64_t batch_size) {
	try {
		CodecConfig cfg;
		cfg.device = CodecConfig::Device::CUDA;
		B200Options opt = B200Backend::defaultOptions();
		opt.cudaDevice = cuda_device;
		VQVAECodec codec(std::make_unique<B200Backend>(cfg, opt));
		std::vector<LeafGrid> grids((size_t)n_grids);
		for (int g = 0; g < n_grids; ++g) {
			grids[g].name = names[g];
			std::memcpy(grids[g].transform, transforms + 16 * g, 64);
			const auto* o = reinterpret_cast<const vqvdb::LeafOrigin*>(origins[g]);
			grids[g].origins.assign(o, o + counts[g]);
			grids[g].voxels.assign(voxels[g], voxels[g] + (size_t)counts[g] * 512);
		}
		codec.compress(grids, out_path, (size_t)batch_size);
		return 0;
	} catch (const std::exception& e) {
		return fail(e);
	}
}

int vqvdb_host_decompress(int cuda_device, const char* in_path, int* n_grids, int64_t* counts, int counts_cap,
                          int32_t* const* origins, float* const* voxels, int64_t batch_size, int fp32_decode) {
	try {
		if (!voxels) {  // size query: container only, no GPU needed
			vqvdb::VqvdbReader rd(in_path);
			int g = 0;
			while (rd.hasNextGrid()) {
				const vqvdb::GridMetadata m = rd.nextGridMetadata();
				if (g < counts_cap) counts[g] = (int64_t)m.totalBlocks;
				++g;
				std::vector<uint8_t> i;
				std::vector<vqvdb::LeafOrigin> o;
				while (rd.hasNext()) rd.nextBatch(1 << 20, i, o);
			}
			*n_grids = g;
			return 0;
		}
		CodecConfig cfg;
		cfg.device = CodecConfig::Device::CUDA;
		B200Options opt = B200Backend::defaultOptions();
		opt.cudaDevice = cuda_device;
		opt.fp32Decode = fp32_decode != 0;
		VQVAECodec codec(std::make_unique<B200Backend>(cfg, opt));
		std::vector<LeafGrid> grids;
		codec.decompress(in_path, grids, (size_t)batch_size);
		*n_grids = (int)grids.size();
		for (size_t g = 0; g < grids.size() && (int)g < counts_cap; ++g) {
			counts[g] = (int64_t)grids[g].leafCount();
			std::memcpy(origins[g], grids[g].origins.data(), grids[g].origins.size() * sizeof(vqvdb::LeafOrigin));
			std::memcpy(voxels[g], grids[g].voxels.data(), grids[g].voxels.size() * sizeof(float));
		}
		return 0;
	} catch (const std::exception& e) {
		return fail(e);
	}
}

int vqvdb_host_backend_create(int cuda_device, vqvdb_host_backend** out) {
	try {
		*out = nullptr;
		CodecConfig cfg;
		cfg.device = CodecConfig::Device::CUDA;
		B200Options opt = B200Backend::defaultOptions();
		opt.cudaDevice = cuda_device;
		auto b = std::make_unique<vqvdb_host_backend>();
		b->codec = std::make_unique<B200Backend>(cfg, opt);
		*out = b.release();
		return 0;
	} catch (const std::exception& e) {
		return fail(e);
	}
}

void vqvdb_host_backend_destroy(vqvdb_host_backend* b) { delete b; }

int vqvdb_host_backend_encode(vqvdb_host_backend* b, const float* leaves, int64_t n, double* seconds) {
	return timed(seconds, [&] {
		const TensorView view{leaves, {n, 1, 8, 8, 8}, DataType::FLOAT32};
		b->last = b->codec->encode(view);
	});
}

int vqvdb_host_backend_decode(vqvdb_host_backend* b, const uint8_t* indices, int64_t n, double* seconds) {
	return timed(seconds, [&] {
		const TensorView view{indices, {n, 4, 4, 4}, DataType::UINT8};
		b->last = b->codec->decode(view);
	});
}

const void* vqvdb_host_backend_result(const vqvdb_host_backend* b, uint64_t* bytes) {
	if (bytes) *bytes = b ? b->last.buffer.size() : 0;
	return b ? b->last.buffer.data() : nullptr;
}

int vqvdb_host_backend_encode_into(vqvdb_host_backend* b, const float* leaves, int64_t n, uint8_t* indices_out, double* seconds) {
	return timed(seconds, [&] { static_cast<const B200Backend&>(*b->codec).encodeInto(leaves, n, indices_out); });
}

int vqvdb_host_backend_decode_into(vqvdb_host_backend* b, const uint8_t* indices, int64_t n, float* voxels_out, double* seconds) {
	return timed(seconds, [&] { static_cast<const B200Backend&>(*b->codec).decodeInto(indices, n, voxels_out); });
}

int vqvdb_host_backend_roundtrip_batched(vqvdb_host_backend* b, const float* leaves, int64_t n, int64_t batch, uint8_t* indices_out,
                                         float* voxels_out, double* seconds) {
	if (!b || batch <= 0 || n < 0) return fail(std::invalid_argument("roundtrip_batched: bad arguments"));
	return timed(seconds, [&] {
		const auto& codec = static_cast<const B200Backend&>(*b->codec);
		for (int64_t first = 0; first < n; first += batch) {  // one synchronous backend call per batch and direction, as the SOPs make them
			const int64_t nb = std::min<int64_t>(batch, n - first);
			codec.encodeInto(leaves + first * 512, nb, indices_out + first * 64);
			codec.decodeInto(indices_out + first * 64, nb, voxels_out + first * 512);
		}
	});
}

int vqvdb_host_orchestrator_accepts(int cuda_device, const char* pack_path) {
	try {
		CodecConfig cfg;
		cfg.device = CodecConfig::Device::CUDA;
		if (pack_path && pack_path[0]) cfg.source = std::filesystem::path(pack_path);
		B200Options opt = B200Backend::defaultOptions();
		opt.cudaDevice = cuda_device;
		VQVAECodec codec(std::make_unique<B200Backend>(cfg, opt));
		return 0;
	} catch (const std::exception& e) {
		return fail(e);
	}
}

}  // extern "C"
