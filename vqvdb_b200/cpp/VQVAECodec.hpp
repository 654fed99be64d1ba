// OpenVDB-free mirror of the reference's orchestrator (src/orchestrator/VQVAECodec.{hpp,cpp}): the batch
// loops of compress() (VQVAECodec.cpp:78-134) and decompress() (:137-208) over a plain leaf container.
//
// The reference walks an openvdb::FloatGrid with a LeafManager, memcpy's each 2 KB leaf buffer into a fresh
// std::vector per batch (VDBInputBlockStreamer::nextBatch, :36-59) and runs one synchronous backend call per
// batch of 64.  OpenVDB / TBB / HDK are not available here, so a grid is the flat result of that
// walk — LeafGrid = {name, transform, origins[n], voxels[n][512]} — and an OpenVDB adapter only has to fill it
// (leaf.origin(), leaf.buffer().data()) or drain it (touchLeaf + memcpy + setValuesOn, :182-192).
// Batching: the B200 backend re-batches internally (16 K-leaf chunks through a 3-deep H2D/compute/D2H
// pipeline), so `batchSize` only bounds how much is handed over per call; 0 means "the whole grid at once".
#pragma once

#include <filesystem>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "IVQVAECodec.hpp"
#include "vqvdb_file.hpp"

struct LeafGrid {
	std::string name;
	float transform[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};  // Mat4s::asPointer() order
	std::vector<vqvdb::LeafOrigin> origins;  // one per active leaf
	std::vector<float> voxels;               // origins.size() * 512, OpenVDB leaf-buffer order
	size_t leafCount() const { return origins.size(); }
};

class VQVAECodec {
   public:
	explicit VQVAECodec(std::unique_ptr<IVQVAECodec> backend);  // throws on nullptr (VQVAECodec.cpp:71-75)

	// `interrupted` is polled between backend calls (the reference documents a Houdini interrupt handler for progress
	// and cancellation on both entry points — VQVAECodec.hpp:38,47 — but never implemented it); a true return aborts
	// with std::runtime_error("Interrupted.").  With batchSize 0 a grid is one backend call; callers that want to be
	// interruptible inside a grid pass a batch size (kInterruptibleBatch is a good one: ~0.25 s of GPU work).
	using InterruptFn = std::function<bool()>;
	static constexpr size_t kInterruptibleBatch = size_t(1) << 20;
	void compress(const std::vector<LeafGrid>& grids, const std::filesystem::path& outPath, size_t batchSize,
	              const InterruptFn& interrupted = {}) const;
	void decompress(const std::filesystem::path& inPath, std::vector<LeafGrid>& grids, size_t batchSize,
	                const InterruptFn& interrupted = {}) const;

	const IVQVAECodec& backend() const { return *backend_; }

   private:
	[[nodiscard]] Tensor encodeBatch(const TensorView& cpuBatch) const;
	[[nodiscard]] Tensor decodeBatch(const TensorView& cpuBatch) const;
	std::unique_ptr<IVQVAECodec> backend_;
};
