// OpenVDB <-> LeafGrid adapter: the only OpenVDB-specific code a host of this library needs.
//
// Header-only and compile-guarded: OpenVDB / TBB are NOT present in the image this repository is built and tested in,
// so nothing here is compiled by `make` or exercised by tests/ — it states, against the OpenVDB public API, the two
// conversions the reference performs inline in its orchestrator:
//   * grid -> leaves  (src/orchestrator/VQVAECodec.cpp:26-65, 84-101): LeafManager over the FloatTree, every leaf's
//     origin and its 512-float buffer (leaf-buffer order x<<6 | y<<3 | z, which is what the kernels consume), the
//     grid name and the affine map as 16 floats;
//   * leaves -> grid  (src/orchestrator/VQVAECodec.cpp:150-200): a linear transform from the 16 floats, one
//     touchLeaf + buffer copy + setValuesOn per decoded leaf.
// Differences from the reference, on purpose: the leaf walk writes straight into ONE contiguous [n][512] buffer (no
// per-batch std::vector, no copy on return), which is what `VQVAECodec::compress` hands to the backend in one call;
// the write-back inserts leaves from a single thread per grid with a ValueAccessor and no per-thread grids + merge —
// decoded leaves never overlap, so leaf insertion order is irrelevant, and tree insertion is no longer next to a
// 64-leaf GPU call but next to a whole-grid one.  With TBB present define VQVDB_B200_ADAPTER_PARALLEL to fill the
// buffers of the inserted leaves in parallel (the tree topology is built first, serially).
//
//   #define VQVDB_B200_WITH_OPENVDB before including, or let __has_include find <openvdb/openvdb.h>.
#pragma once

#if defined(VQVDB_B200_WITH_OPENVDB) || (defined(__has_include) && __has_include(<openvdb/openvdb.h>))

#include <openvdb/openvdb.h>
#include <openvdb/tree/LeafManager.h>

#include <cstring>
#include <stdexcept>
#include <vector>

#ifdef VQVDB_B200_ADAPTER_PARALLEL
#include <tbb/blocked_range.h>
#include <tbb/parallel_for.h>
#endif

#include "VQVAECodec.hpp"

namespace vqvdb_openvdb {

static_assert(openvdb::FloatTree::LeafNodeType::SIZE == 512, "the model is trained on 8^3 leaves");

// Every active leaf of `grid`, in LeafManager order, as one flat [n][512] float buffer + origins.
inline LeafGrid toLeafGrid(const openvdb::FloatGrid& grid) {
	using LeafT = openvdb::FloatTree::LeafNodeType;
	LeafGrid out;
	out.name = grid.getName();
	const openvdb::math::AffineMap::ConstPtr affine = grid.transform().baseMap()->getAffineMap();
	if (!affine) throw std::runtime_error("vqvdb: grid '" + out.name + "' has no affine transform");
	const openvdb::Mat4d m = affine->getMat4();
	for (int i = 0; i < 16; ++i) out.transform[i] = static_cast<float>(m.asPointer()[i]);

	const openvdb::tree::LeafManager<const openvdb::FloatTree> leaves(grid.tree());
	const size_t n = leaves.leafCount();
	out.origins.resize(n);
	out.voxels.resize(n * LeafT::SIZE);
	auto copyLeaf = [&](size_t i) {
		const LeafT& leaf = leaves.leaf(i);
		const openvdb::Coord& o = leaf.origin();
		out.origins[i] = vqvdb::LeafOrigin{o.x(), o.y(), o.z()};
		std::memcpy(out.voxels.data() + i * LeafT::SIZE, leaf.buffer().data(), LeafT::SIZE * sizeof(float));
	};
#ifdef VQVDB_B200_ADAPTER_PARALLEL
	tbb::parallel_for(tbb::blocked_range<size_t>(0, n, 1024), [&](const tbb::blocked_range<size_t>& r) {
		for (size_t i = r.begin(); i != r.end(); ++i) copyLeaf(i);
	});
#else
	for (size_t i = 0; i < n; ++i) copyLeaf(i);
#endif
	return out;
}

// A FloatGrid holding the decoded leaves of `lg`, all 512 voxels of each leaf active (as the reference's write-back).
inline openvdb::FloatGrid::Ptr toFloatGrid(const LeafGrid& lg, float background = 0.0f) {
	using LeafT = openvdb::FloatTree::LeafNodeType;
	if (lg.voxels.size() != lg.origins.size() * LeafT::SIZE) throw std::invalid_argument("vqvdb: LeafGrid voxel / origin count mismatch");
	openvdb::FloatGrid::Ptr grid = openvdb::FloatGrid::create(background);
	openvdb::Mat4R m;
	for (int i = 0; i < 16; ++i) m.asPointer()[i] = static_cast<openvdb::Real>(lg.transform[i]);
	grid->setTransform(openvdb::math::Transform::createLinearTransform(m));
	grid->setName(lg.name);

	const size_t n = lg.leafCount();
	std::vector<LeafT*> nodes(n, nullptr);
	{
		openvdb::FloatGrid::Accessor acc = grid->getAccessor();  // topology first: touchLeaf is not thread-safe
		for (size_t i = 0; i < n; ++i) {
			const vqvdb::LeafOrigin& o = lg.origins[i];
			nodes[i] = acc.touchLeaf(openvdb::Coord(o.x, o.y, o.z));
		}
	}
	auto fillLeaf = [&](size_t i) {
		if (LeafT* leaf = nodes[i]) {
			std::memcpy(leaf->buffer().data(), lg.voxels.data() + i * LeafT::SIZE, LeafT::SIZE * sizeof(float));
			leaf->setValuesOn();
		}
	};
#ifdef VQVDB_B200_ADAPTER_PARALLEL
	tbb::parallel_for(tbb::blocked_range<size_t>(0, n, 1024), [&](const tbb::blocked_range<size_t>& r) {
		for (size_t i = r.begin(); i != r.end(); ++i) fillLeaf(i);
	});
#else
	for (size_t i = 0; i < n; ++i) fillLeaf(i);
#endif
	return grid;
}

// The reference's entry points, on OpenVDB grids: compress a set of float grids to a .vqvdb v3 file / read one back.
// batchSize 0 = the whole grid per backend call (the backend pipelines internally); `interrupted` as in VQVAECodec.hpp.
inline void compress(const VQVAECodec& codec, const std::vector<openvdb::FloatGrid::ConstPtr>& grids, const std::filesystem::path& outPath,
                     size_t batchSize = 0, const VQVAECodec::InterruptFn& interrupted = {}) {
	std::vector<LeafGrid> flat;
	flat.reserve(grids.size());
	for (const auto& g : grids)
		if (g) flat.push_back(toLeafGrid(*g));
	codec.compress(flat, outPath, batchSize, interrupted);
}

inline std::vector<openvdb::FloatGrid::Ptr> decompress(const VQVAECodec& codec, const std::filesystem::path& inPath, float background = 0.0f,
                                                        size_t batchSize = 0, const VQVAECodec::InterruptFn& interrupted = {}) {
	std::vector<LeafGrid> flat;
	codec.decompress(inPath, flat, batchSize, interrupted);
	std::vector<openvdb::FloatGrid::Ptr> grids;
	grids.reserve(flat.size());
	for (const LeafGrid& lg : flat) grids.push_back(toFloatGrid(lg, background));
	return grids;
}

}  // namespace vqvdb_openvdb

#endif  // OpenVDB available
