// .vqvdb v3 container, byte-compatible with the reference's VDBStreamWriter / VDBStreamReader
// (/root/reference/src/Utils/VQVDB_Reader.hpp:30-43, VQVDB_Reader.cpp:81-150,168-300).
//
//   file header, 12 B packed: "VQVDB" | u8 version=3 | u8 numGrids | u32 numEmbeddings | u8 latentDimCount
//   per grid: u32 nameLen | name | f32[16] transform | u16[latentDimCount] latentShape | u32 totalBlocks
//             totalBlocks x { i32 x, y, z (leaf origin) | u8[prod(latentShape)] indices }
//
// OpenVDB-free: a leaf origin is three int32 (layout-identical to openvdb::Coord) and the transform is the
// 16 floats of Mat4s::asPointer().  Differences from the reference reader, on purpose:
// the per-grid byte budget is decremented once, so multi-grid files whose first grid exceeds the read buffer
// decode correctly; grids are read with one bulk read and de-interleaved in place.
#pragma once

#include <cstdint>
#include <fstream>
#include <string>
#include <vector>

namespace vqvdb {

struct LeafOrigin {
	int32_t x, y, z;
};
static_assert(sizeof(LeafOrigin) == 12, "must match openvdb::Coord");

struct GridMetadata {
	std::string name;
	uint8_t fileVersion = 3;
	uint32_t numEmbeddings = 0;
	std::vector<int64_t> latentShape;
	size_t totalBlocks = 0;
	float transform[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
	size_t blockBytes() const {
		size_t n = 1;
		for (int64_t d : latentShape) n *= (size_t)d;
		return n;
	}
};

class VqvdbWriter {
   public:
	explicit VqvdbWriter(const std::string& path);  // throws std::runtime_error
	~VqvdbWriter() noexcept;
	VqvdbWriter(const VqvdbWriter&) = delete;
	VqvdbWriter& operator=(const VqvdbWriter&) = delete;

	void startGrid(const GridMetadata& meta);
	// indices: n x blockBytes, origins: n entries; interleaved into {origin | indices} records.
	void writeBatch(const uint8_t* indices, const LeafOrigin* origins, size_t n);
	void endGrid();
	void close();  // rewrites the file header with the final grid count

   private:
	void writeHeader();
	std::ofstream out_;
	std::vector<char> scratch_;
	size_t blockBytes_ = 0;
	size_t declaredBlocks_ = 0, writtenBlocks_ = 0;
	uint8_t numGrids_ = 0;
	bool haveShared_ = false, inGrid_ = false;
	uint32_t sharedNumEmbeddings_ = 0;
	uint8_t sharedLatentDims_ = 0;
};

class VqvdbReader {
   public:
	explicit VqvdbReader(const std::string& path);  // validates magic and version == 3
	bool hasNextGrid() const noexcept { return gridIndex_ < numGrids_; }
	GridMetadata nextGridMetadata();
	bool hasNext() const noexcept { return blocksRead_ < current_.totalBlocks; }
	// Reads up to maxBatch records of the current grid; returns the count and fills the two arrays
	// (indices: count x blockBytes, origins: count).
	size_t nextBatch(size_t maxBatch, std::vector<uint8_t>& indices, std::vector<LeafOrigin>& origins);
	uint32_t numGrids() const { return numGrids_; }
	uint32_t numEmbeddings() const { return sharedNumEmbeddings_; }

   private:
	void readExact(void* dst, size_t n, const char* what);
	std::ifstream in_;
	uint32_t numGrids_ = 0, gridIndex_ = 0, sharedNumEmbeddings_ = 0;
	uint8_t sharedLatentDims_ = 0;
	uint64_t fileBytes_ = 0;
	GridMetadata current_;
	size_t blocksRead_ = 0;
	std::vector<char> scratch_;
};

}  // namespace vqvdb
