#include "vqvdb_file.hpp"

#include <algorithm>
#include <cstring>
#include <iostream>
#include <stdexcept>

namespace vqvdb {

namespace {
constexpr char kMagic[5] = {'V', 'Q', 'V', 'D', 'B'};
constexpr size_t kHeaderBytes = 12;

void packHeader(char (&h)[kHeaderBytes], uint8_t numGrids, uint32_t numEmbeddings, uint8_t latentDims) {
	std::memcpy(h, kMagic, 5);
	h[5] = 3;
	h[6] = (char)numGrids;
	std::memcpy(h + 7, &numEmbeddings, 4);
	h[11] = (char)latentDims;
}
}  // namespace

// ---------------------------------------------------------------- writer

VqvdbWriter::VqvdbWriter(const std::string& path) {
	out_.open(path, std::ios::binary | std::ios::out | std::ios::trunc);
	if (!out_) throw std::runtime_error("Cannot open output file: " + path);
	writeHeader();  // placeholder; rewritten by close()
}

VqvdbWriter::~VqvdbWriter() noexcept {
	try {
		if (out_.is_open()) close();
	} catch (const std::exception& e) {
		std::cerr << "VqvdbWriter: error while closing: " << e.what() << std::endl;
	}
}

void VqvdbWriter::writeHeader() {
	char h[kHeaderBytes];
	packHeader(h, numGrids_, sharedNumEmbeddings_, sharedLatentDims_);
	const auto pos = out_.tellp();
	out_.seekp(0, std::ios::beg);
	out_.write(h, kHeaderBytes);
	if (pos > (std::streampos)kHeaderBytes) out_.seekp(pos);
	if (!out_) throw std::runtime_error("Failed to write file header.");
}

void VqvdbWriter::startGrid(const GridMetadata& m) {
	if (inGrid_) endGrid();
	if (numGrids_ == 255) throw std::runtime_error("a .vqvdb file holds at most 255 grids");
	if (m.totalBlocks > 0xffffffffull) throw std::runtime_error("a .vqvdb grid holds at most 2^32-1 leaves");
	if (!haveShared_) {
		sharedNumEmbeddings_ = m.numEmbeddings;
		sharedLatentDims_ = (uint8_t)m.latentShape.size();
		haveShared_ = true;
	} else {
		if (m.numEmbeddings != sharedNumEmbeddings_) throw std::runtime_error("Inconsistent number of embeddings across grids.");
		if (m.latentShape.size() != sharedLatentDims_) throw std::runtime_error("Inconsistent latent dimension count across grids.");
	}
	blockBytes_ = m.blockBytes();
	declaredBlocks_ = m.totalBlocks;
	writtenBlocks_ = 0;

	std::vector<char> rec;
	const uint32_t nameLen = (uint32_t)m.name.size();
	rec.resize(4 + nameLen + 64 + 2 * m.latentShape.size() + 4);
	char* p = rec.data();
	std::memcpy(p, &nameLen, 4); p += 4;
	std::memcpy(p, m.name.data(), nameLen); p += nameLen;
	std::memcpy(p, m.transform, 64); p += 64;
	for (int64_t d : m.latentShape) {
		const uint16_t v = (uint16_t)d;
		std::memcpy(p, &v, 2); p += 2;
	}
	const uint32_t total = (uint32_t)m.totalBlocks;
	std::memcpy(p, &total, 4);
	out_.write(rec.data(), (std::streamsize)rec.size());
	if (!out_) throw std::runtime_error("Failed to write grid metadata.");
	++numGrids_;
	inGrid_ = true;
	writeHeader();  // like the reference, the header on disk counts grids *started*
}

void VqvdbWriter::writeBatch(const uint8_t* indices, const LeafOrigin* origins, size_t n) {
	if (!inGrid_) throw std::runtime_error("writeBatch outside startGrid/endGrid");
	if (writtenBlocks_ + n > declaredBlocks_) throw std::runtime_error("more blocks written than the grid metadata declared");
	const size_t rec = sizeof(LeafOrigin) + blockBytes_;
	scratch_.resize(n * rec);
	char* p = scratch_.data();
	for (size_t i = 0; i < n; ++i, p += rec) {
		std::memcpy(p, &origins[i], sizeof(LeafOrigin));
		std::memcpy(p + sizeof(LeafOrigin), indices + i * blockBytes_, blockBytes_);
	}
	out_.write(scratch_.data(), (std::streamsize)scratch_.size());
	if (!out_) throw std::runtime_error("Failed to write buffer to file.");
	writtenBlocks_ += n;
}

void VqvdbWriter::endGrid() {
	if (inGrid_ && writtenBlocks_ != declaredBlocks_)
		throw std::runtime_error("grid closed with fewer blocks than its metadata declared");
	inGrid_ = false;
}

void VqvdbWriter::close() {
	if (!out_.is_open()) return;
	inGrid_ = false;
	writeHeader();
	out_.close();
	if (out_.fail()) throw std::runtime_error("Error closing the output file.");
}

// ---------------------------------------------------------------- reader

VqvdbReader::VqvdbReader(const std::string& path) {
	in_.open(path, std::ios::binary | std::ios::in);
	if (!in_) throw std::runtime_error("Cannot open input file: " + path);
	char h[kHeaderBytes];
	readExact(h, kHeaderBytes, "file header");
	if (std::memcmp(h, kMagic, 5) != 0) throw std::runtime_error("Not a VQVDB file (bad magic).");
	if ((uint8_t)h[5] != 3) throw std::runtime_error("Unsupported VQVDB file version " + std::to_string((int)(uint8_t)h[5]));
	numGrids_ = (uint8_t)h[6];
	std::memcpy(&sharedNumEmbeddings_, h + 7, 4);
	sharedLatentDims_ = (uint8_t)h[11];
	in_.seekg(0, std::ios::end);
	fileBytes_ = (uint64_t)in_.tellg();
	in_.seekg((std::streamoff)kHeaderBytes, std::ios::beg);
}

void VqvdbReader::readExact(void* dst, size_t n, const char* what) {
	in_.read(static_cast<char*>(dst), (std::streamsize)n);
	if ((size_t)in_.gcount() != n) throw std::runtime_error(std::string("Unexpected end of file while reading ") + what);
}

GridMetadata VqvdbReader::nextGridMetadata() {
	if (!hasNextGrid()) throw std::runtime_error("No more grids in file.");
	if (hasNext()) {  // skip what the caller left unread of the previous grid
		const size_t rec = sizeof(LeafOrigin) + current_.blockBytes();
		in_.seekg((std::streamoff)((current_.totalBlocks - blocksRead_) * rec), std::ios::cur);
	}
	GridMetadata m;
	uint32_t nameLen = 0;
	readExact(&nameLen, 4, "grid name length");
	if (nameLen > (1u << 20)) throw std::runtime_error("Corrupt grid name length.");
	m.name.resize(nameLen);
	readExact(m.name.data(), nameLen, "grid name");
	readExact(m.transform, 64, "grid transform");
	m.latentShape.resize(sharedLatentDims_);
	for (auto& d : m.latentShape) {
		uint16_t v;
		readExact(&v, 2, "latent shape");
		d = v;
	}
	uint32_t total = 0;
	readExact(&total, 4, "block count");
	m.totalBlocks = total;
	// a corrupt count must not turn into a multi-terabyte resize() in the caller: the records have to be in the file
	const uint64_t here = (uint64_t)in_.tellg();
	if ((uint64_t)total * (sizeof(LeafOrigin) + m.blockBytes()) > fileBytes_ - std::min(here, fileBytes_))
		throw std::runtime_error("Unexpected end of file while reading block records (grid '" + m.name + "' declares " +
		                         std::to_string(total) + " blocks)");
	m.numEmbeddings = sharedNumEmbeddings_;
	m.fileVersion = 3;
	current_ = m;
	blocksRead_ = 0;
	++gridIndex_;
	return m;
}

size_t VqvdbReader::nextBatch(size_t maxBatch, std::vector<uint8_t>& indices, std::vector<LeafOrigin>& origins) {
	const size_t left = current_.totalBlocks - blocksRead_;
	const size_t n = maxBatch < left ? maxBatch : left;
	const size_t bb = current_.blockBytes(), rec = sizeof(LeafOrigin) + bb;
	indices.resize(n * bb);
	origins.resize(n);
	if (n == 0) return 0;
	scratch_.resize(n * rec);
	readExact(scratch_.data(), n * rec, "block records");
	const char* p = scratch_.data();
	for (size_t i = 0; i < n; ++i, p += rec) {
		std::memcpy(&origins[i], p, sizeof(LeafOrigin));
		std::memcpy(indices.data() + i * bb, p + sizeof(LeafOrigin), bb);
	}
	blocksRead_ += n;
	return n;
}

}  // namespace vqvdb
