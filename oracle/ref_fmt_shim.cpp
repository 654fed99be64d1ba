// TEST INFRASTRUCTURE — C entry points around the reference's own .vqvdb v3 container code, VDBStreamWriter and
// VDBStreamReader (/root/reference/src/Utils/VQVDB_Reader.cpp:20-162, 168-335), compiled unmodified from where it
// lies against the stub header oracle/stub/openvdb/Types.h (oracle/Makefile: fmt -> oracle/_ref/libvqvdb_fmt.so).
// tests/test_vqvdb_file.py uses it both ways: files written by vqvdb_b200/cpp/vqvdb_file.cpp must read back through
// the reference reader, and files written by the reference writer must read back through VqvdbReader.
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "Utils/VQVDB_Reader.hpp"

namespace {
thread_local std::string g_err;
}

struct reffmt_reader {
	std::unique_ptr<VDBStreamReader> r;
};

extern "C" {
#define API __attribute__((visibility("default")))

API const char* reffmt_last_error(void) { return g_err.c_str(); }

// One file with n_grids grids; grid g has counts[g] records, handed to VDBStreamWriter::writeBatch `batch` at a time
// (the orchestrator's loop: VQVAECodec.cpp:108-127).
API int reffmt_write(const char* path, int n_grids, const char* const* names, const float* transforms, const int64_t* latent_shape,
                     int latent_rank, uint32_t num_embeddings, const int64_t* counts, const int32_t* const* origins,
                     const uint8_t* const* indices, int64_t batch) {
	try {
		VDBStreamWriter w(path);
		size_t block = 1;
		for (int i = 0; i < latent_rank; ++i) block *= (size_t)latent_shape[i];
		for (int g = 0; g < n_grids; ++g) {
			VQVDBMetadata m;
			m.name = names[g];
			m.numEmbeddings = num_embeddings;
			m.latentShape.assign(latent_shape, latent_shape + latent_rank);
			m.totalBlocks = (size_t)counts[g];
			m.transform = openvdb::math::Mat4s(transforms + 16 * g);
			w.startGrid(m);
			for (int64_t lo = 0; lo < counts[g]; lo += batch) {
				const int64_t n = std::min<int64_t>(batch, counts[g] - lo);
				Tensor t;
				t.dtype = DataType::UINT8;
				t.shape = {n};
				t.shape.insert(t.shape.end(), latent_shape, latent_shape + latent_rank);
				t.buffer.resize((size_t)n * block);
				std::memcpy(t.buffer.data(), indices[g] + (size_t)lo * block, (size_t)n * block);
				std::vector<openvdb::Coord> org((size_t)n);
				std::memcpy(org.data(), origins[g] + (size_t)lo * 3, (size_t)n * 12);
				w.writeBatch(t, org);
			}
			w.endGrid();
		}
		w.close();
		return 0;
	} catch (const std::exception& e) {
		g_err = e.what();
		return -1;
	}
}

API int reffmt_reader_open(const char* path, reffmt_reader** out) {
	try {
		auto r = std::make_unique<reffmt_reader>();
		r->r = std::make_unique<VDBStreamReader>(path);
		*out = r.release();
		return 0;
	} catch (const std::exception& e) {
		g_err = e.what();
		*out = nullptr;
		return -1;
	}
}
API void reffmt_reader_close(reffmt_reader* r) { delete r; }
API int reffmt_reader_has_next_grid(const reffmt_reader* r) { return r->r->hasNextGrid() ? 1 : 0; }
API int reffmt_reader_has_next(const reffmt_reader* r) { return r->r->hasNext() ? 1 : 0; }

API int reffmt_reader_next_grid(reffmt_reader* r, char* name_buf, int name_cap, float transform[16], int64_t latent_shape[8],
                                int* latent_rank, int64_t* n_blocks, uint32_t* num_embeddings) {
	try {
		const VQVDBMetadata m = r->r->nextGridMetadata();
		std::strncpy(name_buf, m.name.c_str(), (size_t)name_cap - 1);
		name_buf[name_cap - 1] = 0;
		std::memcpy(transform, m.transform.asPointer(), 64);
		*latent_rank = (int)m.latentShape.size();
		for (size_t i = 0; i < m.latentShape.size() && i < 8; ++i) latent_shape[i] = m.latentShape[i];
		*n_blocks = (int64_t)m.totalBlocks;
		*num_embeddings = m.numEmbeddings;
		return 0;
	} catch (const std::exception& e) {
		g_err = e.what();
		return -1;
	}
}

// VDBStreamReader::nextBatch(max_blocks): returns the number of records delivered (possibly 0), or -1.
API int64_t reffmt_reader_next_batch(reffmt_reader* r, int64_t max_blocks, int32_t* origins, uint8_t* indices) {
	try {
		const EncodedBatch b = r->r->nextBatch((size_t)max_blocks);
		const size_t n = b.origins.size();
		if (n) {
			std::memcpy(origins, b.origins.data(), n * 12);
			std::memcpy(indices, b.data.buffer.data(), b.data.buffer.size());
		}
		return (int64_t)n;
	} catch (const std::exception& e) {
		g_err = e.what();
		return -1;
	}
}
}
