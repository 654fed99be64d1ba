/* vqvdb_b200_host.h — C ABI of the host layer (libvqvdb_b200_host.so): the .vqvdb v3 container and the
 * OpenVDB-free compress/decompress batch loops, for callers that are not C++ (tests use it through ctypes).
 *
 * Replaces, above the backend boundary:
 *   VDBStreamWriter / VDBStreamReader     /root/reference/src/Utils/VQVDB_Reader.cpp:20-162, 168-335
 *   VQVAECodec::compress / decompress     /root/reference/src/orchestrator/VQVAECodec.cpp:78-134, 137-208
 * A grid is passed as the flat result of the reference's leaf walk: origins int32[n][3] (openvdb::Coord) and
 * voxels float32[n][512] (leaf.buffer().data()).  Every function returns 0 or a negative code; the message of
 * the calling thread's last failure is vqvdb_host_last_error().
 */
#ifndef VQVDB_B200_HOST_H
#define VQVDB_B200_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VQVDB_HOST_API __attribute__((visibility("default")))

typedef struct vqvdb_host_reader vqvdb_host_reader;

/* ---- container only (no GPU) ---- */
/* Writes one file holding n_grids grids of already-encoded indices. names[g] NUL-terminated; transforms
 * float[n_grids][16]; latent_shape int64[3]; counts[g] leaves; origins[g] int32[counts[g]][3]; indices[g]
 * uint8[counts[g]][64]. */
VQVDB_HOST_API int vqvdb_host_write_file(const char* path, int n_grids, const char* const* names, const float* transforms,
                                         const int64_t* latent_shape, uint32_t num_embeddings, const int64_t* counts,
                                         const int32_t* const* origins, const uint8_t* const* indices);
VQVDB_HOST_API int vqvdb_host_reader_open(const char* path, vqvdb_host_reader** out);
VQVDB_HOST_API void vqvdb_host_reader_close(vqvdb_host_reader* r);
VQVDB_HOST_API int vqvdb_host_reader_num_grids(const vqvdb_host_reader* r);
VQVDB_HOST_API uint32_t vqvdb_host_reader_num_embeddings(const vqvdb_host_reader* r);
/* Advances to the next grid and reports its metadata. name_buf may be NULL. */
VQVDB_HOST_API int vqvdb_host_reader_next_grid(vqvdb_host_reader* r, char* name_buf, int name_cap, float transform[16],
                                               int64_t latent_shape[3], int64_t* n_blocks);
/* Reads up to max_blocks records of the current grid; returns the count read (>= 0) or a negative code. */
VQVDB_HOST_API int64_t vqvdb_host_reader_next_batch(vqvdb_host_reader* r, int64_t max_blocks, int32_t* origins, uint8_t* indices);

/* ---- batch loops through the B200 backend (needs a GPU; no CPU fallback) ---- */
VQVDB_HOST_API int vqvdb_host_compress(int cuda_device, const char* out_path, int n_grids, const char* const* names,
                                       const float* transforms, const int64_t* counts, const int32_t* const* origins,
                                       const float* const* voxels, int64_t batch_size);
/* Decodes every grid of in_path.  Two-step: call with voxels == NULL to learn n_grids/counts (counts_cap entries),
 * then again with voxels[g] / origins[g] buffers sized from counts. */
VQVDB_HOST_API int vqvdb_host_decompress(int cuda_device, const char* in_path, int* n_grids, int64_t* counts, int counts_cap,
                                         int32_t* const* origins, float* const* voxels, int64_t batch_size, int fp32_decode);

/* ---- the backend interface itself, for hosts that are not C++ (bench.py's e2e_pageable; needs a GPU) ----
 * A handle owns one IVQVAECodec created as the reference's SOP caches create theirs (SOP_VQVDB_Encoder.cpp:62-72):
 * CodecConfig{Device::CUDA, EmbeddedModel}, BackendType::B200.  encode / decode make the VIRTUAL calls of
 * IVQVAECodec.hpp:121,130 exactly as the orchestrator does (VQVAECodec.cpp:114-120,172-178): a TensorView over the
 * caller's (pageable) memory in, an owning Tensor out.  The Tensor stays inside the handle until the next call;
 * vqvdb_host_backend_result() exposes its bytes.  *seconds receives the wall time of the virtual call alone.
 * The *_into forms are B200Backend::encodeInto / decodeInto: results written straight into caller memory. */
typedef struct vqvdb_host_backend vqvdb_host_backend;
VQVDB_HOST_API int vqvdb_host_backend_create(int cuda_device, vqvdb_host_backend** out);
VQVDB_HOST_API void vqvdb_host_backend_destroy(vqvdb_host_backend* b);
VQVDB_HOST_API int vqvdb_host_backend_encode(vqvdb_host_backend* b, const float* leaves, int64_t n_leaves, double* seconds);
VQVDB_HOST_API int vqvdb_host_backend_decode(vqvdb_host_backend* b, const uint8_t* indices, int64_t n_leaves, double* seconds);
VQVDB_HOST_API const void* vqvdb_host_backend_result(const vqvdb_host_backend* b, uint64_t* bytes);
VQVDB_HOST_API int vqvdb_host_backend_encode_into(vqvdb_host_backend* b, const float* leaves, int64_t n_leaves, uint8_t* indices_out, double* seconds);
VQVDB_HOST_API int vqvdb_host_backend_decode_into(vqvdb_host_backend* b, const uint8_t* indices, int64_t n_leaves, float* voxels_out, double* seconds);
/* The reference SOPs' calling pattern (SOP_VQVDB_Encoder.cpp:36, VQVAECodec.cpp:108-127,166-196: one synchronous backend call
 * per `batch` leaves) as a native loop: for every batch encodeInto, then decodeInto of its indices.  *seconds = wall time of the
 * whole loop (bench.py's e2e_small_batches: no interpreter between the calls, as there is none in a SOP). */
VQVDB_HOST_API int vqvdb_host_backend_roundtrip_batched(vqvdb_host_backend* b, const float* leaves, int64_t n_leaves, int64_t batch,
                                                        uint8_t* indices_out, float* voxels_out, double* seconds);

/* Constructs the orchestrator (VQVAECodec) over a B200 backend whose CodecConfig::source is the weight pack at
 * `pack_path` (NULL or "" = the embedded model) and destroys it again: 0 if it would accept the model, -1 with the
 * message otherwise.  compress / decompress size every buffer as 512 floats per leaf (FloatGrid only, like the
 * reference: VQVAECodec.hpp:40,49), so a 3-channel model is refused at construction. */
VQVDB_HOST_API int vqvdb_host_orchestrator_accepts(int cuda_device, const char* pack_path);

VQVDB_HOST_API const char* vqvdb_host_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
