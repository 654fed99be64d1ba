// Runtime-described model for the architecture-generic kernels (generic_model.cu).
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace vqvdb {

struct GenericRes {
	const float *gn1_w, *gn1_b, *c1_w, *c1_b, *gn2_w, *gn2_b, *c2_w, *c2_b;  // conv weights transposed [cin][27][cout]
};

// Field names follow oracle/vqvae_oracle.c; all conv weights are transposed to [cin][k^3][cout].
struct GenericModel {
	int cin, D, K;
	int e_c0, e_c1, e_gn0, e_down_k, e_nres, e_red;
	const float *e_pre_w, *e_pre_b, *e_gn_w, *e_gn_b;
	GenericRes e_res0;
	const float *e_down_w, *e_down_b;
	GenericRes e_res[2];
	const float *e_fc0, *e_fc2, *e_proj_w, *e_proj_b;
	const float *e_fc2_t, *d_fc2_t;  // attn.fc.2.weight transposed to [hidden][channels] (coalesced reads in the tensor-core kernels)
	const float *emb, *emb_sq;
	int d_c, d_nres, d_red;
	const float *d_stem_w, *d_stem_b, *d_gn_w, *d_gn_b;
	GenericRes d_res[2];
	const float *d_fc0, *d_fc2, *d_up_w, *d_up_b, *d_fin_w, *d_fin_b;
};

size_t generic_scratch_floats(int grid);
cudaError_t launch_encode_generic(const GenericModel& m, const float* leaves, int64_t n, uint8_t* indices, float* scratch, int grid,
                                  cudaStream_t stream);
// The encoder up to and including the stride-2 conv: down_out [n][e_c1][64] fp32.
cudaError_t launch_encode_generic_front(const GenericModel& m, const float* leaves, int64_t n, float* down_out, float* scratch, int grid,
                                        cudaStream_t stream);
cudaError_t launch_decode_generic(const GenericModel& m, const uint8_t* indices, int64_t n, float* voxels, float* scratch, int grid,
                                  cudaStream_t stream);

}  // namespace vqvdb
