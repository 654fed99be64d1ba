// Device-resident weight tables of the float model (C=1, D=128, K=256) and kernel launchers.
// Layer names follow the reference's state_dict.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace vqvdb {

struct ResWeights {
	const float *gn1_w, *gn1_b, *c1_w, *c1_b, *gn2_w, *gn2_b, *c2_w, *c2_b;  // conv weights transposed [cin][27][cout]
};

struct EncoderWeights {
	const float *pre_w, *pre_b;        // encoder.pre.0      [1][27][16]
	const float *pre_gn_w, *pre_gn_b;  // encoder.pre.1      GroupNorm(4,16)
	ResWeights res16;                  // encoder.pre.3      ResidualBlock(16)
	const float *down_w, *down_b;      // encoder.down       [16][64][32], k4 s2 p1
	ResWeights res32;                  // encoder.res_stack.0
	const float *fc0, *fc2;            // encoder.attn.fc.{0,2}  [8][32], [32][8]
	const float *proj_w, *proj_b;      // encoder.proj       transposed [32][128]
	const float *emb_t;                // quantizer.embedding transposed [128][256]
	const float *emb_sq;               // sum(embedding**2, dim=1) [256]
	const float *emb_norm;             // |e_k| [256], rounded up: only used in the shortlist bound
	const float *emb;                  // quantizer.embedding [256][128] fp32 (exact re-scoring)
	const float *fold_esq;             // |e_k|^2 - 2 proj.bias.e_k [256]   (tensor-core encoder: proj folded into the codebook)
	const float *fold_norm;            // |proj.weight^T e_k| [256] + their maximum [1], rounded up: only used in the shortlist bound
};

// The encoder consumes its weights as a fixed stream of "units" (<= 8 KB) through a shared-memory ring:
// pre.0 (1), res16 conv1/conv2 (4+4, 4 input channels each), pre.0 again (1, for the re-derived residual),
// down (16, one input channel each), res32
// conv1/conv2 (16+16, 2 input channels each), proj (2, 16 input channels each) — slices of the transposed
// fp32 tables above — then the codebook as 8 bf16 tiles [64 codes][64 dims] (pre-swizzled like the decoder's
// weight units) for the tensor-core shortlist pass of the VQ, streamed twice (one round per 96 positions).
constexpr int kEncUnits = 76;
struct EncoderUnits {
	const void* ptr[kEncUnits];   // 16-byte aligned
	uint32_t bytes[kEncUnits];    // multiple of 16, <= 8192
};

struct DecoderWeights {
	const float *emb;                  // quantizer.embedding [256][128]
	const float *stem_w, *stem_b;      // decoder.stem.0     [128][27][64]
	const float *stem_gn_w, *stem_gn_b;
	ResWeights res64;                  // decoder.res_stack.0
	const float *fc0, *fc2;            // decoder.attn.fc.{0,2}  [16][64], [64][16]
	const float *up_w, *up_b;          // decoder.up_conv    [64][27][256]
	const float *fin_w, *fin_b;        // decoder.final      [32][27][1]
};

// fp32 CUDA-core kernels (encode_fp32.cu / decode_fp32.cu).  Return the cudaError of the launch.
cudaError_t launch_encode_fp32(const EncoderWeights& w, const EncoderUnits& units, const float* dev_leaves,
                               int64_t n_leaves, uint8_t* dev_indices, int num_sms, cudaStream_t stream);
cudaError_t launch_decode_fp32(const DecoderWeights& w, const uint8_t* dev_indices, int64_t n_leaves,
                               float* dev_voxels, int num_sms, cudaStream_t stream);
cudaError_t configure_encode_fp32();
cudaError_t configure_decode_fp32();

}  // namespace vqvdb
