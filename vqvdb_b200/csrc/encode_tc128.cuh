// Tensor-core encoder for the reference's vec3 model (EncoderVec3, python/VQVAE_v2.py:278-299, + the codebook argmin of
// InferenceVectorQuantizer.get_indices, python/save_for_inference.py:55-61; BASELINE config 4): weight streams,
// parameter blocks and launchers of encode_tc128.cu.
//
// Index parity needs fp32-level accuracy, so every 3x3x3 convolution runs in the split-fp16 scheme of the
// float encoder (encode_tc.cu): a = a_hi + a_lo / 2048, w = w_hi + w_lo / 2048, three fp16 products
//     a_hi * w_hi  ->  accumulator "hh" ;  a_lo * w_hi + a_hi * w_lo  ->  accumulator "mix" (weighted 1 / 2048)
// with fp32 accumulation in TMEM.  The 1x1 projection is folded into the codebook and the distances of all 256 codes come
// from one more split-fp16 GEMM with a rigorous error bound; rows whose two best codes are closer than that bound
// (near-ties) are re-scored on the fp32 pipes with the summation order of oracle/vqvae_oracle.c.
//
// Two kernels, one per spatial resolution, with the stride-2 output (128 channels at 4^3, fp32, 32 KB per leaf) handed
// over in global memory:
//   front : pre.0 (3 -> 64, fp32 FMA) -> GroupNorm -> ReLU -> ResidualBlock(64) -> down1 (64 -> 128, stride 2)
//   back  : 2 x ResidualBlock(128) -> ChannelAttention -> proj (1x1) -> argmin over the 256 codes
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "encode_tc128_stream.hpp"

namespace vqvdb {

struct Encoder128BackWeights {
	const uint8_t* units;   // kEnc128BackUnits * kEnc128UnitBytes
	const float* par;       // par128e::total floats
	const float* fc0;       // encoder.attn.fc.0.weight [32][128]
	const float* fc2_t;     // encoder.attn.fc.2.weight [128][32] transposed to [32][128]
	const uint8_t* vq_units; // kEnc128VqUnits * kEnc128VqUnitBytes: the codebook with proj folded in, fp16 hi / lo planes
	const float* proj_t;    // encoder.proj.weight transposed to [128 c][128 d]   (near-tie rows: z = W x + b in fp32)
	const float* emb;       // quantizer.embedding [256 k][128 d]                 (near-tie rows: exact re-scoring)
	const float* emb_sq;    // [256] sum_d e_kd^2 (fp32, sequential in d as the oracle's)
};

struct Encoder128FrontWeights {
	const uint8_t* units;   // kEnc128FrontUnits * kEnc128UnitBytes
	const float* par;       // par128f::total floats
};
// encoder.pre.0.weight transposed to [3 ic][27 taps][64 c]: passed to the front kernel BY VALUE (a 20 KB kernel parameter),
// so that its FFMAs read the weights from the constant bank instead of spending shared-memory or L2 loads on them
struct alignas(16) Encoder128PreWeights {
	float w[81 * 64];
};

cudaError_t configure_encode_tc128();
cudaError_t configure_encode_tc128_front();
size_t encode_tc128_front_scratch_floats(int num_sms);
// leaves [n][3][512] fp32 -> y [n][128 ch][64 pos] fp32 (the output of down1).  tap_stage >= 0 additionally writes the
// fp32 activation after {0: pre (GroupNorm + ReLU), 1: the residual block} as [leaf][64][512] to tap_out.
cudaError_t launch_encode_tc128_front(const Encoder128FrontWeights& w, const Encoder128PreWeights& pw, const float* dev_leaves, int64_t n_leaves,
                                      float* dev_y, float* dev_scratch, int num_sms, cudaStream_t stream, int tap_stage = -1,
                                      float* tap_out = nullptr);
// y: [n][128 ch][64 pos] fp32, the output of down1; the kernel uses it as the residual stream and overwrites it.  tap_stage >= 0 additionally writes the fp32 activation after
// {0: res_stack.0, 1: res_stack.1, 2: attention, 3: proj (z)} as [leaf][128][64] to tap_out (bring-up aid).
cudaError_t launch_encode_tc128_back(const Encoder128BackWeights& w, float* dev_y, int64_t n_leaves, uint8_t* dev_indices,
                                     int num_sms, cudaStream_t stream, int tap_stage = -1, float* tap_out = nullptr);

}  // namespace vqvdb
