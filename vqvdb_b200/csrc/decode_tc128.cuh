// Tensor-core decoder for the reference's 128-channel decoder (DecoderVec3, python/VQVAE_v2.py:302-325; BASELINE
// config 4): weight stream, parameter block and launcher of decode_tc128.cu.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "weights.hpp"

namespace vqvdb {

// Every 128 -> 128 convolution of this decoder is computed as TWO passes of 64 output channels, each of exactly the
// shape of the float decoder's stem (decode_tc.cu): K = 128 input channels as two 64-channel halves, the three kw taps
// concatenated along N (N = 192).  A unit is one (kd, kh) tap pair x one input-channel half:
//     [3 kw][64 n][64 k] bf16 = 24 KB, 128-byte rows with 16-byte chunks XOR-swizzled by (n & 7),
// stored in consumption order, so the TMA producer moves one contiguous 24 KB block per unit:
//   stem, res0.conv1, res0.conv2, res1.conv1, res1.conv2 : 5 layers x 2 passes x 9 pairs x 2 halves = 180 units
//   folded tail (see below)                              : 3 passes (output channel x, y, z) x 18     =  54 units
// The tail up_conv(128 -> 256) -> PixelShuffle3D(2) -> final(32 -> 3) is linear, and folds exactly as the float
// model's does (decode_tc.cuh) — once per output channel c:
//     out[c][2p + r] = tanh(final.bias[c] + sum_{eps, p + e(r, eps) in grid} G_c[p + e(r, eps)][r*8 + eps])
//     G_c = conv3x3x3(a; Wg_c) + bg_c,   Wg_c[r*8 + eps][ci][s1] = sum_{s2 -> (r, eps)} sum_oc final.w[c][oc][s2] * up.w[oc*8 + rU(r, s2)][ci][s1]
// i.e. three 128 -> 64 convolutions on the 4^3 grid, each followed by the float decoder's 8-term gather.
constexpr int kDec128ConvLayers = 5;
constexpr int kDec128UnitsPerPass = 18;
constexpr int kDec128Passes = 2 * kDec128ConvLayers + 3;                 // 13 per group of two leaves
constexpr int kDec128Units = kDec128Passes * kDec128UnitsPerPass;        // 234
constexpr uint32_t kDec128UnitBytes = 3 * 8192;

// fp32 parameter block (float offsets): every per-channel vector the epilogues read
namespace par128 {
constexpr int stem_b = 0, stem_gn_w = 128, stem_gn_b = 256;
constexpr int res0 = 384, res_stride = 768;  // per block: gn1_w, gn1_b, c1_b, gn2_w, gn2_b, c2_b (128 each)
constexpr int gn1_w = 0, gn1_b = 128, c1_b = 256, gn2_w = 384, gn2_b = 512, c2_b = 640;
constexpr int fold_b = res0 + 2 * res_stride;  // [3][64]
constexpr int fin_b = fold_b + 192;            // [3] + 1 pad
constexpr int total = fin_b + 4;               // 2116
}  // namespace par128

struct Decoder128Weights {
	const uint8_t* units;           // kDec128Units * kDec128UnitBytes
	const __nv_bfloat16* emb_bf16;  // quantizer.embedding [256][128] as bf16
	const float* par;               // par128::total floats
	const float* fc0;               // decoder.attn.fc.0.weight [32][128]
	const float* fc2_t;             // decoder.attn.fc.2.weight [128][32] transposed to [32][128]
};

// True when the pack is the architecture this kernel is written for: D = 128, K = 256, decoder width 128, two residual
// blocks, attention reduction to 32, up_conv to 256, final 32 -> 3.
bool decoder128_supports(const WeightPack& pack);
// Host-side builders (round-to-nearest-even bf16).
std::vector<uint8_t> build_decoder128_units(const WeightPack& pack);
std::vector<float> build_decoder128_params(const WeightPack& pack);
// Folded tail in fp32 (double accumulation): wg [3][64][128][27], bg [3][64].
void build_decoder128_fold(const WeightPack& pack, std::vector<float>& wg, std::vector<float>& bg);

cudaError_t configure_decode_tc128();
// tap_stage >= 0 additionally writes the fp32 activation after stage {0: stem + GroupNorm + ReLU, 1: both residual
// blocks, 2: channel attention} as [leaf][128 ch][64 pos] to tap_out (bring-up aid; -1 in production).
cudaError_t launch_decode_tc128(const Decoder128Weights& w, const uint8_t* dev_indices, int64_t n_leaves, float* dev_voxels,
                                int num_sms, cudaStream_t stream, int tap_stage = -1, float* tap_out = nullptr);

}  // namespace vqvdb
