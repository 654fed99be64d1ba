// Weight tables and launcher of the tensor-core decoder (decode_tc.cu: tcgen05.mma, bf16 operands, fp32 accumulation).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "model.cuh"
#include "weights.hpp"

namespace vqvdb {

// The conv weights are consumed as a fixed stream of 8 KB "units", each a [64 n][64 k] bf16 tile whose
// 16-byte chunks are XOR-swizzled by (n & 7):
//   stem.0    27 taps x 2 input-channel halves      54 units
//   res conv1 27 taps                                27
//   res conv2 27 taps                                27
//   up_conv   4 output passes (64 ch) x 27 taps     108   (the reference factorisation; kept in the stream, not consumed)
//   folded tail 27 taps (see below)                  27
constexpr int kDecUnitsTotal = 216;
constexpr int kDecUnitsWithFold = kDecUnitsTotal + 27;
// The decoder's tail up_conv -> PixelShuffle3D(2) -> final has no nonlinearity in it (python/VQVAE_v2.py:272-275), so it
// is ONE linear map of the attention output a[64][4^3], zero padding at both resolutions included: for output voxel
// V = 2p + r (p in the 4^3 grid, r in {0,1}^3) every tap s2 of `final` reads the shuffled volume at U = V + s2, i.e. the
// up_conv output of cell p + e, e in {-1, 0, +1}^3 (per axis e = 0 or the one neighbour on r's side), and U is inside
// the 8^3 volume exactly when p + e is inside the 4^3 grid.  Grouping the taps by e gives
//     out[2p + r] = sigmoid(final.bias + sum_{eps in {0,1}^3, p + e(r, eps) in grid} G[p + e(r, eps)][r*8 + eps])
//     G = conv3x3x3(a; Wg) + bg,   Wg[r*8 + eps][ci][s1] = sum_{s2 -> (r, eps)} sum_oc final.w[oc][s2] * up.w[oc*8 + rU(r, s2)][ci][s1]
// a 64 -> 64 convolution on the 4^3 grid (a quarter of up_conv's MACs) followed by an 8-term gather per voxel; the
// weights are folded in double precision on the host (build_decoder_fold) and rounded to bf16 once.

struct DecoderMmaWeights {
	const uint8_t* units;             // kDecUnitsTotal * 8192 bytes, consumption order
	const __nv_bfloat16* emb_bf16;    // quantizer.embedding [256][128] as bf16
	const float *stem_b, *stem_gn_w, *stem_gn_b;
	ResWeights res;                   // only the fp32 vectors (gn*, c1_b, c2_b) are used
	const float *fc0, *fc2;
	const float *up_b;
	const float *fin_w, *fin_b;       // decoder.final transposed [32][27]
	const float* fold_b;              // bias of the folded tail conv [64] (r*8 + eps)
};

// Builds the unit stream and the bf16 codebook on the host (round-to-nearest-even).
std::vector<uint8_t> build_decoder_units(const WeightPack& pack);
void build_decoder_fold(const WeightPack& pack, std::vector<float>& wg /*[64][64][27]*/, std::vector<float>& bg /*[64]*/);
std::vector<uint16_t> build_codebook_bf16(const WeightPack& pack);
// The codebook as 8 weight units [64 codes][64 dims] (code group major, then dim half) for the encoder's
// tensor-core VQ shortlist pass: same tile format as the decoder's units.
std::vector<uint8_t> build_codebook_units(const WeightPack& pack);

// kw taps concatenated along N (one A unit per (kd, kh) pair, N = 192), eight worker warps per 128-row tile, the tail run
// as the folded conv + gather described above (45 units per group of 4 leaves).
cudaError_t configure_decode_tc();
// tap_stage >= 0 additionally writes the activation after stage {0: stem+GN+ReLU, 1: residual block,
// 2: attention} as fp32 [leaf][64 ch][64 pos] to tap_out (bring-up aid; -1 in production).
cudaError_t launch_decode_tc(const DecoderMmaWeights& w, const uint8_t* dev_indices, int64_t n_leaves,
                             float* dev_voxels, int num_sms, cudaStream_t stream, int tap_stage = -1,
                             float* tap_out = nullptr);

}  // namespace vqvdb
