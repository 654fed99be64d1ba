// Weight-stream geometry, parameter-block offsets and host-side builders of the tensor-core vec3 encoder (no CUDA
// headers: shared by the kernels, encode_tc128.cuh, and the host code that prepares the streams).
#pragma once

#include <cstdint>
#include <vector>

#include "weights.hpp"

namespace vqvdb {

// ---- back kernel: the 4^3 stage --------------------------------------------------------------------------------------
// A GEMM tile is 128 rows = the 64 positions of two leaves.  Every 128 -> 128 convolution is two passes of 64 output
// channels; a pass is 9 (kd, kh) tap pairs x 2 input-channel halves = 18 steps, the three kw taps concatenated along N
// (N = 192).  A step consumes two weight units, [3 kw][64 n][64 k] fp16 = 24 KB each (128-byte rows, 16-byte chunks
// XOR-swizzled by n & 7): first the hi plane, then the lo plane, stored in consumption order.
constexpr int kEnc128BackConvs = 4;                                          // res0.conv1, res0.conv2, res1.conv1, res1.conv2
constexpr int kEnc128StepsPerPass = 18;
constexpr int kEnc128BackPasses = 2 * kEnc128BackConvs;                      // 8 per pair of leaves
constexpr int kEnc128BackUnits = kEnc128BackPasses * kEnc128StepsPerPass * 2;  // 288
constexpr uint32_t kEnc128UnitBytes = 3 * 8192;

// fp32 parameter block of the back kernel (float offsets)
namespace par128e {
constexpr int res0 = 0, res_stride = 768;  // per block: gn1_w, gn1_b, c1_b, gn2_w, gn2_b, c2_b (128 each)
constexpr int gn1_w = 0, gn1_b = 128, c1_b = 256, gn2_w = 384, gn2_b = 512, c2_b = 640;
constexpr int proj_b = res0 + 2 * res_stride;  // [128]
// codebook search on the tensor cores (proj folded into the codebook, see build_encoder128_vq_units):
constexpr int vq_esq2 = proj_b + 128;          // [256] |e_k|^2 - 2 b.e_k
constexpr int vq_const = vq_esq2 + 256;        // max_k |M_k| ; an upper bound of |proj.weight|_2 ; |proj.bias| + max_k |e_k| ; 0
constexpr int total = vq_const + 4;            // 1924
}  // namespace par128e

// proj + codebook distances as ONE contraction (the scheme of the float encoder, encode_tc_stream.hpp): z.e_k =
// (W x + b).e_k = x.(W^T e_k) + b.e_k, so the scores of all 256 codes come from a [128 rows][128 c] x [128 c][256 k] GEMM
// on the attention output with M = E W folded in double precision and split into fp16 hi / lo planes.  Eight units of
// 16 KB follow the conv units of a pair of leaves: for each input-channel half, for each half of the codes: M_hi, M_lo,
// a unit = [128 codes][64 c] fp16 in the conv units' row layout (128-byte rows, 16-byte chunks XOR-swizzled by n & 7).
constexpr int kEnc128VqUnits = 8;
constexpr uint32_t kEnc128VqUnitBytes = 16384;

// ---- front kernel: the 8^3 stage -------------------------------------------------------------------------------------
// A GEMM tile is 128 positions of one leaf (two d slices).  The two 64 -> 64 convolutions are 4 tiles x 9 (kd, kh) steps,
// kw along N (N = 192); a step consumes a hi and a lo unit [3 kw][64 n][64 k] fp16 = 24 KB.  The stride-2 conv
// (64 -> 128) is 27 taps on one tile, a step = hi and lo unit [128 n][64 k] fp16 = 16 KB, each stored in a 24 KB slot.
constexpr int kEnc128FrontUnits = 2 * 9 * 2 + 27 * 2;  // 90 slots of kEnc128UnitBytes

namespace par128f {
constexpr int pre_b = 0, pre_gn_w = 64, pre_gn_b = 128;
constexpr int gn1_w = 192, gn1_b = 256, c1_b = 320, gn2_w = 384, gn2_b = 448, c2_b = 512;
constexpr int down_b = 576;  // [128]
constexpr int total = 704;
}  // namespace par128f

// True when the pack is the architecture these kernels are written for (EncoderVec3 with D = 128, K = 256).
bool encoder128_supports(const WeightPack& pack);
std::vector<uint8_t> build_encoder128_back_units(const WeightPack& pack);
std::vector<float> build_encoder128_back_params(const WeightPack& pack);
std::vector<uint8_t> build_encoder128_front_units(const WeightPack& pack);
std::vector<float> build_encoder128_front_params(const WeightPack& pack);
std::vector<uint8_t> build_encoder128_vq_units(const WeightPack& pack);  // kEnc128VqUnits * kEnc128VqUnitBytes
// M = E W as fp32 [256 k][128 c] (what the units hold, before the split), for tests of the fold
std::vector<float> build_encoder128_vq_fold(const WeightPack& pack);

}  // namespace vqvdb
