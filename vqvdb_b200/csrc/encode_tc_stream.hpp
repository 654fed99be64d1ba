// Weight stream of the tensor-core (tcgen05 / TMEM) encoder: plain C++ part shared by the host-side builder
// (encode_tc_host.cpp) and the kernel launcher (encode_tc.cuh / encode_tc.cu).
#pragma once

#include <cstdint>
#include <vector>

#include "weights.hpp"

namespace vqvdb {

// The encoder's GEMM-shaped layers consume their weights as a fixed stream of "units" (<= 16 KB) through a
// shared-memory ring.  Every unit is a sequence of B operands in the canonical no-swizzle K-major UMMA layout
// [k-chunk (2)][n][8 elements = 16 B]; conv weights are split into two fp16 planes, w ~= w_hi + w_lo / 2048:
//   res16 conv1 / conv2 : 3 units (one per kd) = 3 (kh) x [2][96][16 B], n = part*48 + kw*16 + cout; streamed once per TILE
//                         GROUP of the conv (kEncTcConvGroups groups of the five 128-row tiles, each with its own
//                         completion signal, so a group's epilogue overlaps the next group's MMAs); the table entries of
//                         the later groups alias the bytes of the first
//   down                : 16 units (2x2x2-tap space-to-depth form; one per (td, th) tap pair and quarter of the 8 input
//                         parity classes) = 2 parity classes x [2][128][16 B], n = part*64 + tw*32 + cout
//   res32 conv1 / conv2 : 9 units each (one per (kd, kh)) = 2 k-steps x [2][192][16 B], n = part*96 + kw*32 + cout
//   proj x codebook     : 4 units (M_hi, M_lo per 16-channel k-step), each [2][256][16 B], n = code
// (part 0 = w_hi, part 1 = w_lo).
//
// proj (1x1 conv, 32 -> 128) feeds nothing but the codebook distances, and z.e_k = (W x + b).e_k = x.(W^T e_k) + b.e_k, so
// the two GEMMs [64 x 32] x [32 x 128] and [64 x 128] x [128 x 256] are ONE [64 x 32] x [32 x 256] with
//     M[k][c] = sum_d e[k][d] W[d][c],      score_k = |e_k|^2 - 2 b.e_k - 2 x.M_k  (= |z - e_k|^2 - |z|^2)
// folded in double precision on the host (build_encoder_vq_fold).  The tensor-core scores only SHORTLIST (rigorous error
// bound in encode_tc.cu); rows whose shortlist has more than one code compute z = W x + b in fp32 and are re-scored
// with the reference's own formula (python/save_for_inference.py:55-61) exactly as before.
// Measured (592 k leaves, B200, profiles/r2_encode_tile_groups.txt): 2 groups 5.36 M leaves/s, 3 groups 5.22 M, 5 groups
// 5.01 M.  Finer groups do shorten the exposed MMA waits (conv2: 4.2 k -> 2.2 k cycles per leaf) but the convs are
// shared-memory-bandwidth bound (operand reads of 134 B/cycle), so epilogue stores that run under more of the MMAs slow
// the MMAs down by more than the overlap buys, and the weights are re-streamed once per group.
#ifndef VQVDB_ENC_TILE_GROUPS
#define VQVDB_ENC_TILE_GROUPS 2
#endif
constexpr int kEncTcConvGroups = VQVDB_ENC_TILE_GROUPS;    // 5: one tile per group; 3: {0,1},{2,3},{4}; 2: {0,1,2},{3,4}
static_assert(kEncTcConvGroups == 2 || kEncTcConvGroups == 3 || kEncTcConvGroups == 5, "tile groups of the 8^3 convs");
// first tile of group g (g == kEncTcConvGroups gives 5)
constexpr int enc_tc_group_first(int g) {
	return kEncTcConvGroups == 5 ? g : kEncTcConvGroups == 3 ? (g * 2 < 5 ? g * 2 : 5) : (g == 0 ? 0 : g == 1 ? 3 : 5);
}
// Ring geometry.  VQVDB_ENC_RING4 = 0: three 16 KB stages (`down` as 8 units of 4 parity classes, the VQ as 2 units of
// M_hi + M_lo); = 1: four 12 KB stages — one more unit in flight for the 9 KB / 12 KB units of the 3x3x3 convs, whose ~300-cycle
// units outrun an L2 round trip with three stages — with `down` cut into 16 units of 2 parity classes and the VQ into 4
// units (M_hi, M_lo per k-step).
// Measured (592 k leaves): three stages 6.520 M leaves/s, four stages 6.585 M (profiles/r2_encode_experiments.txt).
#ifndef VQVDB_ENC_RING4
#define VQVDB_ENC_RING4 1
#endif
constexpr int kEncTcRingStages = VQVDB_ENC_RING4 ? 4 : 3;
constexpr uint32_t kEncTcStageBytes = VQVDB_ENC_RING4 ? 12288 : 16384;
constexpr int kEncTcDownUnits = VQVDB_ENC_RING4 ? 16 : 8;
constexpr int kEncTcVqUnits = VQVDB_ENC_RING4 ? 4 : 2;
constexpr int kEncTcUnits = 2 * 3 * kEncTcConvGroups + kEncTcDownUnits + 18 + kEncTcVqUnits;

struct EncoderTcStream {
	const uint8_t* units;            // device pointer, 16-byte aligned
	uint32_t off[kEncTcUnits];       // byte offset of each unit
	uint32_t bytes[kEncTcUnits];     // multiple of 16, <= kEncTcStageBytes
};

// Builds the unit stream on the host and fills table.off / table.bytes (table.units is set by the caller after upload).
std::vector<uint8_t> build_encoder_tc_units(const WeightPack& pack, EncoderTcStream& table);
// m [256][32] fp32, esq_fold[k] = |e_k|^2 - 2 b.e_k, m_norm[k] >= |M_k| (rounded up; only used in the shortlist bound), m_norm[256] = max_k
void build_encoder_vq_fold(const WeightPack& pack, std::vector<float>& m, std::vector<float>& esq_fold, std::vector<float>& m_norm);

}  // namespace vqvdb
