// Tensor-core encoder + codebook argmin on the 5th-generation tensor cores (tcgen05.mma, TMEM accumulators):
// leaf voxels in, 64 uint8 indices out, one kernel.  Same contract as encode_fp32.cu — EncoderFloat.forward
// (python/VQVAE_v2.py:231-250) + InferenceVectorQuantizer.get_indices (python/save_for_inference.py:55-61) + the
// int64->uint8 cast of TorchBackend.cpp:150 — with the six GEMM-shaped layers moved off the CUDA-core FMA pipe.
//
// Index parity is a bit-exactness requirement, and bf16 / tf32 operands flip 0.2-2 % of the indices,
// so every convolution runs at fp32-level accuracy on fp16 tensor-core operands:
//   a = a_hi + a_lo / 2048,  w = w_hi + w_lo / 2048   (a_hi = fp16(a), a_lo = fp16((a - a_hi) * 2048): 22 significant bits)
//   a.w ~= a_hi.w_hi + (a_hi.w_lo + a_lo.w_hi) / 2048                                   (dropped term: 2^-22)
// i.e. three fp16 products with fp32 accumulation in two accumulator groups; measured error vs fp64 is below a plain
// fp32 FMA chain's (tools/microbench/umma_conv16_f16x2.cu: max 1.2e-6 vs 2.7e-6 on the 16->16 conv).
//
// Machine mapping
//   * one leaf per CTA pass.  16 "row" warps: warp = g * 4 + quadrant owns TMEM lanes quadrant * 32 .. + 31 (one GEMM
//     row per lane and 128-row tile) and channel group g of 4 (4 of 16 channels at 8^3, 8 of 32 at 4^3, 32 of the
//     128 latent dims, 64 of the 256 codes).  One warp issues every tcgen05.mma through an elected lane; one thread streams weights by TMA.
//   * im2col is never materialised and nothing is gathered: activations live in shared memory FLATTENED with zero
//     halos — 8^3: q = d*72 + h*8 + w (a ninth all-zero row block per d slab, zero slabs around the leaf);
//     4^3: q = d*20 + h*4 + w; the stride-2 conv in its space-to-depth form (2x2x2 taps over a 5^3 grid of 8 parity
//     classes x 16 channels) — channels-last in 8-channel planes, which IS the canonical no-swizzle K-major UMMA
//     layout.  A filter tap is a shifted start address of the same shared-memory descriptor.
//   * the three kw taps of a 3x3x3 conv are concatenated along N (one A read serves three taps, N = 96/48 or
//     192/96); the kw shift is applied in the epilogue as a lane shuffle that never crosses a warp (w = lane & 7 or
//     lane & 3), which also supplies the zero padding along w.
//   * per-leaf reductions (GroupNorm statistics, channel attention) are warp shuffles + one 512-thread named barrier.
//   * the tensor core truncates its fp32 accumulator after every MMA (measured: a bias toward zero that grows with
//     the chain length); the 64-step chain of the stride-2 conv is split over four accumulators added in the epilogue.
//   * pre.0 (Cin = 1, K = 27) stays on FFMA in fp32 (raw voxel values are unbounded; it is 1.4 % of the MACs).
//   * proj + VQ: proj feeds nothing but the codebook distances, so it is folded into the codebook on the host
//     (encode_tc_stream.hpp): split-fp16 tensor-core scores x.M_k for all 256 codes straight from the attention output
//     (K = 32 instead of 128 + the proj GEMM), with a rigorous error bound, then — only for rows whose shortlist holds
//     more than one code (near-ties, ~1 % of the rows) — z = W x + b in fp32 and exact re-scoring with the reference's
//     formula and tie-break: the two-stage scheme of encode_fp32.cu with a 500x tighter first stage.
//   * weights (548 KB per leaf as fp16 hi/lo planes, L2-resident) stream through a 4 x 12 KB shared-memory ring as
//     50 units by 1-D TMA bulk copies; ring slots are released by tcgen05.commit.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "encode_tc.cuh"
#include "leaf_ops.cuh"
#include "ptx_utils.cuh"

namespace vqvdb {

namespace {

constexpr int kGroups = 4;                        // row warps per TMEM lane quadrant: they split the channels of a row
constexpr int kRowWarps = 4 * kGroups;            // 16
constexpr int kRowThreads = 32 * kRowWarps;       // 512
constexpr int kIssuerWarp = kRowWarps, kProducerWarp = kRowWarps + 1;
constexpr int kThreads = kRowThreads + 64;        // + MMA issuer warp + TMA producer warp
constexpr int kStages = kEncTcRingStages;
constexpr uint32_t kStageBytes = kEncTcStageBytes;
constexpr float kLoScale = 2048.f, kLoInv = 1.f / 2048.f;
constexpr int kDownChains = 4;                    // independent accumulators of `down` (see the issuer)
#ifndef VQVDB_ENC_PRE0_SPLIT_AT
#define VQVDB_ENC_PRE0_SPLIT_AT 3   // measured: 0 (all under `down`) 6.422 M leaves/s, 1: 6.461 M, 2: 6.497 M, 3: 6.531 M
#endif
constexpr int kPre0Split = VQVDB_ENC_PRE0_SPLIT_AT;  // kd taps of the next leaf's pre.0 that run inside conv2's MMA wait
// pre.0's FMAs as packed fp32 pairs (FFMA2: one instruction per two channels, each half an IEEE fma: bit-identical results)
#ifndef VQVDB_ENC_PRE0_FFMA2
#define VQVDB_ENC_PRE0_FFMA2 1
#endif
// (Registers: __launch_bounds__(576, 1) makes ptxas stop at 96 per thread with 88 bytes of spills; asking for 112 with
// __maxnreg__ compiles to 52 bytes of spills but the launch fails — "too many resources requested" — on the B200.)
// offsets (floats) of the per-channel parameter vectors staged in shared memory: a global (L2) load at the head of
// every epilogue costs ~300 cycles of exposed latency, there being almost no L1 left beside 222 KB of shared memory
namespace par {
constexpr int pre_b = 0, pre_gn_w = 16, pre_gn_b = 32, r16_gn1_w = 48, r16_gn1_b = 64, r16_c1_b = 80, r16_gn2_w = 96, r16_gn2_b = 112,
              r16_c2_b = 128, down_b = 144, r32_gn1_w = 176, r32_gn1_b = 208, r32_c1_b = 240, r32_gn2_w = 272, r32_gn2_b = 304,
              r32_c2_b = 336, proj_b = 368, fc0 = 496, fc2 = 752;  // 1008 floats in all
}

// ---- shared-memory activation buffers (all UMMA A operands: [precision][8-channel plane][row][16 B]) ----
constexpr int kA8Margin = 80, kA8Rows = 800;                    // 8^3, 16 channels: rows -80 .. 719 around q = d*72 + h*8 + w
constexpr uint32_t kA8Plane = kA8Rows * 16, kA8Prec = 2 * kA8Plane, kA8Bytes = 2 * kA8Prec;      // 51 200
constexpr int kYRows = 160;                                     // space-to-depth input of `down`: q' = md*25 + mh*5 + mw
constexpr uint32_t kYPlane = kYRows * 16, kYPrec = 16 * kYPlane, kYBytes = 2 * kYPrec;           // 81 920
constexpr int kHMargin = 24, kHRows = 176;                      // 4^3, 32 channels: rows -24 .. 151 around q = d*20 + h*4 + w
constexpr uint32_t kHPlane = kHRows * 16, kHPrec = 4 * kHPlane, kHBytes = 2 * kHPrec;            // 22 528
// Overlays of Y, which is idle between the `down` MMAs of a leaf and the conv2 epilogue of the next one (and cleared in
// between): proj.weight transposed [32][128] fp32, the attention output rows [64][36] fp32, and — near-tie rows only —
// the fp32 z rows [64][132] for the exact re-scoring
constexpr uint32_t kWpBytes = 32 * 128 * 4;                     // 16 384
constexpr int kXsPitch = 36;
constexpr uint32_t kXsBytes = 64 * kXsPitch * 4;                // 9 216
constexpr int kZsPitch = 132;
constexpr uint32_t kZsBytes = 64 * kZsPitch * 4;                // 33 792
// the VQ GEMM's A operand: the attention output split into fp16 hi / lo planes, 64 DENSE rows (latent positions) stored
// twice — MMA rows 64..127 repeat rows 0..63, so all four TMEM lane quadrants (= all four SM sub-partitions) share the
// score pass: rows 0..63 take codes 0..127, rows 64..127 codes 128..255
constexpr uint32_t kXhPlane = 128 * 16, kXhPrec = 4 * kXhPlane, kXhBytes = 2 * kXhPrec;          // 16 384
constexpr int kX32Pitch = 36;

constexpr uint32_t kOffRing = 0;
constexpr uint32_t kOffA8 = kOffRing + kStages * kStageBytes;   // 49 152
constexpr uint32_t kOffY = kOffA8 + kA8Bytes;                   // 100 352
constexpr uint32_t kOffWp = kOffY;                              // Y is cleared again once the leaf's indices are out
constexpr uint32_t kOffXs = kOffWp + kWpBytes;
constexpr uint32_t kOffZs = kOffXs + kXsBytes;
constexpr uint32_t kOffXh = kOffZs + kZsBytes;
constexpr uint32_t kOffH = kOffY + kYBytes;                     // 182 272
constexpr uint32_t kOffIn = kOffH + kHBytes;                    // in_halo [10][10][10] fp32 (4096 reserved)
constexpr uint32_t kOffPreW = kOffIn + 4096;                    // pre.0 weights [27][16] fp32
constexpr uint32_t kOffX32 = kOffPreW + 1728;                   // residual of the 4^3 block [64][36] fp32; later the VQ exchange
constexpr uint32_t kOffRed = kOffX32 + 64 * kX32Pitch * 4;      // GroupNorm partials [2 slots][16 warps][2]
constexpr uint32_t kOffAtt = kOffRed + 256;                     // attention: part [16 warps][8], hid [8], scale [32]
constexpr uint32_t kOffCb = kOffAtt + (128 + 8 + 32) * 4;       // emb_sq [256], fold_esq [256], fold_norm [256]
constexpr uint32_t kOffPar = kOffCb + 3088;                     // fold_norm has a 257th entry: its maximum                     // per-channel parameter vectors (ParOff), 1008 floats
constexpr uint32_t kOffXq = kOffPar + 1008 * 4;                 // `down` epilogue: lane-0 values handed to the previous quadrant [4][4][8]
constexpr uint32_t kOffBar = kOffXq + 512;                      // mbarriers
constexpr int kConvGroups = kEncTcConvGroups;                   // tile groups of an 8^3 conv, each with its own completion barrier
constexpr uint32_t kNumBars = 2 * kStages + 3 + kConvGroups + 1;  // ... + the leaf-staging barrier
constexpr uint32_t kOffTmemSlot = kOffBar + kNumBars * 8;
// Leaf staging.  Default (0): 128 row threads issue one coalesced 128-bit ld.global.cs each for the NEXT leaf at the top of a
// leaf's loop and scatter the registers into the haloed fp32 buffer pre.0 reads when its turn comes.  VQVDB_ENC_LEAF_TMA=1:
// the producer warp requests the next leaf (2 KB, one cp.async.bulk) into a staging buffer and the row threads pick it up
// with 128-bit shared loads.  Both are built and give identical indices; measured on 592 k leaves the TMA form is 2 %
// slower (6.38 M vs 6.51 M leaves/s, profiles/r2_encode_experiments.txt) — the copy is 2 KB per 45 k cycles, nothing a bulk
// engine can speed up, and it adds a barrier wait and a shared-memory round trip to the row threads' critical path.
#ifndef VQVDB_ENC_LEAF_TMA
#define VQVDB_ENC_LEAF_TMA 0
#endif
constexpr uint32_t kOffLeafStage = (kOffTmemSlot + 16 + 15) & ~15u;
constexpr uint32_t kSmemBytes = kOffLeafStage + 2048;
static_assert(kOffXh + kXhBytes <= kOffY + kYBytes && kOffXh % 16 == 0, "the VQ overlays fit inside the Y region");
static_assert(kSmemBytes <= 227 * 1024, "encode_tc smem budget");
static_assert(kOffBar % 8 == 0 && kOffA8 % 1024 == 0 && kOffY % 1024 == 0 && kOffH % 1024 == 0, "alignment");

__device__ __forceinline__ uint32_t bar_w_full(uint32_t bars, uint32_t s) { return bars + s * 8; }
__device__ __forceinline__ uint32_t bar_w_empty(uint32_t bars, uint32_t s) { return bars + (kStages + s) * 8; }
__device__ __forceinline__ uint32_t bar_a_ready(uint32_t bars) { return bars + 2 * kStages * 8; }
// Accumulator hand-over barriers.  down / res32.c1 / res32.c2 / VQ use two barriers alternately (4 hand-overs per leaf): a
// waiter may lag a barrier by one phase only.  The tile groups of an 8^3 conv are committed back to back — with the next
// leaf's conv1 running under the VQ of this one, all of its commits can land before the row threads get to their first
// wait — so every group has its own barrier (two phases per leaf: conv1, conv2).
__device__ __forceinline__ uint32_t bar_d_full(uint32_t bars, uint32_t which) { return bars + (2 * kStages + 1 + which) * 8; }
__device__ __forceinline__ uint32_t bar_d8_full(uint32_t bars, uint32_t group) { return bars + (2 * kStages + 3 + group) * 8; }
__device__ __forceinline__ uint32_t bar_leaf(uint32_t bars) { return bars + (2 * kStages + 3 + kConvGroups) * 8; }

// ---- tcgen05 wrappers ----
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// instruction descriptor: D = f32, A/B = f16, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t idesc_f16(uint32_t n) { return (1u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24); }
// shared-memory descriptor, K-major, no swizzle: 8-row core matrices of 128 contiguous bytes; SBO = stride between
// 8-row groups, LBO = stride between the two 8-element K chunks of one K = 16 step
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
	return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
	       ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
	asm volatile(
	    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
	    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
	    "l"(a), "l"(b), "r"(id), "r"(acc)
	    : "memory");
}
// TMEM -> registers, 32 lanes x 4 / 8 / 16 consecutive columns; issue only (tmem_wait_ld() before use)
__device__ __forceinline__ void tmem_ld4_nowait(uint32_t taddr, float (&v)[4]) {
	uint32_t o[4];
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]) : "r"(taddr));
#pragma unroll
	for (int j = 0; j < 4; ++j) v[j] = __uint_as_float(o[j]);
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, float (&v)[8]) {
	uint32_t o[8];
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7])
	             : "r"(taddr));
#pragma unroll
	for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(o[j]);
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float (&v)[16]) {
	uint32_t o[16];
	asm volatile(
	    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
	    : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7]), "=r"(o[8]), "=r"(o[9]),
	      "=r"(o[10]), "=r"(o[11]), "=r"(o[12]), "=r"(o[13]), "=r"(o[14]), "=r"(o[15])
	    : "r"(taddr));
#pragma unroll
	for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(o[j]);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// One lane of a converged warp (the same one every time).
__device__ __forceinline__ bool elect_one() {
	uint32_t pred;
	asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
	return pred != 0;
}
__device__ __forceinline__ void row_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
	asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_shared_v2(uint32_t addr, uint2 v) {
	asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}

// two fp32 values -> packed fp16 hi pair and fp16 lo pair (v ~= hi + lo / 2048)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
	const __half2 hh = __floats2half2_rn(a, b);
	const float2 hf = __half22float2(hh);
	const __half2 ll = __floats2half2_rn((a - hf.x) * kLoScale, (b - hf.y) * kLoScale);
	hi = *reinterpret_cast<const uint32_t*>(&hh);
	lo = *reinterpret_cast<const uint32_t*>(&ll);
}

// The row threads hand the freshly written A operand (and the drained accumulators) to the MMA issuer.
__device__ __forceinline__ void signal_a_ready(uint32_t bars, int lane) {
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core
	tc_fence_before();
	__syncwarp();
	if (lane == 0) mbar_arrive(bar_a_ready(bars));
}

struct RowCtx {
	int warp, lane, g;      // g = channel group of this warp (warp >> 2); TMEM lane quadrant = warp & 3
	uint32_t bars, tlane;   // mbarrier base; TMEM base + (quadrant * 32 << 16)
	uint32_t d_count = 0;   // down / 4^3 / VQ accumulator hand-overs so far (parity of d_full)
	uint32_t d8_count = 0;  // 8^3 convs finished so far (parity of every d8_full)
	float* red;
};
// All 512 row threads park on the barrier themselves (hardware-assisted try_wait); electing one waiter per warp and
// holding the rest at a __syncwarp was measured 4 % slower (profiles/r2_encode_experiments.txt).
__device__ __forceinline__ void row_wait(const RowCtx&, uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }
__device__ __forceinline__ void wait_accumulator(RowCtx& rc) {
	row_wait(rc, bar_d_full(rc.bars, rc.d_count & 1u), (rc.d_count >> 1) & 1u);
	tc_fence_after();
	++rc.d_count;
}
// tile t of the current 8^3 conv: waits for its group when t is the group's first tile
__device__ __forceinline__ void wait_conv_tile(const RowCtx& rc, int t) {
#pragma unroll
	for (int g = 0; g < kConvGroups; ++g)
		if (t == enc_tc_group_first(g)) {
			row_wait(rc, bar_d8_full(rc.bars, g), rc.d8_count & 1u);
			tc_fence_after();
		}
}

// Sum NG (<= 2) per-thread values over the 128 row threads of this thread's channel group: warp shuffle, then the
// group's four warp partials through shared memory.
template <int NG>
__device__ __forceinline__ void group_allreduce(float (&s)[NG], const RowCtx& rc, int slot) {
#pragma unroll
	for (int i = 0; i < NG; ++i) s[i] = warp_sum(s[i]);
	float* red = rc.red + slot * 32;
	if (rc.lane == 0) {
#pragma unroll
		for (int i = 0; i < NG; ++i) red[rc.warp * 2 + i] = s[i];
	}
	// only the group's own four warps (one per SM sub-partition) meet: named barriers 2 .. 5, 128 threads
	asm volatile("bar.sync %0, 128;" ::"r"(2 + rc.g) : "memory");
	const float* r = red + rc.g * 8;  // warps 4g .. 4g+3
#pragma unroll
	for (int i = 0; i < NG; ++i) s[i] = (r[i] + r[2 + i]) + (r[4 + i] + r[6 + i]);
}

// GroupNorm statistics of register-resident values v[R][C] (rows with a clear bit in `valid` do not count); the
// thread's C channels are C / CPG whole groups.
template <int R, int C, int CPG>
__device__ __forceinline__ void gn_stats_regs(const float (&v)[R][C], uint32_t valid, float inv_cnt, const RowCtx& rc,
                                              float (&mean)[C / CPG], float (&rstd)[C / CPG]) {
	constexpr int NG = C / CPG;
	float s[NG];
#pragma unroll
	for (int i = 0; i < NG; ++i) s[i] = 0.f;
#pragma unroll
	for (int t = 0; t < R; ++t)
		if (valid & (1u << t)) {
#pragma unroll
			for (int c = 0; c < C; ++c) s[c / CPG] += v[t][c];
		}
	group_allreduce<NG>(s, rc, 0);
#pragma unroll
	for (int i = 0; i < NG; ++i) {
		mean[i] = s[i] * inv_cnt;
		s[i] = 0.f;
	}
#pragma unroll
	for (int t = 0; t < R; ++t)
		if (valid & (1u << t)) {
#pragma unroll
			for (int c = 0; c < C; ++c) {
				const float dv = v[t][c] - mean[c / CPG];
				s[c / CPG] = fmaf(dv, dv, s[c / CPG]);
			}
		}
	group_allreduce<NG>(s, rc, 1);
#pragma unroll
	for (int i = 0; i < NG; ++i) rstd[i] = 1.f / sqrtf(s[i] * inv_cnt + kGnEps);
}

// Flattened 8^3 row q -> voxel; false for halo / padding rows.
__device__ __forceinline__ bool row8(int q, int& d, int& h, int& w) {
	d = q / 72;
	const int rem = q - d * 72;
	h = rem >> 3;
	w = rem & 7;
	return q < 576 && h < 8;
}

// Output of one 128-row tile of a kw-concatenated 16-channel conv for channels 4g .. 4g+3: the tile's 96 accumulator
// columns are [hh kw0 | hh kw1 | hh kw2 | hl kw0 | hl kw1 | hl kw2] (16 each); folds hi/lo, applies the kw shift.
__device__ __forceinline__ void conv16_tile_out(uint32_t tcol, int g, int w, float (&o)[4]) {
	float hh0[4], hh1[4], hh2[4], hl0[4], hl1[4], hl2[4];
	tmem_ld4_nowait(tcol + g * 4, hh0);
	tmem_ld4_nowait(tcol + 16 + g * 4, hh1);
	tmem_ld4_nowait(tcol + 32 + g * 4, hh2);
	tmem_ld4_nowait(tcol + 48 + g * 4, hl0);
	tmem_ld4_nowait(tcol + 64 + g * 4, hl1);
	tmem_ld4_nowait(tcol + 80 + g * 4, hl2);
	tmem_wait_ld();
#pragma unroll
	for (int c = 0; c < 4; ++c) {
		const float p0 = fmaf(hl0[c], kLoInv, hh0[c]);
		const float p1 = fmaf(hl1[c], kLoInv, hh1[c]);
		const float p2 = fmaf(hl2[c], kLoInv, hh2[c]);
		const float up = __shfl_up_sync(0xffffffffu, p0, 1), dn = __shfl_down_sync(0xffffffffu, p2, 1);
		o[c] = ((w > 0 ? up : 0.f) + p1) + (w < 7 ? dn : 0.f);
	}
}
// Same for a 32-channel conv at 4^3, channels 8g .. 8g+7: 192 columns [hh kw0..2 (32 each) | hl kw0..2 (32 each)].
__device__ __forceinline__ void conv32_tile_out(uint32_t tcol, int g, int w, float (&o)[8]) {
	float hh0[8], hh1[8], hh2[8], hl0[8], hl1[8], hl2[8];
	tmem_ld8_nowait(tcol + g * 8, hh0);
	tmem_ld8_nowait(tcol + 32 + g * 8, hh1);
	tmem_ld8_nowait(tcol + 64 + g * 8, hh2);
	tmem_ld8_nowait(tcol + 96 + g * 8, hl0);
	tmem_ld8_nowait(tcol + 128 + g * 8, hl1);
	tmem_ld8_nowait(tcol + 160 + g * 8, hl2);
	tmem_wait_ld();
#pragma unroll
	for (int c = 0; c < 8; ++c) {
		const float p0 = fmaf(hl0[c], kLoInv, hh0[c]);
		const float p1 = fmaf(hl1[c], kLoInv, hh1[c]);
		const float p2 = fmaf(hl2[c], kLoInv, hh2[c]);
		const float up = __shfl_up_sync(0xffffffffu, p0, 1), dn = __shfl_down_sync(0xffffffffu, p2, 1);
		o[c] = ((w > 0 ? up : 0.f) + p1) + (w < 3 ? dn : 0.f);
	}
}

// channels 4g .. 4g+3 of one 8^3 row -> hi and lo planes (8 bytes each) at byte address `a` (plane, row, half chunk)
__device__ __forceinline__ void store_split4(uint32_t a, uint32_t prec_stride, const float (&v)[4]) {
	uint2 hi, lo;
	split2(v[0], v[1], hi.x, lo.x);
	split2(v[2], v[3], hi.y, lo.y);
	st_shared_v2(a, hi);
	st_shared_v2(a + prec_stride, lo);
}
// channels 8g .. 8g+7 of one 4^3 row -> one 16-byte chunk of the hi plane and one of the lo plane
__device__ __forceinline__ void store_split8(uint32_t a, uint32_t prec_stride, const float (&v)[8]) {
	uint4 hi, lo;
	split2(v[0], v[1], hi.x, lo.x);
	split2(v[2], v[3], hi.y, lo.y);
	split2(v[4], v[5], hi.z, lo.z);
	split2(v[6], v[7], hi.w, lo.w);
	st_shared_v4(a, hi);
	st_shared_v4(a + prec_stride, lo);
}

template <bool kProf> __device__ __forceinline__ long long prof_clock() { return kProf ? clock64() : 0; }

template <bool kProf>
__global__ void __launch_bounds__(kThreads, 1)
encode_tc_kernel(const EncoderWeights w, const EncoderTcStream ws, const float* __restrict__ leaves, int64_t n_leaves,
                 uint8_t* __restrict__ indices, int tap_stage, float* __restrict__ tap_out) {
	extern __shared__ __align__(1024) uint8_t smem[];
	const uint32_t s_base = smem_u32(smem);
	const uint32_t ring = s_base + kOffRing, bars = s_base + kOffBar;
	const uint32_t a8 = s_base + kOffA8, yb = s_base + kOffY, hb = s_base + kOffH, xh = s_base + kOffXh;
	float* in_halo = reinterpret_cast<float*>(smem + kOffIn);
	float* s_prew = reinterpret_cast<float*>(smem + kOffPreW);
	float* x32s = reinterpret_cast<float*>(smem + kOffX32);
	float* att = reinterpret_cast<float*>(smem + kOffAtt);
	float* s_esq = reinterpret_cast<float*>(smem + kOffCb);
	float* s_esq2 = s_esq + 256;  // |e_k|^2 - 2 b.e_k (shortlist scores)
	float* s_mno = s_esq + 512;   // |M_k|, rounded up (shortlist bound)
	float* s_wp = reinterpret_cast<float*>(smem + kOffWp);
	float* xs = reinterpret_cast<float*>(smem + kOffXs);
	const float* __restrict__ sp_c = reinterpret_cast<const float*>(smem + kOffPar);
	float* sp = reinterpret_cast<float*>(smem + kOffPar);
	float* xq = reinterpret_cast<float*>(smem + kOffXq);
	float* zs = reinterpret_cast<float*>(smem + kOffZs);
	uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffTmemSlot);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

	// The first leaf's voxels are requested before anything else: for a small host-pointer call they come from pinned
	// host memory over PCIe (capi.cu: kZeroCopyLeaves), ~2 us away, and the setup below does not depend on them.
#if VQVDB_ENC_LEAF_TMA
	const uint32_t leaf_stage = s_base + kOffLeafStage;
	if (warp == kProducerWarp && lane == 0) {  // this thread owns the staging barrier: nobody else touches it before the block-wide sync below
		mbar_init(bar_leaf(bars), 1);
		mbar_fence_init();
		if ((int64_t)blockIdx.x < n_leaves) {
			mbar_arrive_expect_tx(bar_leaf(bars), 2048);
			tma_load_1d(leaf_stage, leaves + (int64_t)blockIdx.x * 512, 2048, bar_leaf(bars));
		}
	}
#else
	float4 v_first = make_float4(0.f, 0.f, 0.f, 0.f);
	if (tid < 128 && (int64_t)blockIdx.x < n_leaves) v_first = __ldcs(reinterpret_cast<const float4*>(leaves + (int64_t)blockIdx.x * 512) + tid);
#endif

	// ---- one-time setup: zero the operand buffers (halo rows stay zero for the whole kernel), small tables ----
	for (uint32_t i = tid; i < (kA8Bytes + kYBytes + kHBytes + 4096) / 16; i += kThreads)
		reinterpret_cast<uint4*>(smem + kOffA8)[i] = make_uint4(0, 0, 0, 0);
	for (int i = tid; i < 432; i += kThreads) s_prew[i] = __ldg(w.pre_w + i);
	for (int i = tid; i < 256; i += kThreads) {
		s_esq[i] = __ldg(w.emb_sq + i);
		s_esq2[i] = __ldg(w.fold_esq + i);
		s_mno[i] = __ldg(w.fold_norm + i);
		if (i == 0) s_mno[256] = __ldg(w.fold_norm + 256);
		sp[par::fc0 + i] = __ldg(w.fc0 + i);
		sp[par::fc2 + i] = __ldg(w.fc2 + i);
	}
	if (tid < 128) sp[par::proj_b + tid] = __ldg(w.proj_b + tid);
	if (tid < 32) {
		sp[par::down_b + tid] = __ldg(w.down_b + tid);
		sp[par::r32_gn1_w + tid] = __ldg(w.res32.gn1_w + tid);
		sp[par::r32_gn1_b + tid] = __ldg(w.res32.gn1_b + tid);
		sp[par::r32_c1_b + tid] = __ldg(w.res32.c1_b + tid);
		sp[par::r32_gn2_w + tid] = __ldg(w.res32.gn2_w + tid);
		sp[par::r32_gn2_b + tid] = __ldg(w.res32.gn2_b + tid);
		sp[par::r32_c2_b + tid] = __ldg(w.res32.c2_b + tid);
	}
	if (tid < 16) {
		sp[par::pre_b + tid] = __ldg(w.pre_b + tid);
		sp[par::pre_gn_w + tid] = __ldg(w.pre_gn_w + tid);
		sp[par::pre_gn_b + tid] = __ldg(w.pre_gn_b + tid);
		sp[par::r16_gn1_w + tid] = __ldg(w.res16.gn1_w + tid);
		sp[par::r16_gn1_b + tid] = __ldg(w.res16.gn1_b + tid);
		sp[par::r16_c1_b + tid] = __ldg(w.res16.c1_b + tid);
		sp[par::r16_gn2_w + tid] = __ldg(w.res16.gn2_w + tid);
		sp[par::r16_gn2_b + tid] = __ldg(w.res16.gn2_b + tid);
		sp[par::r16_c2_b + tid] = __ldg(w.res16.c2_b + tid);
	}
	if (tid == 0) {
		for (uint32_t s = 0; s < kStages; ++s) {
			mbar_init(bar_w_full(bars, s), 1);
			mbar_init(bar_w_empty(bars, s), 1);
		}
		mbar_init(bar_a_ready(bars), kRowWarps);
		mbar_init(bar_d_full(bars, 0), 1);
		mbar_init(bar_d_full(bars, 1), 1);
		for (int g = 0; g < kConvGroups; ++g) mbar_init(bar_d8_full(bars, g), 1);
		mbar_fence_init();
	}
	if (warp == kIssuerWarp) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
	}
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem = *tmem_slot;
	const int64_t my_leaves = blockIdx.x < n_leaves ? (n_leaves - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

	if (warp == kProducerWarp) {
		// ===================== TMA producer =====================
		if (lane == 0) {
			const uint32_t total = (uint32_t)(my_leaves * kEncTcUnits);
#pragma unroll 1
			for (uint32_t issued = 0; issued < total; ++issued) {
				const uint32_t s = issued % kStages, u = issued % kEncTcUnits;
				mbar_wait(bar_w_empty(bars, s), ((issued / kStages) & 1u) ^ 1u);
#if VQVDB_ENC_LEAF_TMA
				if (u == kStages) {
					// Request this CTA's next leaf.  The staging buffer still holds the current leaf until its front has read it;
					// the ring slot of unit kStages was just released by the MMAs of this leaf's unit 0, and a leaf's first MMA
					// starts only after its front (load, pre.0, GroupNorms) is complete.
					const int64_t nxt = (int64_t)(issued / kEncTcUnits) + 1;
					if (nxt < my_leaves) {
						mbar_arrive_expect_tx(bar_leaf(bars), 2048);
						tma_load_1d(leaf_stage, leaves + ((int64_t)blockIdx.x + nxt * gridDim.x) * 512, 2048, bar_leaf(bars));
					}
				}
#endif
				mbar_arrive_expect_tx(bar_w_full(bars, s), ws.bytes[u]);
				tma_load_1d(ring + s * kStageBytes, ws.units + ws.off[u], ws.bytes[u], bar_w_full(bars, s));
			}
		}
		__syncwarp();
	} else if (warp == kIssuerWarp) {
		// ===================== MMA issuer =====================
		// The WHOLE warp runs the issue loop and one elected lane executes the tcgen05 instructions: with warp-uniform
		// control flow the descriptors and loop state live in uniform registers, so an MMA costs the issuing warp a
		// handful of instructions.  (A single-lane branch keeps them in vector registers and moves them to uniform
		// registers before every MMA; sharing its scheduler with four busy row warps, that issuer could not feed the
		// tensor pipe: measured 93 - 165 cycles per MMA inside the kernel against 50 in isolation.)
#ifdef VQVDB_ENC_ISSUER_LANE0
		const bool leader = true;
		if (lane == 0) {
#else
		const bool leader = elect_one();
		{
#endif
			uint32_t unit = 0, a_count = 0, d_commits = 0;
			long long t_wait_a = 0, t_wait_w = 0, t_issue = 0;
			long long t_wait_w_phase[4] = {0, 0, 0, 0};  // kProf: weight waits per phase (8^3 convs, down, 4^3 convs, VQ)
			long long t_exec_phase[4] = {0, 0, 0, 0};    // kProf: operands handed over -> last MMA of the layer complete
			long long t_phase0 = 0;
			long long t_issue_phase[4] = {0, 0, 0, 0};   // kProf: time inside the issue loops, per phase
			uint32_t convs_done = 0;
			int w_phase = 0;
			const long long t_start = prof_clock<kProf>();
			const uint64_t a8_d = make_desc(a8 + kA8Margin * 16, kA8Plane, 128);
			const uint64_t y_d = make_desc(yb, kYPlane, 128);
			const uint64_t h_d = make_desc(hb + kHMargin * 16, kHPlane, 128);
			const uint64_t xh_d = make_desc(xh, kXhPlane, 128);
			auto wait_a = [&]() {
				const long long c0 = prof_clock<kProf>();
				mbar_wait(bar_a_ready(bars), a_count & 1u);
				tc_fence_after();
				++a_count;
				if (kProf) {
					t_phase0 = prof_clock<kProf>();
					t_wait_a += t_phase0 - c0;
				}
			};
			auto wait_w = [&]() -> uint32_t {
				const long long c0 = prof_clock<kProf>();
				const uint32_t s = unit % kStages;
				mbar_wait(bar_w_full(bars, s), (unit / kStages) & 1u);
				tc_fence_after();
				if (kProf) {
					const long long dt = prof_clock<kProf>() - c0;
					t_wait_w += dt;
					t_wait_w_phase[w_phase] += dt;
				}
				return ring + s * kStageBytes;
			};
			auto wait_w_ahead = [&](uint32_t ahead) -> uint32_t {  // the unit `ahead` positions after the current one
				const uint32_t un = unit + ahead, s = un % kStages;
				mbar_wait(bar_w_full(bars, s), (un / kStages) & 1u);
				tc_fence_after();
				return ring + s * kStageBytes;
			};
			auto commit_d = [&]() {
				if (leader) tc_commit(bar_d_full(bars, d_commits & 1u));
				if (kProf) {  // the issuer has nothing to do before the row threads have read this result anyway
					mbar_wait(bar_d_full(bars, d_commits & 1u), (d_commits >> 1) & 1u);
					t_exec_phase[w_phase] += prof_clock<kProf>() - t_phase0;
				}
				++d_commits;
			};
			auto release_w = [&]() {
				if (leader) tc_commit(bar_w_empty(bars, unit % kStages));
				++unit;
			};
#pragma unroll 1
			for (int64_t it = 0; it < my_leaves; ++it) {
				// ---- res16 conv1, conv2: 5 tiles x 9 (kd, kh) x {N = 96, N = 48}, in kConvGroups tile groups with their own
				//      completion signal: the row threads drain a group while the next group's MMAs run ----
				w_phase = 0;
#pragma unroll 1
				for (int layer = 0; layer < 2; ++layer) {
					wait_a();
#pragma unroll 1
					for (int tg = 0; tg < kConvGroups; ++tg) {
						const int t0 = enc_tc_group_first(tg), t1 = enc_tc_group_first(tg + 1);
#pragma unroll 1
						for (int kd = 0; kd < 3; ++kd) {
							const uint32_t wb = wait_w();
							const long long c0 = prof_clock<kProf>();
#pragma unroll 1
							for (int t = t0; t < t1; ++t) {
#pragma unroll
								for (int kh = 0; kh < 3; ++kh) {
									const int s = (kd - 1) * 72 + (kh - 1) * 8 + 128 * t;
									const uint64_t ad = a8_d + (uint64_t)(int64_t)s;
									const uint64_t bd = make_desc(wb + kh * 3072, 96 * 16, 128);
									if (leader) mma_ss(tmem + t * 96, ad, bd, idesc_f16(96), (kd > 0 || kh > 0) ? 1u : 0u);
									if (leader) mma_ss(tmem + t * 96 + 48, ad + (kA8Prec >> 4), bd, idesc_f16(48), 1u);
								}
							}
							release_w();
							if (kProf) {
								const long long dt = prof_clock<kProf>() - c0;
								t_issue += dt;
								t_issue_phase[w_phase] += dt;
							}
						}
						if (leader) tc_commit(bar_d8_full(bars, tg));
						if (kProf && tg == kConvGroups - 1) {
							mbar_wait(bar_d8_full(bars, tg), convs_done & 1u);
							t_exec_phase[0] += prof_clock<kProf>() - t_phase0;
						}
					}
					++convs_done;
				}
				// ---- down: 4 (td, th) tap pairs x 8 parity classes x {N = 128, N = 64}; the two tw taps of a pair are
				//      concatenated along N like the kw taps of the 3x3x3 convs (their row shift of 1 is applied in the
				//      epilogue).  The tensor core truncates its fp32 accumulator after every MMA, a bias that grows with
				//      the length of the accumulation chain; every tap pair therefore has its own accumulator
				//      (kDownChains chains of 8 steps) and the epilogue adds them in fp32. ----
				w_phase = 1;
				wait_a();
#pragma unroll 1
				constexpr int kPclPerUnit = 32 / kEncTcDownUnits;  // parity classes per weight unit: 4 (8 units) or 2 (16 units)
#pragma unroll 1
				for (int u = 0; u < kEncTcDownUnits; ++u) {  // u = (td, th) tap pair * units per pair + slice of the parity classes
					const uint32_t wb = wait_w();
					const long long c0 = prof_clock<kProf>();
					const int pair = u / (kEncTcDownUnits / 4), half = u % (kEncTcDownUnits / 4);
					const int s = (pair >> 1) * 25 + (pair & 1) * 5;
					const uint32_t dcol = tmem + (uint32_t)pair * 128;
#pragma unroll
					for (int pcl = 0; pcl < kPclPerUnit; ++pcl) {
						const uint64_t ad = y_d + (uint64_t)(s + (half * kPclPerUnit + pcl) * 2 * (int)(kYPlane >> 4));
						const uint64_t bd = make_desc(wb + pcl * 4096, 128 * 16, 128);
						if (leader) mma_ss(dcol, ad, bd, idesc_f16(128), (half > 0 || pcl > 0) ? 1u : 0u);
						if (leader) mma_ss(dcol + 64, ad + (kYPrec >> 4), bd, idesc_f16(64), 1u);
					}
					release_w();
					if (kProf) {
								const long long dt = prof_clock<kProf>() - c0;
								t_issue += dt;
								t_issue_phase[w_phase] += dt;
							}
				}
				commit_d();
				// ---- res32 conv1, conv2: 9 (kd, kh) x 2 k-steps x {N = 192, N = 96} ----
#pragma unroll 1
				w_phase = 2;
				for (int layer = 0; layer < 2; ++layer) {
					wait_a();
#pragma unroll 1
					for (int kk = 0; kk < 9; ++kk) {
						const uint32_t wb = wait_w();
						const long long c0 = prof_clock<kProf>();
						const int s = (kk / 3 - 1) * 20 + (kk % 3 - 1) * 4;
#pragma unroll
						for (int ks = 0; ks < 2; ++ks) {
							const uint64_t ad = h_d + (uint64_t)(int64_t)(s + ks * 2 * (int)(kHPlane >> 4));
							const uint64_t bd = make_desc(wb + ks * 6144, 192 * 16, 128);
							if (leader) mma_ss(tmem, ad, bd, idesc_f16(192), (kk > 0 || ks > 0) ? 1u : 0u);
							if (leader) mma_ss(tmem + 96, ad + (kHPrec >> 4), bd, idesc_f16(96), 1u);
						}
						release_w();
						if (kProf) {
								const long long dt = prof_clock<kProf>() - c0;
								t_issue += dt;
								t_issue_phase[w_phase] += dt;
							}
					}
					commit_d();
				}
				// ---- VQ scores straight from the attention output (proj folded into the codebook): 2 k-steps x
				//      {x_hi.M_hi -> cols 0..255 ; x_hi.M_lo + x_lo.M_hi -> cols 256..511}, N = 256 ----
				w_phase = 3;
				wait_a();
#pragma unroll 1
				for (int ks = 0; ks < 2; ++ks) {
					const uint32_t wb = wait_w();
					const uint32_t wl = kEncTcVqUnits == 2 ? wb + 8192 : wait_w_ahead(1);  // M_lo: second half of the unit, or the next unit
					const long long c0 = prof_clock<kProf>();
					const uint64_t ad = xh_d + (uint64_t)(ks * 2 * (int)(kXhPlane >> 4));
					const uint64_t bh = make_desc(wb, 256 * 16, 128), bl = make_desc(wl, 256 * 16, 128);
					if (leader) mma_ss(tmem, ad, bh, idesc_f16(256), ks > 0 ? 1u : 0u);
					if (leader) mma_ss(tmem + 256, ad, bl, idesc_f16(256), ks > 0 ? 1u : 0u);
					if (leader) mma_ss(tmem + 256, ad + (kXhPrec >> 4), bh, idesc_f16(256), 1u);
					release_w();
					if (kEncTcVqUnits == 4) release_w();
					if (kProf) {
								const long long dt = prof_clock<kProf>() - c0;
								t_issue += dt;
								t_issue_phase[w_phase] += dt;
							}
				}
				commit_d();
			}
			if (kProf && tap_out && leader) {
				float* o = tap_out + (size_t)blockIdx.x * 64 + 32;
				o[0] = (float)t_wait_a; o[1] = (float)t_wait_w; o[2] = (float)t_issue; o[3] = (float)(prof_clock<kProf>() - t_start);
#pragma unroll
				for (int i = 0; i < 4; ++i) o[4 + i] = (float)t_wait_w_phase[i];
#pragma unroll
				for (int i = 0; i < 4; ++i) o[8 + i] = (float)t_exec_phase[i];
#pragma unroll
				for (int i = 0; i < 4; ++i) o[12 + i] = (float)t_issue_phase[i];
			}
		}
		__syncwarp();
	} else {
		// ===================== row threads: FFMA pre.0, epilogues, VQ =====================
		// warp = g * 4 + quadrant: TMEM row = quadrant * 32 + lane; channel group g of kGroups.
		RowCtx rc;
		rc.warp = warp; rc.lane = lane; rc.g = warp >> 2;
		rc.bars = bars;
		rc.tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
		rc.red = reinterpret_cast<float*>(smem + kOffRed);
		const int g = rc.g, row = (warp & 3) * 32 + lane;
		long long prof[20];
#pragma unroll
		for (int i = 0; i < 20; ++i) prof[i] = 0;
		long long pc0 = prof_clock<kProf>();
		auto lap = [&](int slot) {
			if (kProf) {
				const long long c = prof_clock<kProf>();
				prof[slot] += c - pc0;
				pc0 = c;
			}
		};
		// 4^3 row of this thread (res32 / proj / VQ tiles): q4 = row
		const int d4 = row / 20, h4 = (row - d4 * 20) >> 2, w4 = row & 3;
		const bool valid4 = row < 80 && h4 < 4;
		const int p4 = d4 * 16 + h4 * 4 + w4;
		// `down` output row of this thread: q' = row over the 5^3 grid
		const int jd = row / 25, jh = (row - jd * 25) / 5, jw = row % 5;
		const bool validd = row < 100 && jh < 4 && jw < 4;
		const int pd = jd * 16 + jh * 4 + jw;
		// VQ exchange arrays (overlay the x32 staging, dead by then): [4 groups][128 rows]
		float* vq_umin = x32s;        // near-tie rows: the re-scored minima
		float* vq_lo1 = x32s + 512;   // smallest / second-smallest lower bound of a thread's 64 codes
		float* vq_lo2 = x32s + 1024;
		int* vq_k1 = reinterpret_cast<int*>(x32s + 1536);  // code of lo1; later the re-scored arg-min
		static_assert(4 * 512 * 4 <= 64 * kX32Pitch * 4, "VQ exchange arrays fit in the dead x32 staging area");
		// the four per-group partials of |x|^2 (and of |z|^2) of a row live in the padding of its xs (zs) row
		static_assert(kXsPitch >= 32 + 4 && kZsPitch >= 128 + 4, "row padding holds the four partial sums");

		// Rows of this thread in the five 8^3 tiles (the same for every leaf).
		uint32_t valid8 = 0;
		int base8[5];
#pragma unroll
		for (int t = 0; t < 5; ++t) {
			int d, h, w8;
			const bool ok = row8(t * 128 + row, d, h, w8);
			valid8 |= ok ? (1u << t) : 0u;
			base8[t] = ok ? d * 100 + h * 10 + w8 : 0;
		}
		const uint32_t a8_mine = a8 + (uint32_t)(g >> 1) * kA8Plane + (uint32_t)(kA8Margin + row) * 16 + (uint32_t)(g & 1) * 8;

		// ---- the front of a leaf, software-pipelined into the MMA waits of the previous leaf ----
		// (a) stage the leaf (2048 B, 128-bit coalesced) into the haloed fp32 buffer
#if VQVDB_ENC_LEAF_TMA
		auto front_load = [&](int64_t ordinal) {  // ordinal = index of the leaf among this CTA's leaves
			if (tid < 128) {
				mbar_wait(bar_leaf(bars), (uint32_t)ordinal & 1u);
				float4 v;
				asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(leaf_stage + (uint32_t)tid * 16));
#else
		auto front_load = [&](const float4& v) {
			if (tid < 128) {
#endif
				const int p = tid * 4, d = p >> 6, h = (p >> 3) & 7, w0 = p & 7;
				float* dst = in_halo + (d + 1) * 100 + (h + 1) * 10 + w0 + 1;
				dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
			}
			row_bar();
		};
		// (b) pre.0: Conv3d(1,16,k3) on FFMA, channels 4g..4g+3, + bias
		// kd taps [kd0, kd1) of the filter; the first part zeroes the accumulators, the last one adds the bias
#if VQVDB_ENC_PRE0_FFMA2
		// The accumulators are packed fp32 pairs (FFMA2, channel pairs 4g + {0,1}, 4g + {2,3}; each half an IEEE fma).
		auto front_pre0_part = [&](float (&x)[5][4], int kd0, int kd1) {
			uint64_t x2[5][2];
#pragma unroll
			for (int t = 0; t < 5; ++t) {
				x2[t][0] = kd0 == 0 ? 0ull : pack_f32x2(x[t][0], x[t][1]);
				x2[t][1] = kd0 == 0 ? 0ull : pack_f32x2(x[t][2], x[t][3]);
			}
#pragma unroll 1
			for (int kd = kd0; kd < kd1; ++kd) {
#pragma unroll
				for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
					for (int kw = 0; kw < 3; ++kw) {
						const int tap = (kd * 3 + kh) * 3 + kw;
						const float4 f = *reinterpret_cast<const float4*>(s_prew + tap * 16 + g * 4);
						const uint64_t w01 = pack_f32x2(f.x, f.y), w23 = pack_f32x2(f.z, f.w);
#pragma unroll
						for (int t = 0; t < 5; ++t) {
							const float xv = in_halo[base8[t] + kd * 100 + kh * 10 + kw];
							const uint64_t xx = pack_f32x2(xv, xv);
							x2[t][0] = fma_f32x2(xx, w01, x2[t][0]);
							x2[t][1] = fma_f32x2(xx, w23, x2[t][1]);
						}
					}
				}
			}
#pragma unroll
			for (int t = 0; t < 5; ++t) {
				unpack_f32x2(x2[t][0], x[t][0], x[t][1]);
				unpack_f32x2(x2[t][1], x[t][2], x[t][3]);
			}
			if (kd1 == 3) {
#pragma unroll
				for (int c = 0; c < 4; ++c) {
					const float b = sp_c[par::pre_b + g * 4 + c];
#pragma unroll
					for (int t = 0; t < 5; ++t) x[t][c] += b;
				}
			}
		};
#else
		auto front_pre0_part = [&](float (&x)[5][4], int kd0, int kd1) {
			if (kd0 == 0) {
#pragma unroll
				for (int t = 0; t < 5; ++t)
#pragma unroll
					for (int c = 0; c < 4; ++c) x[t][c] = 0.f;
			}
#pragma unroll 1
			for (int kd = kd0; kd < kd1; ++kd) {
#pragma unroll
				for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
					for (int kw = 0; kw < 3; ++kw) {
						const int tap = (kd * 3 + kh) * 3 + kw;
						const float4 f = *reinterpret_cast<const float4*>(s_prew + tap * 16 + g * 4);
						const float wv[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
						for (int t = 0; t < 5; ++t) {
							const float xv = in_halo[base8[t] + kd * 100 + kh * 10 + kw];
#pragma unroll
							for (int c = 0; c < 4; ++c) x[t][c] = fmaf(xv, wv[c], x[t][c]);
						}
					}
				}
			}
			if (kd1 == 3) {
#pragma unroll
				for (int c = 0; c < 4; ++c) {
					const float b = sp_c[par::pre_b + g * 4 + c];
#pragma unroll
					for (int t = 0; t < 5; ++t) x[t][c] += b;
				}
			}
		};
#endif
		auto front_pre0 = [&](float (&x)[5][4]) { front_pre0_part(x, 0, 3); };
		// (c) pre.1: GroupNorm(4,16) + ReLU -> x (the residual, kept in registers)
		auto front_gn_pre1 = [&](float (&x)[5][4]) {
			float mean[1], rstd[1];  // this thread's 4 channels are exactly GroupNorm group g
			gn_stats_regs<5, 4, 4>(x, valid8, 1.f / 2048.f, rc, mean, rstd);
#pragma unroll
			for (int c = 0; c < 4; ++c) {
				const float ga = sp_c[par::pre_gn_w + g * 4 + c], be = sp_c[par::pre_gn_b + g * 4 + c];
#pragma unroll
				for (int t = 0; t < 5; ++t) x[t][c] = relu_f((x[t][c] - mean[0]) * rstd[0] * ga + be);
			}
		};
		// (d) res16.gn1 + ReLU -> A8 (conv1 input)
		auto front_gn1_to_a8 = [&](const float (&x)[5][4], int64_t leaf) {
			float mean[2], rstd[2];
			gn_stats_regs<5, 4, 2>(x, valid8, 1.f / 1024.f, rc, mean, rstd);
			float ga[4], be[4];
#pragma unroll
			for (int c = 0; c < 4; ++c) {
				ga[c] = sp_c[par::r16_gn1_w + g * 4 + c];
				be[c] = sp_c[par::r16_gn1_b + g * 4 + c];
			}
#pragma unroll
			for (int t = 0; t < 5; ++t) {
				if (valid8 & (1u << t)) {
					float a[4];
#pragma unroll
					for (int c = 0; c < 4; ++c) a[c] = relu_f((x[t][c] - mean[c >> 1]) * rstd[c >> 1] * ga[c] + be[c]);
					store_split4(a8_mine + t * 2048, kA8Prec, a);
					if (tap_stage == 0) {
						int d, h, w8;
						row8(t * 128 + row, d, h, w8);
#pragma unroll
						for (int c = 0; c < 4; ++c) tap_out[(leaf * 16 + g * 4 + c) * 512 + d * 64 + h * 8 + w8] = x[t][c];
					}
				}
			}
		};

		float xn[5][4];  // x of the leaf whose front is in progress
		if (my_leaves > 0) {
#if VQVDB_ENC_LEAF_TMA
			front_load(0);
#else
			front_load(v_first);
#endif
			front_pre0(xn);
			front_gn_pre1(xn);
			front_gn1_to_a8(xn, blockIdx.x);
			signal_a_ready(bars, lane);
		}
		lap(1);

#pragma unroll 1
		for (int64_t it = 0; it < my_leaves; ++it) {
			const int64_t leaf = blockIdx.x + it * gridDim.x;
			const bool has_next = it + 1 < my_leaves;
#if !VQVDB_ENC_LEAF_TMA
			float4 nx_v = make_float4(0.f, 0.f, 0.f, 0.f);  // the next leaf's voxels: in flight while this leaf's 8^3 layers run
			if (has_next && tid < 128) nx_v = __ldcs(reinterpret_cast<const float4*>(leaves + (leaf + gridDim.x) * 512) + tid);
#endif
			float xr[5][4];
#pragma unroll
			for (int t = 0; t < 5; ++t)
#pragma unroll
				for (int c = 0; c < 4; ++c) xr[t][c] = xn[t][c];

			// ---- conv1 epilogue: + bias, res16.gn2 + ReLU -> A8 (conv2 input) ----
			wait_conv_tile(rc, 0);
			lap(3);
			{
				float v[5][4];
				float bs[4];
#pragma unroll
				for (int c = 0; c < 4; ++c) bs[c] = sp_c[par::r16_c1_b + g * 4 + c];
#pragma unroll
				for (int t = 0; t < 5; ++t) {
					if (t > 0) wait_conv_tile(rc, t);  // the later tile groups
					conv16_tile_out(rc.tlane + t * 96, g, lane & 7, v[t]);
#pragma unroll
					for (int c = 0; c < 4; ++c) v[t][c] += bs[c];
					if (tap_stage == 6 && (valid8 & (1u << t))) {
						int d, h, w8;
						row8(t * 128 + row, d, h, w8);
#pragma unroll
						for (int c = 0; c < 4; ++c) tap_out[(leaf * 16 + g * 4 + c) * 512 + d * 64 + h * 8 + w8] = v[t][c];
					}
				}
				++rc.d8_count;
				lap(13);
				float mean[2], rstd[2];
				gn_stats_regs<5, 4, 2>(v, valid8, 1.f / 1024.f, rc, mean, rstd);
				lap(14);
				float ga[4], be[4];
#pragma unroll
				for (int c = 0; c < 4; ++c) {
					ga[c] = sp_c[par::r16_gn2_w + g * 4 + c];
					be[c] = sp_c[par::r16_gn2_b + g * 4 + c];
				}
#pragma unroll
				for (int t = 0; t < 5; ++t) {
					if (valid8 & (1u << t)) {
#pragma unroll
						for (int c = 0; c < 4; ++c) v[t][c] = relu_f((v[t][c] - mean[c >> 1]) * rstd[c >> 1] * ga[c] + be[c]);
						store_split4(a8_mine + t * 2048, kA8Prec, v[t]);
					}
				}
			}
			signal_a_ready(bars, lane);
			lap(4);
#if VQVDB_ENC_PRE0_SPLIT_AT > 0
			// conv2's first tile group takes ~3 k cycles in which the row threads would only wait: stage the next leaf and
			// run the first kPre0Split kd-slices of its pre.0 here; the rest runs under the `down` MMAs
			if (has_next) {
#if VQVDB_ENC_LEAF_TMA
				front_load(it + 1);
#else
				front_load(nx_v);
#endif
				front_pre0_part(xn, 0, kPre0Split);
			}
			lap(0);
#endif

			// ---- conv2 epilogue: x2 = x + 0.1 (conv2 + b) -> Y, the space-to-depth input of `down` ----
			wait_conv_tile(rc, 0);
			lap(5);
			{
				float bs[4];
#pragma unroll
				for (int c = 0; c < 4; ++c) bs[c] = sp_c[par::r16_c2_b + g * 4 + c];
#pragma unroll
				for (int t = 0; t < 5; ++t) {
					if (t > 0) wait_conv_tile(rc, t);  // the later tile groups
					float o[4];
					conv16_tile_out(rc.tlane + t * 96, g, lane & 7, o);
					int d, h, w8;
					const bool ok = row8(t * 128 + row, d, h, w8);
					if (ok) {
#pragma unroll
						for (int c = 0; c < 4; ++c) o[c] = xr[t][c] + kResScale * (o[c] + bs[c]);
						if (tap_stage == 1) {
#pragma unroll
							for (int c = 0; c < 4; ++c) tap_out[(leaf * 16 + g * 4 + c) * 512 + d * 64 + h * 8 + w8] = o[c];
						}
						const int pcl = (((d + 1) & 1) << 2) | (((h + 1) & 1) << 1) | ((w8 + 1) & 1);
						const int qy = ((d + 1) >> 1) * 25 + ((h + 1) >> 1) * 5 + ((w8 + 1) >> 1);
						store_split4(yb + (uint32_t)(pcl * 2 + (g >> 1)) * kYPlane + (uint32_t)qy * 16 + (uint32_t)(g & 1) * 8, kYPrec, o);
					}
				}
				++rc.d8_count;
			}
			signal_a_ready(bars, lane);
			lap(6);
#if VQVDB_ENC_PRE0_SPLIT_AT > 0
			if (has_next && kPre0Split < 3) front_pre0_part(xn, kPre0Split, 3);
#else
			if (has_next) {  // the `down` MMAs run for ~6 k cycles: stage the next leaf and run its pre.0
#if VQVDB_ENC_LEAF_TMA
				front_load(it + 1);
#else
				front_load(nx_v);
#endif
				front_pre0(xn);
			}
#endif
			lap(0);

			// ---- down epilogue: sum the accumulation chains, + bias -> x32 (residual, to shared memory) ; res32.gn1 + ReLU -> H32 ----
			wait_accumulator(rc);
			lap(7);
			// the `down` MMAs were the last readers of Y: proj.weight (transposed, fp32) goes there for the near-tie rows of the VQ
			const float4 wp0 = __ldg(reinterpret_cast<const float4*>(w.proj_w) + tid);
			const float4 wp1 = __ldg(reinterpret_cast<const float4*>(w.proj_w) + 512 + tid);
			{
				float v[1][8];
				{
					// chain c holds [hh tw0 | hh tw1 | hl tw0 | hl tw1] (32 columns each) in columns c*128 ..
					float h0[8], h1[8], l0[8], l1[8];
#pragma unroll
					for (int c = 0; c < kDownChains; ++c) {
						float a0[8], a1[8], b0[8], b1[8];
						tmem_ld8_nowait(rc.tlane + c * 128 + g * 8, a0);
						tmem_ld8_nowait(rc.tlane + c * 128 + 32 + g * 8, a1);
						tmem_ld8_nowait(rc.tlane + c * 128 + 64 + g * 8, b0);
						tmem_ld8_nowait(rc.tlane + c * 128 + 96 + g * 8, b1);
						tmem_wait_ld();
#pragma unroll
						for (int i = 0; i < 8; ++i) {
							h0[i] = c == 0 ? a0[i] : h0[i] + a0[i];
							h1[i] = c == 0 ? a1[i] : h1[i] + a1[i];
							l0[i] = c == 0 ? b0[i] : l0[i] + b0[i];
							l1[i] = c == 0 ? b1[i] : l1[i] + b1[i];
						}
					}
					// out[q'] = P_tw0[q'] + P_tw1[q' + 1]: lane + 1, or lane 0 of the next quadrant through shared memory
#pragma unroll
					for (int i = 0; i < 8; ++i) {
						h0[i] = fmaf(l0[i], kLoInv, h0[i]);
						h1[i] = fmaf(l1[i], kLoInv, h1[i]);
					}
					if (lane == 0) {
#pragma unroll
						for (int i = 0; i < 8; ++i) xq[((warp & 3) * 4 + g) * 8 + i] = h1[i];
					}
					row_bar();
#pragma unroll
					for (int i = 0; i < 8; ++i) {
						float nx = __shfl_down_sync(0xffffffffu, h1[i], 1);
						if (lane == 31) nx = xq[((((warp & 3) + 1) & 3) * 4 + g) * 8 + i];  // row 127 never matters
						v[0][i] = (h0[i] + nx) + sp_c[par::down_b + g * 8 + i];
					}
				}
				if (validd) {
					*reinterpret_cast<float4*>(x32s + pd * kX32Pitch + g * 8) = make_float4(v[0][0], v[0][1], v[0][2], v[0][3]);
					*reinterpret_cast<float4*>(x32s + pd * kX32Pitch + g * 8 + 4) = make_float4(v[0][4], v[0][5], v[0][6], v[0][7]);
					if (tap_stage == 2) {
#pragma unroll
						for (int c = 0; c < 8; ++c) tap_out[(leaf * 32 + g * 8 + c) * 64 + pd] = v[0][c];
					}
				}
				float mean[2], rstd[2];
				gn_stats_regs<1, 8, 4>(v, validd ? 1u : 0u, 1.f / 256.f, rc, mean, rstd);
				if (validd) {
#pragma unroll
					for (int c = 0; c < 8; ++c)
						v[0][c] = relu_f((v[0][c] - mean[c >> 2]) * rstd[c >> 2] * sp_c[par::r32_gn1_w + g * 8 + c] + sp_c[par::r32_gn1_b + g * 8 + c]);
					store_split8(hb + (uint32_t)g * kHPlane + (uint32_t)(kHMargin + jd * 20 + jh * 4 + jw) * 16, kHPrec, v[0]);
				}
			}
			signal_a_ready(bars, lane);
			reinterpret_cast<float4*>(s_wp)[tid] = wp0;
			reinterpret_cast<float4*>(s_wp)[512 + tid] = wp1;
			lap(8);
			if (has_next) front_gn_pre1(xn);
			lap(1);

			// ---- res32 conv1 epilogue: + bias, gn2 + ReLU -> H32 ----
			const uint32_t h_mine = hb + (uint32_t)g * kHPlane + (uint32_t)(kHMargin + row) * 16;
			wait_accumulator(rc);
			lap(9);
			{
				float v[1][8];
				conv32_tile_out(rc.tlane, g, w4, v[0]);
#pragma unroll
				for (int c = 0; c < 8; ++c) v[0][c] += sp_c[par::r32_c1_b + g * 8 + c];
				if (valid4 && tap_stage == 7) {
#pragma unroll
					for (int c = 0; c < 8; ++c) tap_out[(leaf * 32 + g * 8 + c) * 64 + p4] = v[0][c];
				}
				float mean[2], rstd[2];
				gn_stats_regs<1, 8, 4>(v, valid4 ? 1u : 0u, 1.f / 256.f, rc, mean, rstd);
				if (valid4) {
#pragma unroll
					for (int c = 0; c < 8; ++c)
						v[0][c] = relu_f((v[0][c] - mean[c >> 2]) * rstd[c >> 2] * sp_c[par::r32_gn2_w + g * 8 + c] + sp_c[par::r32_gn2_b + g * 8 + c]);
					store_split8(h_mine, kHPrec, v[0]);
				}
			}
			signal_a_ready(bars, lane);
			lap(10);
			if (has_next) front_gn1_to_a8(xn, leaf + gridDim.x);  // A8 has been free since conv2's MMAs completed
			lap(2);

			// ---- res32 conv2 epilogue: x3 = x32 + 0.1 (conv2 + b) ; ChannelAttention(32) -> H32 (proj input) ----
			wait_accumulator(rc);
			lap(11);
			{
				float v[8];
				conv32_tile_out(rc.tlane, g, w4, v);
				if (valid4) {
					const float4 x0 = *reinterpret_cast<const float4*>(x32s + p4 * kX32Pitch + g * 8);
					const float4 x1 = *reinterpret_cast<const float4*>(x32s + p4 * kX32Pitch + g * 8 + 4);
					const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
					for (int c = 0; c < 8; ++c) v[c] = xv[c] + kResScale * (v[c] + sp_c[par::r32_c2_b + g * 8 + c]);
					if (tap_stage == 3) {
#pragma unroll
						for (int c = 0; c < 8; ++c) tap_out[(leaf * 32 + g * 8 + c) * 64 + p4] = v[c];
					}
				}
				float* part = att;          // [16 warps][8]
				float* hid = att + 128;     // [8]
				float* scale = att + 136;   // [32]
#pragma unroll
				for (int c = 0; c < 8; ++c) {
					const float s = warp_sum(valid4 ? v[c] : 0.f);
					if (lane == 0) part[warp * 8 + c] = s;
				}
				row_bar();
				if (tid < 8) {
					float s = 0.f;
#pragma unroll 8
					for (int c = 0; c < 32; ++c) {
						const float* pp = part + (c >> 3) * 32 + (c & 7);  // channel c = 8 g' + i lives in warps 4g' .. 4g'+3
						const float m = ((pp[0] + pp[8]) + (pp[16] + pp[24])) * (1.f / 64.f);
						s = fmaf(sp_c[par::fc0 + tid * 32 + c], m, s);
					}
					hid[tid] = relu_f(s);
				}
				row_bar();
				if (tid < 32) {
					float s = 0.f;
#pragma unroll
					for (int j = 0; j < 8; ++j) s = fmaf(sp_c[par::fc2 + tid * 8 + j], hid[j], s);
					scale[tid] = sigmoid_f(s);
				}
				row_bar();
				if (valid4) {
#pragma unroll
					for (int c = 0; c < 8; ++c) v[c] *= scale[g * 8 + c];
					if (tap_stage == 4) {
#pragma unroll
						for (int c = 0; c < 8; ++c) tap_out[(leaf * 32 + g * 8 + c) * 64 + p4] = v[c];
					}
					// the VQ GEMM's A operand (dense rows, stored twice) ...
					store_split8(xh + (uint32_t)g * kXhPlane + (uint32_t)p4 * 16, kXhPrec, v);
					store_split8(xh + (uint32_t)g * kXhPlane + (uint32_t)(64 + p4) * 16, kXhPrec, v);
					// ... an fp32 copy of the row (z = W x + b of the near-tie rows) and this thread's share of |x|^2 (shortlist bound)
					*reinterpret_cast<float4*>(xs + p4 * kXsPitch + g * 8) = make_float4(v[0], v[1], v[2], v[3]);
					*reinterpret_cast<float4*>(xs + p4 * kXsPitch + g * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
					float xxp = 0.f;
#pragma unroll
					for (int c = 0; c < 8; ++c) xxp = fmaf(v[c], v[c], xxp);
					xs[p4 * kXsPitch + 32 + g] = xxp;
				}
			}
			signal_a_ready(bars, lane);
			lap(12);

			// ---- VQ: argmin_k (|z|^2 + |e_k|^2) - 2 z.e_k (save_for_inference.py:55-61), first minimum wins ; z = W x + b ----
			//  1. a_k = (|e_k|^2 - 2 b.e_k) - 2 (x_hi.M_hi + (x_hi.M_lo + x_lo.M_hi) / 2048) from the tensor cores, M = proj.weight^T e
			//     folded on the host in double.  Each operand keeps 22 significant bits, so |a_k - (exact score - |z|^2)| <=
			//     2 * (3 * 2^-22 [split + dropped lo.lo term] + 2 * 2^-22 [accumulator truncation, 2 steps]) * sum|x_c M_kc|
			//     <= 2.4e-6 |x| |M_k|; used bound B_k = 4e-6 |x| |M_k| + 1e-4, the 1e-4 covering the fp32 evaluation noise of the
			//     reference's own z and distance formula.
			//  2. one pass over the scores keeps, per thread, the two smallest a_k and the code of the smallest.  With
			//     Bmax = max_k B_k, every code that can be the fp32 arg-min has a_k <= min_j a_j + 2 Bmax; a row whose
			//     second-smallest score exceeds that has exactly one such code: done (~99 % of the rows).
			//  3. otherwise (near-tie): z = W x + b in fp32 for that row, every code with a_k <= min_j a_j + 2 Bmax that is also
			//     within 2 Bmax of its own thread's minimum (a superset of {a_k - B_k <= min_j (a_j + B_j)}) is re-scored with
			//     the reference's fp32 formula, sequential in d; the fp32 arg-min and all its ties are in that shortlist.  The
			//     step is entered by the whole CTA when any of its 64 rows needs it.
			//  GEMM row r = latent position r & 63; the eight threads of a position (4 channel groups x 2 row copies) take 32
			//  codes each, ascending with o8 = (r >> 6) * 4 + g.  tcgen05.ld is warp-collective, so every lane runs the loads.
			const int vp = row & 63, o8 = (row >> 6) * 4 + g, kb = o8 * 32;
			wait_accumulator(rc);
			row_bar();  // the |x|^2 partials and fp32 x rows of all four groups are in place
			lap(15);
			{
				const float* xp = xs + vp * kXsPitch + 32;
				const float xx = (xp[0] + xp[1]) + (xp[2] + xp[3]);
				const float cb = 4e-6f * sqrtf(xx);
				const float bmax = fmaf(cb, s_mno[256], 1e-4f);
				float a1 = INFINITY, a2 = INFINITY;
				int k1 = 0;
				uint32_t near_mask = 0u;  // this thread's codes within 2 Bmax of its smallest score
				{
					float hh[2][16], mx[2][16];
					tmem_ld16_nowait(rc.tlane + kb, hh[0]);
					tmem_ld16_nowait(rc.tlane + 256 + kb, mx[0]);
					tmem_ld16_nowait(rc.tlane + kb + 16, hh[1]);
					tmem_ld16_nowait(rc.tlane + 256 + kb + 16, mx[1]);
					tmem_wait_ld();
					// The scores are in registers and nothing below reads TMEM again: the accumulators are drained, so the
					// next leaf's conv1 — its input has been in A8 since the res32 phase — starts under the rest of the VQ.
					if (has_next) signal_a_ready(bars, lane);
#pragma unroll
					for (int j = 0; j < 32; ++j) {
						const int k = kb + j;
						const float a = s_esq2[k] - 2.f * fmaf(mx[j >> 4][j & 15], kLoInv, hh[j >> 4][j & 15]);
						hh[j >> 4][j & 15] = a;
						k1 = a < a1 ? k : k1;  // codes ascend: strict < keeps the lowest code among equal scores
						a2 = fminf(a2, fmaxf(a1, a));
						a1 = fminf(a1, a);
					}
					const float near = a1 + 2.f * bmax;
#pragma unroll
					for (int j = 0; j < 32; ++j)
						if (!(hh[j >> 4][j & 15] > near)) near_mask |= 1u << j;
				}
				vq_lo1[o8 * 64 + vp] = a1;
				vq_lo2[o8 * 64 + vp] = a2;
				vq_k1[o8 * 64 + vp] = k1;
				row_bar();
				lap(16);
				float l1 = INFINITY, l2 = INFINITY;
				int kbest = 0;
#pragma unroll
				for (int o = 0; o < 8; ++o) {  // ascending code ranges
					const float b1 = vq_lo1[o * 64 + vp], b2 = vq_lo2[o * 64 + vp];
					if (b1 < l1) {
						l2 = l1;
						l1 = b1;
						kbest = vq_k1[o * 64 + vp];
					} else {
						l2 = fminf(l2, b1);
					}
					l2 = fminf(l2, b2);
				}
				const float umin = l1 + bmax;  // >= min_j (a_j + B_j)
				// a non-finite row (inf / nan inputs; fminf would skip a nan score) takes the exact path
				const bool amb = !(l2 > umin + bmax) || !(xx < INFINITY) || tap_stage == 5;
				uint32_t any_amb;
				asm volatile(
				    "{\n.reg .pred p, q;\nsetp.ne.u32 q, %1, 0;\nbar.red.or.pred p, 1, 512, q;\nselp.u32 %0, 1, 0, p;\n}\n"
				    : "=r"(any_amb)
				    : "r"((uint32_t)amb)
				    : "memory");  // also: every thread is done reading the exchange arrays
				if (!any_amb) {
					if (o8 == 0) indices[leaf * 64 + vp] = (uint8_t)kbest;  // vp = (d*4+h)*4+w == view(B,4,4,4)
					lap(17);
				} else {
					// z rows of the near-tie positions (the warps of row copy 0): lane l computes dim 32g + l, sequential in c
					// from the bias
					if (row < 64) {
						for (uint32_t m = __ballot_sync(0xffffffffu, amb); m; m &= m - 1) {
							const int r = __ffs((int)m) - 1;
							const int rp = __shfl_sync(0xffffffffu, vp, r);
							const float* xr = xs + rp * kXsPitch;
							float acc = sp_c[par::proj_b + g * 32 + lane];
#pragma unroll 8
							for (int c = 0; c < 32; ++c) acc = fmaf(xr[c], s_wp[c * 128 + g * 32 + lane], acc);
							zs[rp * kZsPitch + g * 32 + lane] = acc;
						}
					}
					row_bar();
					if (amb && row < 64) {  // |z|^2 as four per-group partial sums, each sequential in d
						float zzp = 0.f;
#pragma unroll 8
						for (int d = 0; d < 32; ++d) {
							const float zv = zs[vp * kZsPitch + g * 32 + d];
							zzp = fmaf(zv, zv, zzp);
							if (tap_stage == 5) tap_out[(leaf * 128 + g * 32 + d) * 64 + vp] = zv;
						}
						zs[vp * kZsPitch + 128 + g] = zzp;
					}
					// Shortlist of a near-tie row: every code that can be the fp32 arg-min has a_k <= min_j a_j + 2 Bmax, hence lies
					// within 2 Bmax of the smallest score of ITS thread, whose own minimum must lie in the row's window too —
					// a superset of {a_k - B_k <= min_j (a_j + B_j)} that needs no second look at the (released) accumulators.
					uint32_t mask = !(a1 > umin + bmax) ? near_mask : 0u;
					if (!amb) mask = 0u;
					row_bar();
					const float* zrow = zs + vp * kZsPitch;
					const float zz = (zrow[128] + zrow[129]) + (zrow[130] + zrow[131]);
					float best = INFINITY;
					int bi = 0x7fffffff;
					while (mask) {  // two candidates per trip: two independent FMA chains
						const int b0 = __ffs((int)mask) - 1;
						mask &= mask - 1;
						const bool two = mask != 0u;
						const int b1 = two ? __ffs((int)mask) - 1 : b0;
						mask &= mask - 1;  // no-op on zero
						const int code0 = kb + b0, code1 = kb + b1;
						const float4* e0 = reinterpret_cast<const float4*>(w.emb + code0 * 128);
						const float4* e1 = reinterpret_cast<const float4*>(w.emb + code1 * 128);
						float dot0 = 0.f, dot1 = 0.f;
#pragma unroll 4
						for (int q = 0; q < 32; ++q) {
							const float4 ea = __ldg(e0 + q), eb = __ldg(e1 + q);
							const float4 zv = *reinterpret_cast<const float4*>(zrow + q * 4);
							dot0 = fmaf(zv.x, ea.x, dot0); dot1 = fmaf(zv.x, eb.x, dot1);
							dot0 = fmaf(zv.y, ea.y, dot0); dot1 = fmaf(zv.y, eb.y, dot1);
							dot0 = fmaf(zv.z, ea.z, dot0); dot1 = fmaf(zv.z, eb.z, dot1);
							dot0 = fmaf(zv.w, ea.w, dot0); dot1 = fmaf(zv.w, eb.w, dot1);
						}
						const float dist0 = (zz + s_esq[code0]) - 2.f * dot0, dist1 = (zz + s_esq[code1]) - 2.f * dot1;
						if (dist0 < best) {  // codes ascend, so strict < keeps the first minimum
							best = dist0;
							bi = code0;
						}
						if (two && dist1 < best) {
							best = dist1;
							bi = code1;
						}
					}
					lap(17);
					vq_umin[o8 * 64 + vp] = best;
					vq_k1[o8 * 64 + vp] = bi;
					row_bar();
					if (o8 == 0) {
						if (amb) {
#pragma unroll
							for (int o = 1; o < 8; ++o) {  // ascending code ranges: strict < keeps the lowest code among equal distances
								const float ob = vq_umin[o * 64 + vp];
								if (ob < best) {
									best = ob;
									bi = vq_k1[o * 64 + vp];
								}
							}
							if (bi == 0x7fffffff) bi = kbest;  // every distance was nan: keep the shortlist's choice
						} else {
							bi = kbest;
						}
						indices[leaf * 64 + vp] = (uint8_t)bi;
					}
				}
			}
			// ---- leaf done: Y (dirtied by the VQ overlays) is cleared for the next conv2 epilogue, and the next leaf's conv1 —
			//      its input has been in A8 since the res32 phase — may start now that the accumulators are drained ----
			if (has_next) {
				row_bar();  // every row thread is done with the VQ overlays
				// conv1 has been running since the scores left TMEM; the clear runs under its MMAs, and the GroupNorm barriers
				// of the conv1 epilogue order it before the conv2 epilogue's stores into Y
				for (uint32_t i = tid; i < (kOffXh - kOffY + kXhBytes) / 16; i += kRowThreads) reinterpret_cast<uint4*>(smem + kOffY)[i] = make_uint4(0, 0, 0, 0);
			}
			lap(18);
		}
		if (kProf && tap_out && tid == 0) {
			float* o = tap_out + (size_t)blockIdx.x * 64;
#pragma unroll
			for (int i = 0; i < 20; ++i) o[i] = (float)prof[i];
		}
	}

	// ---- teardown ----
	tc_fence_before();
	__syncthreads();
	if (warp == kIssuerWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

}  // namespace

cudaError_t configure_encode_tc() {
	cudaError_t e = cudaFuncSetAttribute(encode_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
	if (e != cudaSuccess) return e;
	return cudaFuncSetAttribute(encode_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
}

cudaError_t launch_encode_tc(const EncoderWeights& w, const EncoderTcStream& ws, const float* dev_leaves, int64_t n_leaves,
                             uint8_t* dev_indices, int num_sms, cudaStream_t stream, int tap_stage, float* tap_out) {
	if (n_leaves <= 0) return cudaSuccess;
	const int grid = (int)(n_leaves < (int64_t)num_sms ? n_leaves : (int64_t)num_sms);
	if (tap_stage == 100)
		encode_tc_kernel<true><<<grid, kThreads, kSmemBytes, stream>>>(w, ws, dev_leaves, n_leaves, dev_indices, -1, tap_out);
	else
		encode_tc_kernel<false><<<grid, kThreads, kSmemBytes, stream>>>(w, ws, dev_leaves, n_leaves, dev_indices, tap_stage, tap_out);
	return cudaGetLastError();
}

}  // namespace vqvdb
