// Tensor-core decoder (tcgen05.mma, TMEM accumulators, bf16 operands, fp32 accumulation): 64 uint8 indices in, 512 voxels
// out, one kernel.  Reference: python/save_for_inference.py:91-104 (gather + permute) + python/VQVAE_v2.py:253-275
// (DecoderFloat.forward); weight stream and the folded tail are described in decode_tc.cuh.
//
//   * the gathered A operand goes through TMEM (TS-mode MMA): each of a tile's 128 rows (the 64 latent positions of two
//     leaves) is staged by its own threads — the tap-shifted neighbour's channels, or a zero row outside the leaf —
//     from the channels-last bf16 activation buffer into its TMEM lane (tcgen05.st).
//   * a unit is one (kd, kh) tap PAIR: the A rows are shifted along d and h only, and the three kw taps ride along N —
//     B = [3 kw x 64 cout][64 cin] = three consecutive 8 KB tiles of the weight stream, N = 192 (96 cycles per MMA, the
//     pipe's floor).  The accumulator holds three partial convolutions per row; the kw shift is applied when it is read:
//     out(w) = P0(w-1) + P1(w) + P2(w+1), a lane shuffle that never leaves the warp (w = lane & 3) and supplies the zero
//     padding along w.  (First generation, one unit per tap and N = 64: 5.5 M leaves/s, tensor pipe 28 % busy,
//     profiles/r1d_decode_tc_pipeline.txt; this scheme: 9.0 M before the fold below.)
//   * the accumulator is 192 columns, so a CTA holds 2 tiles (2 x 2 leaves): TMEM = 2 x 192 (D) + 2 x 2 x 32 (A,
//     double-buffered).  Each 128-row tile is served by EIGHT warps: warp = (TMEM lane quadrant, channel half) — a thread
//     owns 32 of the 64 channels of its row for staging and for every epilogue, GroupNorm groups never straddle the
//     halves, and the two halves of a leaf only meet for the channel attention and the final store.
//   * up_conv -> PixelShuffle3D -> final is one linear map, folded on the host into a 64 -> 64 convolution G + an 8-term
//     gather per output voxel (decode_tc.cuh): 45 units per group of leaves and no CUDA-core convolution at all.  G is
//     parked as fp32 [64 ch][64 pos] (channel-major, skewed so the gather is bank-conflict free) over the leaf's by
//     then idle activation region; thread (row R, channel half) sums the in-grid neighbours for its four voxels,
//     applies the sigmoid and stores 16 bytes.
// Measured (1 M leaves, B200): 15.4 M leaves/s.
#include <cuda_bf16.h>

#include "decode_tc.cuh"
#include "leaf_ops.cuh"
#include "ptx_utils.cuh"

namespace vqvdb {

namespace {

constexpr int kTiles = 2;
constexpr int kLeavesPerCta = 2 * kTiles;                  // 4
constexpr int kWorkWarps = 8 * kTiles;                     // 16: (tile, channel half, lane quadrant)
constexpr int kCtrlWarps = kTiles + 1;                     // one MMA issuer per tile + the TMA producer
constexpr int kThreads = (kWorkWarps + kCtrlWarps) * 32;   // 608
constexpr int kStages = 4;
constexpr uint32_t kSrcUnitBytes = 8192;                   // one [64 n][64 k] tile of the weight stream
constexpr uint32_t kUnitBytes = 3 * kSrcUnitBytes;         // [3 kw x 64 n][64 k]
// units per group of leaves: stem (2 input halves), res conv1, res conv2, the folded tail conv (decode_tc.cuh:
// up_conv -> PixelShuffle3D -> final as one 64 -> 64 conv + an 8-term gather)
constexpr int kUnitsPerGroup = 18 + 9 + 9 + 9;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kDCols = 192;                           // tile t accumulator: columns [t*192, t*192 + 192)
constexpr uint32_t kColA = kTiles * kDCols;                // tile t A buffers: columns 384 + t*64 + buf*32

// instruction descriptor: D = f32, A = B = bf16, both K-major, N = 192, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((192u >> 3) << 17) | ((128u >> 4) << 24);

// per-channel parameters staged in shared memory (float offsets)
namespace par {
constexpr int stem_b = 0, stem_gn_w = 64, stem_gn_b = 128, gn1_w = 192, gn1_b = 256, c1_b = 320, gn2_w = 384, gn2_b = 448,
              c2_b = 512, fin_b = 576, fc0 = 580, fc2 = fc0 + 16 * 72, fold_b = fc2 + 64 * 17, total = fold_b + 64;
constexpr int fc0_pitch = 72, fc2_pitch = 17;
}

// shared memory map (bytes)
constexpr uint32_t kOffRing = 0;
constexpr uint32_t kOffLeaf = kOffRing + kStages * kUnitBytes;            // 4 x 16 KB
constexpr uint32_t kLeafBytes = 16384 + 16;                               // + the 3-word skew of the fp32 G planes (see the folded tail)
constexpr uint32_t kOffZero = kOffLeaf + kLeavesPerCta * kLeafBytes;      // 128 zero bytes: the source row of out-of-leaf taps
constexpr uint32_t kOffBar = kOffZero + 128;
constexpr uint32_t kNumBars = 2 * kStages + 4 * kTiles + kTiles;          // w_full, w_empty, a_full[t][2], a_empty[t][2], d_full[t]
constexpr uint32_t kOffTmemSlot = kOffBar + kNumBars * 8;
constexpr uint32_t kOffPar = kOffTmemSlot + 16;
constexpr uint32_t kOffScratch = kOffPar + par::total * 4;
// per-leaf scratch (floats): exch [2 slots][2 warps][2 halves][8], part [2][64], scale [64], hid [16], idx [16 words]
constexpr uint32_t kScrExch = 0, kScrPart = 64, kScrScale = 192, kScrHid = 256, kScrIdx = 272, kScratchFloats = 288;
constexpr uint32_t kSmemBytes = kOffScratch + kLeavesPerCta * kScratchFloats * 4;
static_assert(kSmemBytes <= 227 * 1024, "decode_tc smem budget");
static_assert(kOffBar % 8 == 0 && kOffPar % 16 == 0 && kOffScratch % 16 == 0, "alignment");

__device__ __forceinline__ uint32_t bar_w_full(uint32_t bars, uint32_t s) { return bars + s * 8; }
__device__ __forceinline__ uint32_t bar_w_empty(uint32_t bars, uint32_t s) { return bars + (kStages + s) * 8; }
__device__ __forceinline__ uint32_t bar_a_full(uint32_t bars, uint32_t t, uint32_t b) { return bars + (2 * kStages + t * 2 + b) * 8; }
__device__ __forceinline__ uint32_t bar_a_empty(uint32_t bars, uint32_t t, uint32_t b) { return bars + (2 * kStages + 2 * kTiles + t * 2 + b) * 8; }
__device__ __forceinline__ uint32_t bar_d_full(uint32_t bars, uint32_t t) { return bars + (2 * kStages + 4 * kTiles + t) * 8; }

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem_d] (+)= A[tmem_a] (128 x 16 bf16, TMEM) * B[desc] (192 x 16 bf16, shared)^T
__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t accumulate) {
	asm volatile(
	    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
	    "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
	    "r"(tmem_a), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
	    : "memory");
}
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
	// K-major SWIZZLE_128B: 128-byte rows, 8-row groups 1024 B apart
	return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
	       ((uint64_t)2 << 61);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
	asm volatile(
	    "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
	    "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
	    "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
	    : "memory");
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
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
	asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
	uint4 v;
	asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
	return v;
}
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
	asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

// Per-thread view of the work: which tile / TMEM lane quadrant / channel half / leaf / latent position this thread is.
struct Worker {
	int tile, quad, chalf, lane, row;  // row = quad*32 + lane in [0,128)
	int leaf_slot;                     // tile*2 + (row >> 6)
	int pos, d, h, w;                  // latent position within the leaf
	int wil;                           // warp-in-leaf along the rows: 0 or 1
	uint32_t unit = 0;                 // units staged so far (same sequence in every worker of the tile and in its issuer)
	uint32_t passes = 0;               // accumulator hand-overs so far (parity of d_full)
	uint32_t reds = 0;                 // half-leaf reductions so far (alternates the exchange slot)
	uint32_t bars, tmem_lane, zero_row;
	long long t_wait = 0, t_stage = 0, t_acc = 0;  // kProf only
};
template <bool P> __device__ __forceinline__ long long prof_clock() { return P ? clock64() : 0; }

__device__ __forceinline__ void leaf_bar(const Worker& wk) { named_bar_sync(1 + wk.leaf_slot, 128); }
__device__ __forceinline__ void half_bar(const Worker& wk) { named_bar_sync(1 + kLeavesPerCta + wk.leaf_slot * 2 + wk.chalf, 64); }

// Stage one A unit: this thread's 32 channels (16 words) of the source row -> its TMEM lane, then hand the buffer to
// the MMA.  `seg` = shared address of the 128-byte (64-channel) segment of the source row, chunks swizzled by `swz`;
// out-of-leaf taps pass the zero row.  The loads are issued before the wait for the free buffer.
template <bool kProf>
__device__ __forceinline__ void stage_unit(Worker& wk, uint32_t seg, uint32_t swz) {
	const uint32_t buf = wk.unit & 1u;
	uint32_t r[16];
#pragma unroll
	for (int q = 0; q < 4; ++q) {
		const uint4 v = lds128(seg + ((((uint32_t)(wk.chalf * 4 + q)) ^ swz) << 4));
		r[4 * q] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
	}
	const long long c0 = prof_clock<kProf>();
	if (wk.lane == 0) mbar_wait(bar_a_empty(wk.bars, wk.tile, buf), ((wk.unit >> 1) & 1u) ^ 1u);
	__syncwarp();
	tc_fence_after();
	const long long c1 = prof_clock<kProf>();
	tmem_st16(wk.tmem_lane + kColA + wk.tile * 64 + buf * 32 + wk.chalf * 16, r);
	asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
	tc_fence_before();
	__syncwarp();
	if (wk.lane == 0) mbar_arrive(bar_a_full(wk.bars, wk.tile, buf));
	++wk.unit;
	if (kProf) {
		wk.t_wait += c1 - c0;
		wk.t_stage += prof_clock<kProf>() - c1;
	}
}

// The (kd, kh) = t tap pair of a conv whose input rows are HALVES * 128 bytes: HALVES units.
template <int HALVES, bool kProf>
__device__ __forceinline__ void stage_tap_pair(Worker& wk, uint32_t act_base, int t) {
	const int td = t / 3, th = t - td * 3;
	const bool ok = (unsigned)(wk.d + td - 1) < 4u && (unsigned)(wk.h + th - 1) < 4u;
	const int p2 = wk.pos + (td - 1) * 16 + (th - 1) * 4;
	const uint32_t row = ok ? act_base + (uint32_t)p2 * (HALVES * 128) : wk.zero_row;
	const uint32_t swz = ok ? ((uint32_t)p2 & 7u) : 0u;
#pragma unroll
	for (int half = 0; half < HALVES; ++half) stage_unit<kProf>(wk, row + (ok ? half * 128 : 0), swz);
}
template <int HALVES, bool kProf>
__device__ __forceinline__ void stage_conv(Worker& wk, uint32_t act_base) {
#pragma unroll 1
	for (int t = 0; t < 9; ++t) stage_tap_pair<HALVES, kProf>(wk, act_base, t);
}

// Wait until this tile's accumulator holds the finished layer.
template <bool kProf>
__device__ __forceinline__ void wait_accumulator(Worker& wk) {
	const long long c0 = prof_clock<kProf>();
	mbar_wait(bar_d_full(wk.bars, wk.tile), wk.passes & 1u);
	tc_fence_after();
	++wk.passes;
	if (kProf) wk.t_acc += prof_clock<kProf>() - c0;
}

// This thread's 32 output channels of the finished conv: the three kw partials combined across neighbouring rows.
__device__ __forceinline__ void load_conv32(const Worker& wk, float (&v)[32]) {
	const uint32_t base = wk.tmem_lane + wk.tile * kDCols + wk.chalf * 32;
	const bool has_lo = wk.w > 0, has_hi = wk.w < 3;
#pragma unroll
	for (int part = 0; part < 2; ++part) {
		float a[16], b[16], c[16];
		tmem_ld16_nowait(base + part * 16, a);        // kw = 0: belongs to the row at w + 1
		tmem_ld16_nowait(base + 64 + part * 16, b);   // kw = 1
		tmem_ld16_nowait(base + 128 + part * 16, c);  // kw = 2: belongs to the row at w - 1
		tmem_wait_ld();
#pragma unroll
		for (int j = 0; j < 16; ++j) {
			const float lo = __shfl_up_sync(0xffffffffu, a[j], 1);
			const float hi = __shfl_down_sync(0xffffffffu, c[j], 1);
			v[part * 16 + j] = b[j] + (has_lo ? lo : 0.f) + (has_hi ? hi : 0.f);
		}
	}
}

// Sum N per-thread values over the 64 rows of this thread's leaf, among the threads of its channel half (2 warps).
template <int N>
__device__ __forceinline__ void half_allreduce(float (&v)[N], Worker& wk, float* exch /* [2 slots][2 warps][2 halves][8] */) {
	static_assert(N <= 8, "exchange slot size");
#pragma unroll
	for (int i = 0; i < N; ++i) v[i] = warp_sum(v[i]);
	float* e = exch + (wk.reds & 1u) * 32;
	if (wk.lane == 0) {
#pragma unroll
		for (int i = 0; i < N; ++i) e[(wk.wil * 2 + wk.chalf) * 8 + i] = v[i];
	}
	half_bar(wk);
#pragma unroll
	for (int i = 0; i < N; ++i) v[i] = e[wk.chalf * 8 + i] + e[(2 + wk.chalf) * 8 + i];
	++wk.reds;  // the next reduction uses the other slot; this one is rewritten only after another barrier has been passed
}

// GroupNorm(8, 64) over the leaf for this thread's four groups: v -> statistics.
__device__ __forceinline__ void gn_stats(const float (&v)[32], Worker& wk, float* exch, float (&mean)[4], float (&rstd)[4]) {
	float st[8];
#pragma unroll
	for (int i = 0; i < 8; ++i) st[i] = 0.f;
#pragma unroll
	for (int j = 0; j < 32; ++j) {
		st[j >> 3] += v[j];
		st[4 + (j >> 3)] = fmaf(v[j], v[j], st[4 + (j >> 3)]);
	}
	half_allreduce<8>(st, wk, exch);
#pragma unroll
	for (int g = 0; g < 4; ++g) {
		mean[g] = st[g] * (1.f / 512.f);
		rstd[g] = 1.f / sqrtf(fmaxf(st[4 + g] * (1.f / 512.f) - mean[g] * mean[g], 0.f) + kGnEps);
	}
}

// row-major [64 pos][64 ch] bf16 buffer, 128-B rows, 16-B chunks swizzled by pos & 7: this thread's 32 channels
__device__ __forceinline__ void store_row_half(uint32_t buf_base, const Worker& wk, const float (&v)[32]) {
#pragma unroll
	for (int q = 0; q < 4; ++q) {
		const uint32_t a = buf_base + (uint32_t)wk.pos * 128 + ((((uint32_t)(wk.chalf * 4 + q)) ^ ((uint32_t)wk.pos & 7u)) << 4);
		sts128(a, pack_bf16(v[8 * q], v[8 * q + 1]), pack_bf16(v[8 * q + 2], v[8 * q + 3]), pack_bf16(v[8 * q + 4], v[8 * q + 5]),
		       pack_bf16(v[8 * q + 6], v[8 * q + 7]));
	}
}
__device__ __forceinline__ void load_row_chunk(uint32_t buf_base, const Worker& wk, int q, float (&x)[8]) {
	const uint4 raw = lds128(buf_base + (uint32_t)wk.pos * 128 + ((((uint32_t)(wk.chalf * 4 + q)) ^ ((uint32_t)wk.pos & 7u)) << 4));
	float2 f;
	f = unpack_bf16(raw.x); x[0] = f.x; x[1] = f.y;
	f = unpack_bf16(raw.y); x[2] = f.x; x[3] = f.y;
	f = unpack_bf16(raw.z); x[4] = f.x; x[5] = f.y;
	f = unpack_bf16(raw.w); x[6] = f.x; x[7] = f.y;
}

// Transposing butterfly: afterwards v[0] of lane L = sum over the warp's 32 lanes of the original v[L].  Destroys v.
__device__ __forceinline__ float column_sums(float (&v)[32], int lane) {
#pragma unroll
	for (int step = 16; step >= 1; step >>= 1) {
		const bool upper = (lane & step) != 0;
#pragma unroll
		for (int i = 0; i < step; ++i) {
			const float send = upper ? v[i] : v[i + step];
			const float keep = upper ? v[i + step] : v[i];
			v[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
		}
	}
	return v[0];
}

template <bool kProf>
__global__ void __launch_bounds__(kThreads, 1)
decode_tc_kernel(const DecoderMmaWeights w, const uint8_t* __restrict__ indices, int64_t n_leaves, float* __restrict__ voxels,
                 int active_tiles, int tap_stage, float* __restrict__ tap_out) {
	extern __shared__ __align__(1024) uint8_t smem[];
	const uint32_t s_base = smem_u32(smem);
	const uint32_t ring = s_base + kOffRing;
	const uint32_t bars = s_base + kOffBar;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	// active_tiles = 2: a group is 4 leaves (two 128-row tiles side by side).  Small calls (<= 2 leaves per SM) run with
	// ONE tile per CTA — twice the CTAs, each with half the work and the tensor pipe to itself; tile 1's warps idle.
	const int leaves_per_group = 2 * active_tiles;
	const int64_t n_groups = (n_leaves + leaves_per_group - 1) / leaves_per_group;
	uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffTmemSlot);
	float* s_par = reinterpret_cast<float*>(smem + kOffPar);

	// ---- per-channel parameters, final-conv and attention weights -> shared memory ----
	for (int i = threadIdx.x; i < 64; i += kThreads) {
		s_par[par::stem_b + i] = __ldg(w.stem_b + i);
		s_par[par::stem_gn_w + i] = __ldg(w.stem_gn_w + i);
		s_par[par::stem_gn_b + i] = __ldg(w.stem_gn_b + i);
		s_par[par::gn1_w + i] = __ldg(w.res.gn1_w + i);
		s_par[par::gn1_b + i] = __ldg(w.res.gn1_b + i);
		s_par[par::c1_b + i] = __ldg(w.res.c1_b + i);
		s_par[par::gn2_w + i] = __ldg(w.res.gn2_w + i);
		s_par[par::gn2_b + i] = __ldg(w.res.gn2_b + i);
		s_par[par::c2_b + i] = __ldg(w.res.c2_b + i);
		s_par[par::fold_b + i] = __ldg(w.fold_b + i);
	}
	for (int i = threadIdx.x; i < 1024; i += kThreads) {
		s_par[par::fc0 + (i >> 6) * par::fc0_pitch + (i & 63)] = __ldg(w.fc0 + i);  // [16][64]
		s_par[par::fc2 + (i >> 4) * par::fc2_pitch + (i & 15)] = __ldg(w.fc2 + i);  // [64][16]
	}
	if (threadIdx.x == 0) s_par[par::fin_b] = __ldg(w.fin_b);
	if (threadIdx.x < 32) reinterpret_cast<uint32_t*>(smem + kOffZero)[threadIdx.x] = 0u;
	if (threadIdx.x == 0) {
		for (uint32_t s = 0; s < kStages; ++s) {
			mbar_init(bar_w_full(bars, s), 1);
			mbar_init(bar_w_empty(bars, s), active_tiles);  // one tcgen05.commit per active tile issuer
		}
		for (uint32_t t = 0; t < kTiles; ++t) {
			for (uint32_t b = 0; b < 2; ++b) {
				mbar_init(bar_a_full(bars, t, b), 8);   // one arrival per worker warp of the tile
				mbar_init(bar_a_empty(bars, t, b), 1);  // tcgen05.commit
			}
			mbar_init(bar_d_full(bars, t), 1);
		}
		mbar_fence_init();
	}
	if (warp == kWorkWarps) {  // the first control warp owns the TMEM allocation
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
	}
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem = *tmem_slot;
	const int64_t my_groups = blockIdx.x < n_groups ? (n_groups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

	if (warp >= kWorkWarps) {
		// ===================== control warps =====================
		if (warp < kWorkWarps + kTiles) {
#ifdef VQVDB_DEC_ISSUER_LANE0
			const bool leader = true;
			if (lane == 0 && warp - kWorkWarps < active_tiles) {
#else
			// The whole warp runs the issue loop, one elected lane executes the tcgen05 instructions: warp-uniform control
			// flow keeps descriptors and loop state in uniform registers (see encode_tc.cu's issuer for the measurement).
			const bool leader = elect_one();
			if (warp - kWorkWarps < active_tiles) {
#endif
				// MMA issuer of tile t.  The issuers are independent, so the tiles drift apart and one tile's epilogue
				// overlaps the other tile's MMAs.
				const uint32_t t = warp - kWorkWarps;
				uint32_t unit = 0;
				long long tw = 0, ta = 0, ti = 0;
				const long long tstart = prof_clock<kProf>();
				for (int64_t g = 0; g < my_groups; ++g) {
#pragma unroll 1
					for (int u = 0; u < kUnitsPerGroup; ++u) {
						const uint32_t s = unit % kStages, buf = unit & 1u;
						// position of this unit inside its layer pass: stem = 18 units, then six passes of 9
						const int in_pass = u < 18 ? u : (u - 18) % 9;
						const bool last = u < 18 ? (u == 17) : (in_pass == 8);
						const long long c0 = prof_clock<kProf>();
						mbar_wait(bar_w_full(bars, s), (unit / kStages) & 1u);
						const long long c1 = prof_clock<kProf>();
						mbar_wait(bar_a_full(bars, t, buf), (unit >> 1) & 1u);
						tc_fence_after();
						const long long c2 = prof_clock<kProf>();
						const uint64_t bdesc = make_desc_sw128(ring + s * kUnitBytes);
#pragma unroll
						for (uint32_t kk = 0; kk < 4; ++kk)
							if (leader) tc_mma_ts(tmem + t * kDCols, tmem + kColA + t * 64 + buf * 32 + kk * 8, bdesc + (uint64_t)(kk * 2),
							          (in_pass > 0 || kk > 0) ? 1u : 0u);
						if (leader) tc_commit(bar_a_empty(bars, t, buf));
						if (last && leader) tc_commit(bar_d_full(bars, t));
						if (leader) tc_commit(bar_w_empty(bars, s));
						++unit;
						if (kProf) {
							tw += c1 - c0;
							ta += c2 - c1;
							ti += prof_clock<kProf>() - c2;
						}
					}
				}
				if (kProf && tap_out && leader) {
					float* o = tap_out + ((size_t)blockIdx.x * kThreads + threadIdx.x) * 4;
					o[0] = (float)tw; o[1] = (float)ta; o[2] = (float)ti; o[3] = (float)(prof_clock<kProf>() - tstart);
				}
			}
		} else if (lane == 0) {
			// TMA producer: one 24 KB unit per (kd, kh) pair = three consecutive 8 KB tiles of the stream (kw = 0, 1, 2);
			// the stem's stream interleaves its two input-channel halves, so its units are gathered by three copies.
			const uint32_t total = (uint32_t)(my_groups * kUnitsPerGroup);
#pragma unroll 1
			for (uint32_t issued = 0; issued < total; ++issued) {
				const uint32_t s = issued % kStages, u = issued % kUnitsPerGroup;
				mbar_wait(bar_w_empty(bars, s), ((issued / kStages) & 1u) ^ 1u);
				mbar_arrive_expect_tx(bar_w_full(bars, s), kUnitBytes);
				const uint32_t dst = ring + s * kUnitBytes;
				if (u < 18) {
					const uint32_t pair = u >> 1, half = u & 1u;
#pragma unroll
					for (uint32_t kw = 0; kw < 3; ++kw)
						tma_load_1d(dst + kw * kSrcUnitBytes, w.units + (size_t)((pair * 3 + kw) * 2 + half) * kSrcUnitBytes, kSrcUnitBytes,
						            bar_w_full(bars, s));
				} else {
					const uint32_t src = u >= 36 ? (uint32_t)kDecUnitsTotal + (u - 36) * 3 : 54 + (u - 18) * 3;
					tma_load_1d(dst, w.units + (size_t)src * kSrcUnitBytes, kUnitBytes, bar_w_full(bars, s));
				}
			}
		}
		__syncwarp();
	} else {
		// ===================== worker warps: A staging + epilogues =====================
		Worker wk;
		wk.tile = warp >> 3;
		wk.chalf = (warp >> 2) & 1;
		wk.quad = warp & 3;
		wk.lane = lane;
		wk.row = wk.quad * 32 + lane;
		wk.leaf_slot = wk.tile * 2 + (wk.row >> 6);
		wk.wil = (wk.row >> 5) & 1;
		wk.pos = wk.row & 63;
		wk.d = wk.pos >> 4;
		wk.h = (wk.pos >> 2) & 3;
		wk.w = wk.pos & 3;
		wk.bars = bars;
		wk.tmem_lane = tmem + ((uint32_t)(wk.quad * 32) << 16);
		wk.zero_row = s_base + kOffZero;
		uint8_t* region = smem + kOffLeaf + wk.leaf_slot * kLeafBytes;
		const uint32_t a_base = s_base + kOffLeaf + wk.leaf_slot * kLeafBytes;  // Q [64][128] for the stem, then A [64][64]
		const uint32_t x_base = a_base + 8192;                                   // residual x [64][64] bf16; later the pixel-shuffled planes
		float* scratch = reinterpret_cast<float*>(smem + kOffScratch) + wk.leaf_slot * kScratchFloats;
		float* exch = scratch + kScrExch;
		float* s_part = scratch + kScrPart;
		float* s_scale = scratch + kScrScale;
		float* s_hid = scratch + kScrHid;
		uint32_t* s_idx = reinterpret_cast<uint32_t*>(scratch + kScrIdx);
		const int tl = wk.chalf * 64 + wk.pos;  // thread index within the leaf, 0..127
		const int c0 = wk.chalf * 32;           // first of this thread's 32 channels
		long long ep[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // kProf only: cycles per section of this thread
		long long ep0 = prof_clock<kProf>();
		auto lap = [&](int slot) {
			if (kProf) {
				const long long c = prof_clock<kProf>();
				ep[slot] += c - ep0;
				ep0 = c;
			}
		};

		// The leaf's 64 indices (16 words) come from HBM, ~2 k cycles away: the words of the NEXT group's leaf are requested
		// at the top of a group and have long arrived when that group's gather needs them.
		const int64_t my_tile_groups = wk.tile < active_tiles ? my_groups : 0;
		auto load_index_word = [&](int64_t g) -> uint32_t {
			const int64_t lf = (blockIdx.x + g * gridDim.x) * leaves_per_group + wk.leaf_slot;
			return (tl < 16 && g < my_tile_groups && lf < n_leaves) ? __ldcs(reinterpret_cast<const uint32_t*>(indices + lf * 64) + tl) : 0u;
		};
		uint32_t idx_word = load_index_word(0);
		for (int64_t g = 0; g < my_tile_groups; ++g) {
			const int64_t grp = blockIdx.x + g * gridDim.x;
			const int64_t leaf = grp * leaves_per_group + wk.leaf_slot;
			const bool leaf_ok = leaf < n_leaves;

			lap(7);
			// ---- gather: Q[pos][0..127] = codebook_bf16[idx[pos]] (spare slots decode code 0) ----
			if (tl < 16) s_idx[tl] = idx_word;
			idx_word = load_index_word(g + 1);
			leaf_bar(wk);  // also: every thread of the leaf is done with the previous group's planes
			{
				const uint8_t* idx8 = reinterpret_cast<const uint8_t*>(s_idx);
#pragma unroll
				for (int i = tl; i < 64 * 16; i += 128) {  // eight L2 loads per thread, all in flight
					const int pos = i >> 4, c = i & 15;
					const uint4 v = __ldg(reinterpret_cast<const uint4*>(w.emb_bf16 + (size_t)idx8[pos] * 128) + c);
					const uint32_t pc = (c & 8) | ((c & 7) ^ (pos & 7));
					*reinterpret_cast<uint4*>(region + pos * 256 + pc * 16) = v;
				}
			}
			leaf_bar(wk);
			lap(0);

			float v[32];
			float mean[4], rstd[4];
			// ---- stem.0 (128->64) ; stem.1 GroupNorm + ReLU -> x ; gn1 + ReLU -> conv1 input ----
			stage_conv<2, kProf>(wk, a_base);
			wait_accumulator<kProf>(wk);  // all stem MMAs done => every read of Q is done too
			lap(6);
			load_conv32(wk, v);
#pragma unroll
			for (int j = 0; j < 32; ++j) v[j] += s_par[par::stem_b + c0 + j];
			gn_stats(v, wk, exch, mean, rstd);
#pragma unroll
			for (int j = 0; j < 32; ++j)
				v[j] = fmaxf((v[j] - mean[j >> 3]) * rstd[j >> 3] * s_par[par::stem_gn_w + c0 + j] + s_par[par::stem_gn_b + c0 + j], 0.f);
			if (tap_stage == 0 && leaf_ok) {
#pragma unroll
				for (int j = 0; j < 32; ++j) tap_out[leaf * 4096 + (c0 + j) * 64 + wk.pos] = v[j];
			}
			store_row_half(x_base, wk, v);  // residual x (thread-private half row)
			gn_stats(v, wk, exch, mean, rstd);
#pragma unroll
			for (int j = 0; j < 32; ++j)
				v[j] = fmaxf((v[j] - mean[j >> 3]) * rstd[j >> 3] * s_par[par::gn1_w + c0 + j] + s_par[par::gn1_b + c0 + j], 0.f);
			store_row_half(a_base, wk, v);
			half_bar(wk);  // the rows of this channel half are in place (staging reads its own half only)
			lap(1);

			// ---- res conv1 ; gn2 + ReLU -> conv2 input ----
			stage_conv<1, kProf>(wk, a_base);
			wait_accumulator<kProf>(wk);
			lap(6);
			load_conv32(wk, v);
#pragma unroll
			for (int j = 0; j < 32; ++j) v[j] += s_par[par::c1_b + c0 + j];
			gn_stats(v, wk, exch, mean, rstd);
#pragma unroll
			for (int j = 0; j < 32; ++j)
				v[j] = fmaxf((v[j] - mean[j >> 3]) * rstd[j >> 3] * s_par[par::gn2_w + c0 + j] + s_par[par::gn2_b + c0 + j], 0.f);
			store_row_half(a_base, wk, v);
			half_bar(wk);
			lap(2);

			// ---- res conv2 ; x + 0.1 * (.) ; ChannelAttention(64) -> up_conv input ----
			stage_conv<1, kProf>(wk, a_base);
			wait_accumulator<kProf>(wk);
			lap(6);
			load_conv32(wk, v);
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				float xr[8];
				load_row_chunk(x_base, wk, q, xr);
#pragma unroll
				for (int i = 0; i < 8; ++i) v[8 * q + i] = xr[i] + kResScale * (v[8 * q + i] + s_par[par::c2_b + c0 + 8 * q + i]);
			}
			if (tap_stage == 1 && leaf_ok) {
#pragma unroll
				for (int j = 0; j < 32; ++j) tap_out[leaf * 4096 + (c0 + j) * 64 + wk.pos] = v[j];
			}
			store_row_half(x_base, wk, v);  // x' (thread-private), re-read below once the channel scales are known
			{
				const float cs = column_sums(v, lane);  // channel c0 + lane over this warp's 32 rows
				s_part[wk.wil * 64 + c0 + lane] = cs;
			}
			leaf_bar(wk);
			{
				// hidden = relu(fc0 [16][64] . mean): 8 threads per hidden unit, 8 channels each
				const int unit = tl >> 3, part = tl & 7;
				float s = 0.f;
#pragma unroll
				for (int i = 0; i < 8; ++i) {
					const int c = i * 8 + part;
					s = fmaf(s_par[par::fc0 + unit * par::fc0_pitch + c], (s_part[c] + s_part[64 + c]) * (1.f / 64.f), s);
				}
				s += __shfl_xor_sync(0xffffffffu, s, 1);
				s += __shfl_xor_sync(0xffffffffu, s, 2);
				s += __shfl_xor_sync(0xffffffffu, s, 4);
				if (part == 0) s_hid[unit] = fmaxf(s, 0.f);
			}
			leaf_bar(wk);
			if (tl < 64) {
				float s = 0.f;
#pragma unroll
				for (int j = 0; j < 16; ++j) s = fmaf(s_par[par::fc2 + tl * par::fc2_pitch + j], s_hid[j], s);
				s_scale[tl] = sigmoid_f(s);
			}
			leaf_bar(wk);
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				float xr[8];
				load_row_chunk(x_base, wk, q, xr);
#pragma unroll
				for (int i = 0; i < 8; ++i) v[8 * q + i] = xr[i] * s_scale[c0 + 8 * q + i];
			}
			if (tap_stage == 2 && leaf_ok) {
#pragma unroll
				for (int j = 0; j < 32; ++j) tap_out[leaf * 4096 + (c0 + j) * 64 + wk.pos] = v[j];
			}
			store_row_half(a_base, wk, v);
			half_bar(wk);
			lap(3);

			// ---- folded tail: G = conv(a; Wg) + bg, then out[2p + r] = sigmoid(fin_b + sum over the in-grid cells
			// p + e(r, eps) of G[p + e][r*8 + eps]) ----
			stage_conv<1, kProf>(wk, a_base);
			wait_accumulator<kProf>(wk);  // all reads of the conv input are done: the whole leaf region is free
			lap(6);
			load_conv32(wk, v);
			// The accumulator hand-over already orders every staging read of this leaf's rows before the stores below (a
			// unit's loads feed the tcgen05.st that precedes its a_full arrival); the barrier states the same thing in
			// terms compute-sanitizer's racecheck can follow, and costs nothing: all 128 threads just left the same wait.
			leaf_bar(wk);
			// G as fp32 [64 ch][64 pos] over the leaf region, channel c at word c*64 + ((c >> 4) & 3): a warp's store of one
			// channel is 32 consecutive words, and in the gather below the lanes of a warp — (rd, rh) x (pd & 1, ph) — land in
			// 32 different banks: (pd, ph) walk the multiples of 4 words, the skew separates the four (rd, rh) classes.
			// (The position-major layout this replaces put the 32 lanes of every gather load into 2 banks.)
#pragma unroll
			for (int j = 0; j < 32; ++j) {
				const int c = c0 + j;
				const uint32_t a = a_base + (uint32_t)(c * 64 + ((c >> 4) & 3) + wk.pos) * 4;
				asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v[j] + s_par[par::fold_b + c]) : "memory");
			}
			leaf_bar(wk);
			lap(4);
			// this thread: output row R = pos (D = R>>3, H = R&7), voxels W = chalf*4 .. +3
			{
				const int D = wk.pos >> 3, H = wk.pos & 7;
				const int rd = D & 1, rh = H & 1, pd = D >> 1, ph = H >> 1;
				const float fb = s_par[par::fin_b];
				float o[4];
#pragma unroll
				for (int j = 0; j < 4; ++j) {
					const int rw = j & 1, pw = wk.chalf * 2 + (j >> 1);
					const int r = rd * 4 + rh * 2 + rw;
					float sum = fb;
#pragma unroll
					for (int eps = 0; eps < 8; ++eps) {
						const int ed = (eps >> 2) & 1, eh = (eps >> 1) & 1, ew = eps & 1;
						const int qd = pd + (ed ? (rd ? 1 : -1) : 0), qh = ph + (eh ? (rh ? 1 : -1) : 0), qw = pw + (ew ? (rw ? 1 : -1) : 0);
						const bool ok = (unsigned)qd < 4u && (unsigned)qh < 4u && (unsigned)qw < 4u;
						const int c = r * 8 + eps;
						const uint32_t a = a_base + (uint32_t)(c * 64 + (r >> 1) + qd * 16 + qh * 4 + qw) * 4;  // (c >> 4) & 3 == r >> 1
						if (ok) {
							float g;
							asm volatile("ld.shared.f32 %0, [%1];" : "=f"(g) : "r"(a));
							sum += g;
						}
					}
					o[j] = sigmoid_f(sum);
				}
				if (leaf_ok) __stcs(reinterpret_cast<float4*>(voxels + leaf * 512 + wk.pos * 8 + wk.chalf * 4), make_float4(o[0], o[1], o[2], o[3]));
			}
			lap(5);
		}
		lap(7);
		if (kProf && tap_out) {
			float* o = tap_out + ((size_t)blockIdx.x * kThreads + threadIdx.x) * 4;
			o[0] = (float)wk.t_wait; o[1] = (float)wk.t_stage; o[2] = (float)wk.t_acc; o[3] = 0.f;
			if (threadIdx.x == 0) {  // sections of thread 0, behind the per-thread records
				float* e = tap_out + (size_t)gridDim.x * kThreads * 4 + (size_t)blockIdx.x * 8;
#pragma unroll
				for (int i = 0; i < 8; ++i) e[i] = (float)ep[i];
			}
		}
	}

	// ---- teardown: everybody is done with TMEM before the owner frees it ----
	tc_fence_before();
	__syncthreads();
	if (warp == kWorkWarps) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
}

}  // namespace

cudaError_t configure_decode_tc() {
	cudaError_t e = cudaFuncSetAttribute(decode_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
	if (e == cudaSuccess) e = cudaFuncSetAttribute(decode_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
	return e;
}

cudaError_t launch_decode_tc(const DecoderMmaWeights& w, const uint8_t* dev_indices, int64_t n_leaves, float* dev_voxels,
                             int num_sms, cudaStream_t stream, int tap_stage, float* tap_out) {
	if (n_leaves <= 0) return cudaSuccess;
	const int tiles = n_leaves <= 2 * (int64_t)num_sms ? 1 : kTiles;
	const int64_t groups = (n_leaves + 2 * tiles - 1) / (2 * tiles);
	const int grid = (int)(groups < (int64_t)num_sms ? groups : (int64_t)num_sms);
	if (tap_stage == 100)  // timing instrumentation: 4 floats per thread + 8 per CTA (tools/tc_pipeline_prof.py)
		decode_tc_kernel<true><<<grid, kThreads, kSmemBytes, stream>>>(w, dev_indices, n_leaves, dev_voxels, tiles, -1, tap_out);
	else
		decode_tc_kernel<false><<<grid, kThreads, kSmemBytes, stream>>>(w, dev_indices, n_leaves, dev_voxels, tiles, tap_stage, tap_out);
	return cudaGetLastError();
}

}  // namespace vqvdb
