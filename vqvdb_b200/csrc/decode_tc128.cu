// Tensor-core decoder for the 128-channel decoder of the reference's vec3 model (DecoderVec3, python/VQVAE_v2.py:302-325,
// with InferenceVQVAE.decode's gather + permute in front: python/save_for_inference.py:91-104): 64 uint8 indices in,
// 3 x 512 voxels out, one kernel, bf16 operands with fp32 accumulation on tcgen05 / TMEM.  BASELINE.json configs[3].
//
// It reuses the float decoder's scheme (decode_tc.cu) with the shapes changed:
//   * a GEMM tile is 128 rows = the 64 latent positions of two leaves; the gathered A operand goes through TMEM (TS-mode
//     tcgen05.mma): the tap-shifted neighbour row's channels, or a zero row outside the leaf, are copied from the
//     channels-last bf16 activation buffer into the row's TMEM lane by the row's own thread;
//   * every 128 -> 128 convolution runs as TWO passes of 64 output channels; a pass is 9 (kd, kh) tap pairs x 2
//     input-channel halves = 18 units, the three kw taps ride along N (N = 192) and are recombined when the accumulator
//     is read (a lane shuffle inside the warp, which also supplies the zero padding along w);
//   * the linear tail up_conv -> PixelShuffle3D -> final is folded on the host into three 128 -> 64 convolutions (one
//     per output channel) + the 8-term gather per voxel (decode_tc128.cuh), then tanh.
// What differs structurally: one CTA holds ONE tile (the activations of a leaf are 3 x 16 KB: conv input, conv output,
// residual), so the overlap the float kernel gets from two tiles drifting apart comes from two things instead:
//   * the accumulator is double-buffered (2 x 192 TMEM columns): pass p + 1's MMAs run while pass p's accumulator is read;
//   * staging and epilogues are done by DIFFERENT warps — 8 stager warps feed the MMAs and run ahead within a layer
//     (both passes of a conv read the same input), 8 epilogue warps own GroupNorm / residual / attention / the tail and
//     hand every finished layer input to the stagers through an mbarrier.
// Warp roles (576 threads): 0-7 epilogue (TMEM lane quadrant, channel half of the pass), 8-15 stagers (quadrant, channel
// half of the unit), 16 MMA issuer (whole warp, one elected lane), 17 TMA producer.
#include <cuda_bf16.h>

#include "decode_tc128.cuh"
#include "leaf_ops.cuh"
#include "tc128_ops.cuh"

namespace vqvdb {

namespace {

using tc128::column_sums;
using tc128::elect_one;
using tc128::lds128;
using tc128::make_desc_sw128;
using tc128::named_bar_sync;
using tc128::tc_commit;
using tc128::tc_fence_after;
using tc128::tc_fence_before;
using tc128::tmem_ld16_nowait;
using tc128::tmem_st16;
using tc128::tmem_wait_ld;

constexpr int kEpiWarps = 8, kStageWarps = 8;
constexpr int kIssuerWarp = kEpiWarps + kStageWarps, kProducerWarp = kIssuerWarp + 1;
constexpr int kThreads = (kProducerWarp + 1) * 32;        // 576
constexpr int kStages = 4;
constexpr uint32_t kUnitBytes = kDec128UnitBytes;        // [3 kw x 64 n][64 k] bf16
constexpr int kUnitsPerPass = kDec128UnitsPerPass;       // 18
constexpr int kPasses = kDec128Passes;                   // 13 per group
constexpr int kLayers = kDec128ConvLayers + 1;           // 5 convs + the folded tail (3 passes)
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kDCols = 192;                         // accumulator b: columns [b*192, b*192 + 192)
constexpr uint32_t kColA = 2 * kDCols;                   // A buffers: columns 384 + buf*32
#ifndef VQVDB_DEC128_A_BUFS
#define VQVDB_DEC128_A_BUFS 2  // 4 fit TMEM and were measured: 3.07 M vs 3.06 M leaves/s — the stager -> issuer hand-over is not the limiter
#endif
constexpr uint32_t kABufs = VQVDB_DEC128_A_BUFS;         // staged-A units in flight (the hand-over latency stager -> issuer hides behind them)
static_assert((kABufs & (kABufs - 1)) == 0 && kColA + kABufs * 32 <= kTmemCols, "A buffers: a power of two that fits TMEM");

// instruction descriptor: D = f32, A = B = bf16, both K-major, N = 192, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((192u >> 3) << 17) | ((128u >> 4) << 24);

// shared memory map (bytes)
constexpr uint32_t kOffRing = 0;
constexpr uint32_t kBufBytes = 16384;                    // [64 pos][128 ch] bf16, 256-byte rows, 16-byte chunks swizzled by pos & 7
constexpr uint32_t kXBytes = 16384 + 16;                 // residual x; later the fp32 G planes with their 3-word skew
constexpr uint32_t kLeafBytes = 2 * kBufBytes + kXBytes;
constexpr uint32_t kOffLeaf = kOffRing + kStages * kUnitBytes;
constexpr uint32_t kOffZero = kOffLeaf + 2 * kLeafBytes; // 256 zero bytes: the source row of out-of-leaf taps
constexpr uint32_t kOffBar = kOffZero + 256;
constexpr uint32_t kNumBars = 2 * kStages + 2 * kABufs + 2 + 2 + 1;  // w_full, w_empty, a_full[], a_empty[], d_full[2], d_empty[2], in_ready
constexpr uint32_t kOffTmemSlot = kOffBar + kNumBars * 8;
constexpr uint32_t kOffPar = (kOffTmemSlot + 16 + 15) & ~15u;
constexpr uint32_t kOffScratch = kOffPar + par128::total * 4;
// per-leaf scratch (floats): exch [2 slots][2 warps][2 halves][8], part [2 wil][128], scale [128], hid [32], idx [16 words]
constexpr uint32_t kScrExch = 0, kScrPart = 64, kScrScale = 320, kScrHid = 448, kScrIdx = 480, kScratchFloats = 496;
constexpr uint32_t kSmemBytes = kOffScratch + 2 * kScratchFloats * 4;
static_assert(kSmemBytes <= 227 * 1024, "decode_tc128 smem budget");
static_assert(kOffBar % 8 == 0 && kOffPar % 16 == 0 && kOffScratch % 16 == 0 && kLeafBytes % 16 == 0, "alignment");

__device__ __forceinline__ uint32_t bar_w_full(uint32_t bars, uint32_t s) { return bars + s * 8; }
__device__ __forceinline__ uint32_t bar_w_empty(uint32_t bars, uint32_t s) { return bars + (kStages + s) * 8; }
__device__ __forceinline__ uint32_t bar_a_full(uint32_t bars, uint32_t b) { return bars + (2 * kStages + b) * 8; }
__device__ __forceinline__ uint32_t bar_a_empty(uint32_t bars, uint32_t b) { return bars + (2 * kStages + kABufs + b) * 8; }
__device__ __forceinline__ uint32_t bar_d_full(uint32_t bars, uint32_t b) { return bars + (2 * kStages + 2 * kABufs + b) * 8; }
__device__ __forceinline__ uint32_t bar_d_empty(uint32_t bars, uint32_t b) { return bars + (2 * kStages + 2 * kABufs + 2 + b) * 8; }
__device__ __forceinline__ uint32_t bar_in_ready(uint32_t bars) { return bars + (2 * kStages + 2 * kABufs + 4) * 8; }

// D[tmem_d] (+)= A[tmem_a] (128 x 16 bf16, TMEM) * B[desc] (192 x 16 bf16, shared)^T
__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t accumulate) {
	tc128::tc_mma_ts(tmem_d, tmem_a, bdesc, kIdesc, accumulate);
}
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) { tc128::sts128(a, make_uint4(x, y, z, w)); }

// Physical byte offset, inside a [64 pos][128 ch] bf16 buffer, of the 16-byte chunk holding channels 8*c16 .. 8*c16+7 of row pos.
__device__ __forceinline__ uint32_t chunk_off(int pos, int c16) {
	return (uint32_t)pos * 256u + ((uint32_t)((c16 & 8) | ((c16 & 7) ^ (pos & 7))) << 4);
}

// An epilogue thread: one GEMM row (latent position of one of the two leaves) x 32 of the 64 channels of a pass.
struct Epi {
	int quad, chalf, lane, row, leaf_slot, pos, d, h, w, wil;
	uint32_t bars, tmem_lane;
	uint32_t passes = 0;  // accumulator hand-overs so far: buffer = passes & 1, phase = (passes >> 1) & 1
	uint32_t reds = 0;    // half-leaf reductions so far (alternates the exchange slot)
};
__device__ __forceinline__ void leaf_bar(const Epi& e) { named_bar_sync(1 + e.leaf_slot, 128); }
__device__ __forceinline__ void half_bar(const Epi& e) { named_bar_sync(3 + e.leaf_slot * 2 + e.chalf, 64); }

// This thread's 32 output channels of the finished pass (the three kw partials combined across neighbouring rows); the
// accumulator buffer is handed back to the issuer as soon as it has been read.
__device__ __forceinline__ void take_accumulator(Epi& e, float (&v)[32]) {
	const uint32_t b = e.passes & 1u;
	mbar_wait(bar_d_full(e.bars, b), (e.passes >> 1) & 1u);
	tc_fence_after();
	const uint32_t base = e.tmem_lane + b * kDCols + e.chalf * 32;
	const bool has_lo = e.w > 0, has_hi = e.w < 3;
#pragma unroll
	for (int part = 0; part < 2; ++part) {
		float a[16], m[16], c[16];
		tmem_ld16_nowait(base + part * 16, a);        // kw = 0: belongs to the row at w + 1
		tmem_ld16_nowait(base + 64 + part * 16, m);   // kw = 1
		tmem_ld16_nowait(base + 128 + part * 16, c);  // kw = 2: belongs to the row at w - 1
		tmem_wait_ld();
#pragma unroll
		for (int j = 0; j < 16; ++j) {
			const float lo = __shfl_up_sync(0xffffffffu, a[j], 1);
			const float hi = __shfl_down_sync(0xffffffffu, c[j], 1);
			v[part * 16 + j] = m[j] + (has_lo ? lo : 0.f) + (has_hi ? hi : 0.f);
		}
	}
	tc_fence_before();
	__syncwarp();
	if (e.lane == 0) mbar_arrive(bar_d_empty(e.bars, b));
	++e.passes;
}

// Sum N per-thread values over the 64 rows of this thread's leaf, among the threads of its channel half (2 warps).
template <int N>
__device__ __forceinline__ void half_allreduce(float (&v)[N], Epi& e, float* exch /* [2 slots][2 warps][2 halves][8] */) {
	static_assert(N <= 8, "exchange slot size");
#pragma unroll
	for (int i = 0; i < N; ++i) v[i] = warp_sum(v[i]);
	float* x = exch + (e.reds & 1u) * 32;
	if (e.lane == 0) {
#pragma unroll
		for (int i = 0; i < N; ++i) x[(e.wil * 2 + e.chalf) * 8 + i] = v[i];
	}
	half_bar(e);
#pragma unroll
	for (int i = 0; i < N; ++i) v[i] = x[e.chalf * 8 + i] + x[(2 + e.chalf) * 8 + i];
	++e.reds;  // the next reduction uses the other slot; this one is rewritten only after another barrier has been passed
}

// GroupNorm(8, 128) over the leaf for this thread's two groups of 16 channels: v -> normalised (+ ReLU) in place.
__device__ __forceinline__ void group_norm_relu(float (&v)[32], Epi& e, float* exch, const float* gamma, const float* beta) {
	float st[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
	for (int j = 0; j < 32; ++j) {
		st[j >> 4] += v[j];
		st[2 + (j >> 4)] = fmaf(v[j], v[j], st[2 + (j >> 4)]);
	}
	half_allreduce<4>(st, e, exch);
#pragma unroll
	for (int g = 0; g < 2; ++g) {
		const float mean = st[g] * (1.f / 1024.f);
		const float rstd = 1.f / sqrtf(fmaxf(st[2 + g] * (1.f / 1024.f) - mean * mean, 0.f) + kGnEps);
#pragma unroll
		for (int j = 0; j < 16; ++j) v[g * 16 + j] = fmaxf((v[g * 16 + j] - mean) * rstd * gamma[g * 16 + j] + beta[g * 16 + j], 0.f);
	}
}

// this thread's 32 channels (first channel 8*c16_0) of row pos -> bf16
__device__ __forceinline__ void store_row32(uint32_t buf, int pos, int c16_0, const float (&v)[32]) {
#pragma unroll
	for (int q = 0; q < 4; ++q)
		sts128(buf + chunk_off(pos, c16_0 + q), pack_bf16(v[8 * q], v[8 * q + 1]), pack_bf16(v[8 * q + 2], v[8 * q + 3]),
		       pack_bf16(v[8 * q + 4], v[8 * q + 5]), pack_bf16(v[8 * q + 6], v[8 * q + 7]));
}
__device__ __forceinline__ void load_row32(uint32_t buf, int pos, int c16_0, float (&x)[32]) {
#pragma unroll
	for (int q = 0; q < 4; ++q) {
		const uint4 raw = lds128(buf + chunk_off(pos, c16_0 + q));
		float2 f;
		f = unpack_bf16(raw.x); x[8 * q] = f.x; x[8 * q + 1] = f.y;
		f = unpack_bf16(raw.y); x[8 * q + 2] = f.x; x[8 * q + 3] = f.y;
		f = unpack_bf16(raw.z); x[8 * q + 4] = f.x; x[8 * q + 5] = f.y;
		f = unpack_bf16(raw.w); x[8 * q + 6] = f.x; x[8 * q + 7] = f.y;
	}
}

__global__ void __launch_bounds__(kThreads, 1)
decode_tc128_kernel(const Decoder128Weights w, const uint8_t* __restrict__ indices, int64_t n_leaves, float* __restrict__ voxels,
                    int tap_stage, float* __restrict__ tap_out) {
	extern __shared__ __align__(1024) uint8_t smem[];
	const uint32_t s_base = smem_u32(smem);
	const uint32_t ring = s_base + kOffRing;
	const uint32_t bars = s_base + kOffBar;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int64_t n_groups = (n_leaves + 1) / 2;
	uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffTmemSlot);
	float* s_par = reinterpret_cast<float*>(smem + kOffPar);

	for (int i = threadIdx.x; i < par128::total; i += kThreads) s_par[i] = __ldg(w.par + i);
	if (threadIdx.x < 64) reinterpret_cast<uint32_t*>(smem + kOffZero)[threadIdx.x] = 0u;
	if (threadIdx.x == 0) {
		for (uint32_t s = 0; s < kStages; ++s) {
			mbar_init(bar_w_full(bars, s), 1);
			mbar_init(bar_w_empty(bars, s), 1);
		}
		for (uint32_t b = 0; b < kABufs; ++b) {
			mbar_init(bar_a_full(bars, b), kStageWarps);
			mbar_init(bar_a_empty(bars, b), 1);
		}
		for (uint32_t b = 0; b < 2; ++b) {
			mbar_init(bar_d_full(bars, b), 1);
			mbar_init(bar_d_empty(bars, b), kEpiWarps);
		}
		mbar_init(bar_in_ready(bars), kEpiWarps);
		mbar_fence_init();
	}
	if (warp == kIssuerWarp) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
	}
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem = *tmem_slot;
	const int64_t my_groups = blockIdx.x < n_groups ? (n_groups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

	if (warp == kProducerWarp) {
		// ===================== TMA producer: one contiguous 24 KB unit per (pass, tap pair, input half) =====================
		if (lane == 0) {
			const uint32_t total = (uint32_t)(my_groups * kDec128Units);
#pragma unroll 1
			for (uint32_t issued = 0; issued < total; ++issued) {
				const uint32_t s = issued % kStages, u = issued % kDec128Units;
				mbar_wait(bar_w_empty(bars, s), ((issued / kStages) & 1u) ^ 1u);
				mbar_arrive_expect_tx(bar_w_full(bars, s), kUnitBytes);
				tma_load_1d(ring + s * kUnitBytes, w.units + (size_t)u * kUnitBytes, kUnitBytes, bar_w_full(bars, s));
			}
		}
		__syncwarp();
	} else if (warp == kIssuerWarp) {
		// ===================== MMA issuer (whole warp, one elected lane issues) =====================
		const bool leader = elect_one();
		uint32_t unit = 0, pass = 0;
#pragma unroll 1
		for (int64_t g = 0; g < my_groups; ++g) {
#pragma unroll 1
			for (int p = 0; p < kPasses; ++p, ++pass) {
				const uint32_t db = pass & 1u;
				// the epilogue warps have read the previous result out of this accumulator buffer
				mbar_wait(bar_d_empty(bars, db), ((pass >> 1) & 1u) ^ 1u);
				tc_fence_after();
#pragma unroll 1
				for (int u = 0; u < kUnitsPerPass; ++u, ++unit) {
					const uint32_t s = unit % kStages, ab = unit % kABufs;
					mbar_wait(bar_w_full(bars, s), (unit / kStages) & 1u);
					mbar_wait(bar_a_full(bars, ab), (unit / kABufs) & 1u);
					tc_fence_after();
					const uint64_t bdesc = make_desc_sw128(ring + s * kUnitBytes);
#pragma unroll
					for (uint32_t kk = 0; kk < 4; ++kk)
						if (leader) tc_mma_ts(tmem + db * kDCols, tmem + kColA + ab * 32 + kk * 8, bdesc + (uint64_t)(kk * 2), (u > 0 || kk > 0) ? 1u : 0u);
					if (leader) tc_commit(bar_a_empty(bars, ab));
					if (leader) tc_commit(bar_w_empty(bars, s));
					if (u == kUnitsPerPass - 1 && leader) tc_commit(bar_d_full(bars, db));
				}
			}
		}
		__syncwarp();
	} else if (warp >= kEpiWarps) {
		// ===================== stager warps: tap-shifted activation rows -> TMEM A buffers =====================
		const int sw = warp - kEpiWarps, quad = sw & 3, chalf = sw >> 2;
		const int row = quad * 32 + lane, leaf_slot = row >> 6, pos = row & 63;
		const int pd = pos >> 4, ph = (pos >> 2) & 3;
		const uint32_t leaf_base = s_base + kOffLeaf + (uint32_t)leaf_slot * kLeafBytes;
		const uint32_t zero_row = s_base + kOffZero;
		const uint32_t tmem_lane = tmem + ((uint32_t)(quad * 32) << 16);
		uint32_t unit = 0, layer = 0;
#pragma unroll 1
		for (int64_t g = 0; g < my_groups; ++g) {
#pragma unroll 1
			for (int l = 0; l < kLayers; ++l, ++layer) {
				// the layer's input (buffer l & 1 of this leaf) is complete
				if (lane == 0) mbar_wait(bar_in_ready(bars), layer & 1u);
				__syncwarp();
				const uint32_t in_buf = leaf_base + (uint32_t)(l & 1) * kBufBytes;
				const int passes = l < kDec128ConvLayers ? 2 : 3;
#pragma unroll 1
				for (int pu = 0; pu < passes * kUnitsPerPass; ++pu, ++unit) {
					const int u = pu % kUnitsPerPass, t = u >> 1, khalf = u & 1;
					const int td = t / 3, th = t - td * 3;
					const bool ok = (unsigned)(pd + td - 1) < 4u && (unsigned)(ph + th - 1) < 4u;
					const int p2 = pos + (td - 1) * 16 + (th - 1) * 4;
					uint32_t r[16];
#pragma unroll
					for (int q = 0; q < 4; ++q) {
						const uint4 v = lds128(ok ? in_buf + chunk_off(p2, khalf * 8 + chalf * 4 + q) : zero_row + q * 16);
						r[4 * q] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
					}
					const uint32_t ab = unit % kABufs;
					if (lane == 0) mbar_wait(bar_a_empty(bars, ab), ((unit / kABufs) & 1u) ^ 1u);
					__syncwarp();
					tc_fence_after();
					tmem_st16(tmem_lane + kColA + ab * 32 + chalf * 16, r);
					asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
					tc_fence_before();
					__syncwarp();
					if (lane == 0) mbar_arrive(bar_a_full(bars, ab));
				}
			}
		}
	} else {
		// ===================== epilogue warps =====================
		Epi e;
		e.quad = warp & 3;
		e.chalf = warp >> 2;
		e.lane = lane;
		e.row = e.quad * 32 + lane;
		e.leaf_slot = e.row >> 6;
		e.wil = (e.row >> 5) & 1;
		e.pos = e.row & 63;
		e.d = e.pos >> 4;
		e.h = (e.pos >> 2) & 3;
		e.w = e.pos & 3;
		e.bars = bars;
		e.tmem_lane = tmem + ((uint32_t)(e.quad * 32) << 16);
		const uint32_t leaf_base = s_base + kOffLeaf + (uint32_t)e.leaf_slot * kLeafBytes;
		const uint32_t buf0 = leaf_base, buf1 = leaf_base + kBufBytes, xbuf = leaf_base + 2 * kBufBytes;
		uint8_t* buf0_g = smem + kOffLeaf + e.leaf_slot * kLeafBytes;
		float* scratch = reinterpret_cast<float*>(smem + kOffScratch) + e.leaf_slot * kScratchFloats;
		float* exch = scratch + kScrExch;
		float* s_part = scratch + kScrPart;
		float* s_scale = scratch + kScrScale;
		float* s_hid = scratch + kScrHid;
		uint32_t* s_idx = reinterpret_cast<uint32_t*>(scratch + kScrIdx);
		const int tl = e.chalf * 64 + e.pos;  // thread index within the leaf, 0..127
		// layer input l lives in buffer l & 1; its output goes to buffer (l + 1) & 1
		auto signal_input_ready = [&]() {
			__syncwarp();
			if (lane == 0) mbar_arrive(bar_in_ready(bars));
		};

		// the index words of the NEXT group's leaf are requested at the top of a group (HBM, ~2 k cycles away)
		auto load_index_word = [&](int64_t g) -> uint32_t {
			const int64_t lf = (blockIdx.x + g * gridDim.x) * 2 + e.leaf_slot;
			return (tl < 16 && g < my_groups && lf < n_leaves) ? __ldcs(reinterpret_cast<const uint32_t*>(indices + lf * 64) + tl) : 0u;
		};
		uint32_t idx_word = load_index_word(0);
#pragma unroll 1
		for (int64_t g = 0; g < my_groups; ++g) {
			const int64_t grp = blockIdx.x + g * gridDim.x;
			const int64_t leaf = grp * 2 + e.leaf_slot;
			const bool leaf_ok = leaf < n_leaves;

			// ---- gather: Q[pos][0..127] = codebook_bf16[idx[pos]] -> buffer 0 (a spare slot decodes code 0) ----
			if (tl < 16) s_idx[tl] = idx_word;
			idx_word = load_index_word(g + 1);
			leaf_bar(e);  // also: every thread of the leaf is done with the previous group's G planes and buffers
			{
				const uint8_t* idx8 = reinterpret_cast<const uint8_t*>(s_idx);
#pragma unroll
				for (int i = tl; i < 64 * 16; i += 128) {  // eight L2 loads per thread, all in flight
					const int pos = i >> 4, c = i & 15;
					const uint4 v = __ldg(reinterpret_cast<const uint4*>(w.emb_bf16 + (size_t)idx8[pos] * 128) + c);
					*reinterpret_cast<uint4*>(buf0_g + chunk_off(pos, c)) = v;
				}
			}
			signal_input_ready();  // layer 0 (stem): the mbarrier's release/acquire orders the stores above before the stagers' loads

			float v[32];
			// ---- layer 0: stem (128 -> 128) ; GroupNorm + ReLU -> x ; res0.gn1 + ReLU -> conv1 input (buffer 1) ----
#pragma unroll 1
			for (int hh = 0; hh < 2; ++hh) {
				const int c0 = hh * 64 + e.chalf * 32, c16 = c0 >> 3;
				take_accumulator(e, v);
#pragma unroll
				for (int j = 0; j < 32; ++j) v[j] += s_par[par128::stem_b + c0 + j];
				group_norm_relu(v, e, exch, s_par + par128::stem_gn_w + c0, s_par + par128::stem_gn_b + c0);
				if (tap_stage == 0 && leaf_ok) {
#pragma unroll
					for (int j = 0; j < 32; ++j) tap_out[leaf * 8192 + (c0 + j) * 64 + e.pos] = v[j];
				}
				store_row32(xbuf, e.pos, c16, v);  // residual x (thread-private)
				group_norm_relu(v, e, exch, s_par + par128::res0 + par128::gn1_w + c0, s_par + par128::res0 + par128::gn1_b + c0);
				store_row32(buf1, e.pos, c16, v);
			}
			signal_input_ready();  // layer 1

			// ---- two residual blocks: conv1 (layer 1 + 2r: in buffer 1, out buffer 0), conv2 (layer 2 + 2r: in 0, out 1) ----
#pragma unroll 1
			for (int r = 0; r < 2; ++r) {
				const float* rp = s_par + par128::res0 + r * par128::res_stride;
#pragma unroll 1
				for (int hh = 0; hh < 2; ++hh) {
					const int c0 = hh * 64 + e.chalf * 32;
					take_accumulator(e, v);
#pragma unroll
					for (int j = 0; j < 32; ++j) v[j] += rp[par128::c1_b + c0 + j];
					group_norm_relu(v, e, exch, rp + par128::gn2_w + c0, rp + par128::gn2_b + c0);
					store_row32(buf0, e.pos, c0 >> 3, v);
				}
				signal_input_ready();  // conv2's input
#pragma unroll 1
				for (int hh = 0; hh < 2; ++hh) {
					const int c0 = hh * 64 + e.chalf * 32, c16 = c0 >> 3;
					take_accumulator(e, v);
					float xo[32];
					load_row32(xbuf, e.pos, c16, xo);
#pragma unroll
					for (int j = 0; j < 32; ++j) v[j] = xo[j] + kResScale * (v[j] + rp[par128::c2_b + c0 + j]);
					store_row32(xbuf, e.pos, c16, v);  // x' (thread-private)
					if (r == 0) {
						group_norm_relu(v, e, exch, rp + par128::res_stride + par128::gn1_w + c0, rp + par128::res_stride + par128::gn1_b + c0);
						store_row32(buf1, e.pos, c16, v);
					} else {
						if (tap_stage == 1 && leaf_ok) {
#pragma unroll
							for (int j = 0; j < 32; ++j) tap_out[leaf * 8192 + (c0 + j) * 64 + e.pos] = v[j];
						}
						const float cs = column_sums(v, lane);  // channel c0 + lane over this warp's 32 rows
						s_part[e.wil * 128 + c0 + lane] = cs;
					}
				}
				if (r == 0) signal_input_ready();  // res1.conv1's input
			}

			// ---- ChannelAttention(128): mean over the leaf -> 128 -> 32 -> 128 -> sigmoid ; x' * scale -> tail input (buffer 1) ----
			leaf_bar(e);
			{
				// hidden = relu(fc0 [32][128] . mean): 4 threads per hidden unit, 32 channels each
				const int unit = tl >> 2, part = tl & 3;
				float s = 0.f;
#pragma unroll 8
				for (int i = 0; i < 32; ++i) {
					const int c = i * 4 + part;
					s = fmaf(__ldg(w.fc0 + unit * 128 + c), (s_part[c] + s_part[128 + c]) * (1.f / 64.f), s);
				}
				s += __shfl_xor_sync(0xffffffffu, s, 1);
				s += __shfl_xor_sync(0xffffffffu, s, 2);
				if (part == 0) s_hid[unit] = fmaxf(s, 0.f);
			}
			leaf_bar(e);
			{
				float s = 0.f;
#pragma unroll 8
				for (int j = 0; j < 32; ++j) s = fmaf(__ldg(w.fc2_t + j * 128 + tl), s_hid[j], s);  // [hidden][channel]: a warp reads one line
				s_scale[tl] = sigmoid_f(s);
			}
			leaf_bar(e);
#pragma unroll 1
			for (int hh = 0; hh < 2; ++hh) {
				const int c0 = hh * 64 + e.chalf * 32, c16 = c0 >> 3;
				load_row32(xbuf, e.pos, c16, v);
#pragma unroll
				for (int j = 0; j < 32; ++j) v[j] *= s_scale[c0 + j];
				if (tap_stage == 2 && leaf_ok) {
#pragma unroll
					for (int j = 0; j < 32; ++j) tap_out[leaf * 8192 + (c0 + j) * 64 + e.pos] = v[j];
				}
				store_row32(buf1, e.pos, c16, v);
			}
			signal_input_ready();  // layer 5: the folded tail

			// ---- folded tail, once per output channel: G_c = conv(a; Wg_c) + bg_c as fp32 [64 ch][64 pos] over the (now idle)
			//      residual region, then out[c][2p + r] = tanh(fin_b[c] + sum over the in-grid cells p + e(r, eps) of G_c[p + e][r*8 + eps]) ----
			const int ch0 = e.chalf * 32;
#pragma unroll 1
			for (int c = 0; c < 3; ++c) {
				take_accumulator(e, v);
				leaf_bar(e);  // the previous channel's gather is done: G may be overwritten
				// channel n at word n*64 + ((n >> 4) & 3): stores and the gather below are bank-conflict free (decode_tc.cu)
#pragma unroll
				for (int j = 0; j < 32; ++j) {
					const int n = ch0 + j;
					const uint32_t a = xbuf + (uint32_t)(n * 64 + ((n >> 4) & 3) + e.pos) * 4;
					asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v[j] + s_par[par128::fold_b + c * 64 + n]) : "memory");
				}
				leaf_bar(e);
				{
					// this thread: output row R = pos (D = R>>3, H = R&7), voxels W = chalf*4 .. +3
					const int D = e.pos >> 3, H = e.pos & 7;
					const int rd = D & 1, rh = H & 1, pd = D >> 1, ph = H >> 1;
					const float fb = s_par[par128::fin_b + c];
					float o[4];
#pragma unroll
					for (int j = 0; j < 4; ++j) {
						const int rw = j & 1, pw = e.chalf * 2 + (j >> 1);
						const int r = rd * 4 + rh * 2 + rw;
						float sum = fb;
#pragma unroll
						for (int eps = 0; eps < 8; ++eps) {
							const int ed = (eps >> 2) & 1, eh = (eps >> 1) & 1, ew = eps & 1;
							const int qd = pd + (ed ? (rd ? 1 : -1) : 0), qh = ph + (eh ? (rh ? 1 : -1) : 0), qw = pw + (ew ? (rw ? 1 : -1) : 0);
							const bool ok = (unsigned)qd < 4u && (unsigned)qh < 4u && (unsigned)qw < 4u;
							const int n = r * 8 + eps;
							const uint32_t a = xbuf + (uint32_t)(n * 64 + (r >> 1) + qd * 16 + qh * 4 + qw) * 4;  // (n >> 4) & 3 == r >> 1
							if (ok) {
								float gv;
								asm volatile("ld.shared.f32 %0, [%1];" : "=f"(gv) : "r"(a));
								sum += gv;
							}
						}
						o[j] = tanhf(sum);  // DecoderVec3 ends in tanh (VQVAE_v2.py:325)
					}
					if (leaf_ok) __stcs(reinterpret_cast<float4*>(voxels + leaf * 1536 + c * 512 + e.pos * 8 + e.chalf * 4), make_float4(o[0], o[1], o[2], o[3]));
				}
			}
		}
	}

	// ---- teardown: everybody is done with TMEM before the owner frees it ----
	tc_fence_before();
	__syncthreads();
	if (warp == kIssuerWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
}

}  // namespace

cudaError_t configure_decode_tc128() {
	return cudaFuncSetAttribute(decode_tc128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
}

cudaError_t launch_decode_tc128(const Decoder128Weights& w, const uint8_t* dev_indices, int64_t n_leaves, float* dev_voxels,
                                int num_sms, cudaStream_t stream, int tap_stage, float* tap_out) {
	if (n_leaves <= 0) return cudaSuccess;
	const int64_t groups = (n_leaves + 1) / 2;
	const int grid = (int)(groups < (int64_t)num_sms ? groups : (int64_t)num_sms);
	decode_tc128_kernel<<<grid, kThreads, kSmemBytes, stream>>>(w, dev_indices, n_leaves, dev_voxels, tap_stage, tap_out);
	return cudaGetLastError();
}

}  // namespace vqvdb
