// Tensor-core decoder on the 5th-generation tensor cores: tcgen05.mma with TMEM accumulators, the gathered
// (im2col) A operand written into TMEM by the threads themselves, weights streamed by TMA.
//
// Same arithmetic contract as decode_mma.cu (bf16 operands, fp32 accumulate; reference:
// python/save_for_inference.py:91-104 + python/VQVAE_v2.py:253-275), different machine mapping:
//
//   * GEMM tile = 128 rows x 64 columns: the 64 latent positions of TWO leaves x 64 output channels,
//     one tcgen05.mma.cta_group::1.kind::f16 per 16 input channels.  A CTA keeps 4 such tiles (8 leaves) in flight;
//     their fp32 accumulators live in TMEM (4 x 64 columns), never in registers.
//   * a 3x3x3 convolution over a 4^3 leaf cannot be described to the tensor core by a shared-memory descriptor
//     (rows of 4 voxels, zero padding), so the A operand goes through TMEM: for every (tap, 64-channel) unit each
//     of a tile's 128 threads loads ITS row — the shifted neighbour's 64 channels, or zeros outside the leaf — from
//     the channels-last bf16 activation buffer in shared memory and writes it to its TMEM lane (tcgen05.st).
//     A is double-buffered per tile (2 x 32 columns), so staging unit u+1 overlaps the MMAs of unit u.
//   * B (weights) is the same stream of 216 pre-swizzled 8 KB units [64 n][64 k] as decode_mma.cu: exactly the
//     canonical SWIZZLE_128B K-major layout a UMMA shared-memory descriptor expects.  An 8-stage ring is filled by
//     1-D TMA bulk copies; all 8 leaves share every unit.
//   * warp roles: 16 worker warps (4 per tile = the four TMEM lane quadrants) stage A and run the epilogues
//     (GroupNorm, residual, channel attention, pixel-shuffle + final conv on FFMA, sigmoid, stores); 4 control
//     warps, one elected lane each, issue the MMAs of one tile and signal completion with tcgen05.commit, so the
//     tiles run out of phase and one tile's epilogue hides behind the others' MMAs; one more lane issues the TMA.
//     (Letting a worker thread issue its tile's MMAs after a tile barrier was measured 27 % slower.)
//   * synchronisation is mbarrier-only on the MMA path: w_full/w_empty per ring stage, a_full/a_empty per (tile,
//     A buffer), d_full per tile; the two warps that share a leaf meet on a 64-thread named barrier for the
//     per-leaf reductions.
#include <cuda_bf16.h>

#include "decode_mma.cuh"
#include "leaf_ops.cuh"
#include "ptx_utils.cuh"

namespace vqvdb {

namespace {

constexpr int kTiles = 4;
constexpr int kLeavesPerCta = 2 * kTiles;
constexpr int kWorkWarps = 4 * kTiles;
constexpr int kCtrlWarps = kTiles;               // one MMA issuer per tile
constexpr int kThreads = (kWorkWarps + kCtrlWarps) * 32;  // 640
constexpr int kStages = 8;
constexpr uint32_t kUnitBytes = 8192;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColD = 0;      // tile t accumulator: columns [t*64, t*64+64)
constexpr uint32_t kColA = 256;    // tile t A buffers:   columns 256 + t*64 + buf*32

// instruction descriptor: D=f32, A=B=bf16, both K-major, N=64, M=128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

// shared memory map (bytes)
constexpr uint32_t kOffRing = 0;
constexpr uint32_t kOffLeaf = kOffRing + kStages * kUnitBytes;           // 8 x 16 KB
constexpr uint32_t kLeafBytes = 16384;
constexpr uint32_t kOffBar = kOffLeaf + kLeavesPerCta * kLeafBytes;      // mbarriers
constexpr uint32_t kNumBars = 2 * kStages + 4 * kTiles + kTiles;         // w_full, w_empty, a_full[t][2], a_empty[t][2], d_full[t]
constexpr uint32_t kOffTmemSlot = kOffBar + kNumBars * 8;
constexpr uint32_t kOffFinW = kOffTmemSlot + 16;                         // 864 floats
constexpr uint32_t kOffScratch = kOffFinW + 864 * 4;                     // per leaf: 256 floats
constexpr uint32_t kScratchFloats = 256;
constexpr uint32_t kSmemBytes = kOffScratch + kLeavesPerCta * kScratchFloats * 4;
static_assert(kSmemBytes <= 227 * 1024, "decode_tc smem budget");

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
// D[tmem_d] (+)= A[tmem_a] (128 x 16 bf16, TMEM) * B[desc] (64 x 16 bf16, shared)^T
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
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
	asm volatile(
	    "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
	    ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
	    "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
	    "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
	    "r"(r[30]), "r"(r[31])
	    : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
	uint32_t o[32];
	asm volatile(
	    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
	    "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
	    : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7]), "=r"(o[8]), "=r"(o[9]),
	      "=r"(o[10]), "=r"(o[11]), "=r"(o[12]), "=r"(o[13]), "=r"(o[14]), "=r"(o[15]), "=r"(o[16]), "=r"(o[17]), "=r"(o[18]), "=r"(o[19]),
	      "=r"(o[20]), "=r"(o[21]), "=r"(o[22]), "=r"(o[23]), "=r"(o[24]), "=r"(o[25]), "=r"(o[26]), "=r"(o[27]), "=r"(o[28]), "=r"(o[29]),
	      "=r"(o[30]), "=r"(o[31])
	    : "r"(taddr));
	asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
	for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(o[j]);
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
	asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Per-thread view of the work: which tile / TMEM lane quadrant / leaf / latent position this thread is.
struct Worker {
	int tile, quad, lane, row;      // row = quad*32 + lane in [0,128)
	int leaf_slot;                  // tile*2 + (row >> 6)
	int pos, d, h, w;               // latent position within the leaf
	int wil;                        // warp-in-leaf: 0 or 1
	uint32_t unit = 0;              // units staged so far (same sequence in every worker and in the control thread)
	uint32_t passes = 0;            // accumulator hand-overs so far (parity of d_full)
	uint32_t bars, tmem_lane;       // mbarrier base; tmem base + (quad*32 << 16)
	long long t_wait = 0, t_stage = 0, t_acc = 0;  // kProf only: cycles waiting for a free A buffer / staging / waiting for d_full
};
template <bool P> __device__ __forceinline__ long long prof_clock() { return P ? clock64() : 0; }

// Stage one A unit: this thread's row (64 bf16 channels = 32 words) -> its TMEM lane, then hand the buffer to the MMA.
template <bool kProf>
__device__ __forceinline__ void stage_unit(Worker& wk, bool valid, uint32_t src_row /*smem addr of the 128-B half row*/, uint32_t swz) {
	const uint32_t buf = wk.unit & 1u;
	const long long c0 = prof_clock<kProf>();
	if (wk.lane == 0) mbar_wait(bar_a_empty(wk.bars, wk.tile, buf), ((wk.unit >> 1) & 1u) ^ 1u);
	__syncwarp();
	tc_fence_after();
	const long long c1 = prof_clock<kProf>();
	uint32_t r[32];
	if (valid) {
#pragma unroll
		for (int q = 0; q < 8; ++q) {
			const uint32_t a = src_row + (((uint32_t)q ^ swz) << 4);
			asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[4 * q]), "=r"(r[4 * q + 1]), "=r"(r[4 * q + 2]), "=r"(r[4 * q + 3]) : "r"(a));
		}
	} else {
#pragma unroll
		for (int j = 0; j < 32; ++j) r[j] = 0u;
	}
	tmem_st32(wk.tmem_lane + kColA + wk.tile * 64 + buf * 32, r);
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

// One 3x3x3 conv: stage 27 * HALVES units from the leaf's activation rows (ROW_BYTES = HALVES * 128).
template <int HALVES, bool kProf>
__device__ __forceinline__ void stage_conv(Worker& wk, uint32_t act_base) {
	constexpr uint32_t ROW_BYTES = HALVES * 128;
#pragma unroll 1
	for (int tap = 0; tap < 27; ++tap) {
		const int td = tap / 9, th = (tap / 3) % 3, tw = tap % 3;
		const bool ok = (unsigned)(wk.d + td - 1) < 4u && (unsigned)(wk.h + th - 1) < 4u && (unsigned)(wk.w + tw - 1) < 4u;
		const int p2 = wk.pos + (td - 1) * 16 + (th - 1) * 4 + (tw - 1);
#pragma unroll
		for (int half = 0; half < HALVES; ++half) stage_unit<kProf>(wk, ok, act_base + (uint32_t)p2 * ROW_BYTES + half * 128, (uint32_t)p2 & 7u);
	}
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

// Sum `n` per-thread values over the 64 rows of this thread's leaf (2 warps): butterfly + one exchange.
template <int N>
__device__ __forceinline__ void leaf_allreduce(float (&v)[N], const Worker& wk, float* exch /*[2][N] per leaf*/) {
#pragma unroll
	for (int i = 0; i < N; ++i) v[i] = warp_sum(v[i]);
	if (wk.lane == 0) {
#pragma unroll
		for (int i = 0; i < N; ++i) exch[wk.wil * N + i] = v[i];
	}
	named_bar_sync(1 + wk.leaf_slot, 64);
#pragma unroll
	for (int i = 0; i < N; ++i) v[i] = exch[i] + exch[N + i];
	named_bar_sync(1 + wk.leaf_slot, 64);  // exch may be reused right away
}

// row-major [64 pos][64 ch] bf16 buffer, 128-B rows, 16-B chunks swizzled by pos & 7: store this thread's 32 values
// (channels c0..c0+31) of its row.
__device__ __forceinline__ void store_row_half(uint32_t buf_base, int pos, int half, const float (&v)[32]) {
#pragma unroll
	for (int q = 0; q < 4; ++q) {
		const uint32_t chunk = (uint32_t)(half * 4 + q) ^ ((uint32_t)pos & 7u);
		const uint32_t a = buf_base + (uint32_t)pos * 128 + (chunk << 4);
		asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(pack_bf16(v[8 * q], v[8 * q + 1])),
		             "r"(pack_bf16(v[8 * q + 2], v[8 * q + 3])), "r"(pack_bf16(v[8 * q + 4], v[8 * q + 5])),
		             "r"(pack_bf16(v[8 * q + 6], v[8 * q + 7])));
	}
}
__device__ __forceinline__ void load_row_half(uint32_t buf_base, int pos, int half, float (&v)[32]) {
#pragma unroll
	for (int q = 0; q < 4; ++q) {
		const uint32_t chunk = (uint32_t)(half * 4 + q) ^ ((uint32_t)pos & 7u);
		const uint32_t a = buf_base + (uint32_t)pos * 128 + (chunk << 4);
		uint32_t w0, w1, w2, w3;
		asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(a));
		float2 f;
		f = unpack_bf16(w0); v[8 * q] = f.x; v[8 * q + 1] = f.y;
		f = unpack_bf16(w1); v[8 * q + 2] = f.x; v[8 * q + 3] = f.y;
		f = unpack_bf16(w2); v[8 * q + 4] = f.x; v[8 * q + 5] = f.y;
		f = unpack_bf16(w3); v[8 * q + 6] = f.x; v[8 * q + 7] = f.y;
	}
}

// GroupNorm(8, 64) statistics of (accumulator + bias) over the leaf: mean/rstd per group.
__device__ __forceinline__ void gn_stats_from_tmem(const Worker& wk, const float* __restrict__ bias, float* exch, float (&mean)[8],
                                                   float (&rstd)[8]) {
	float st[16];
#pragma unroll
	for (int i = 0; i < 16; ++i) st[i] = 0.f;
#pragma unroll
	for (int half = 0; half < 2; ++half) {
		float v[32];
		tmem_ld32(wk.tmem_lane + kColD + wk.tile * 64 + half * 32, v);
#pragma unroll
		for (int j = 0; j < 32; ++j) {
			const float x = v[j] + (bias ? __ldg(bias + half * 32 + j) : 0.f);
			st[half * 4 + (j >> 3)] += x;
			st[8 + half * 4 + (j >> 3)] = fmaf(x, x, st[8 + half * 4 + (j >> 3)]);
		}
	}
	leaf_allreduce<16>(st, wk, exch);
#pragma unroll
	for (int g = 0; g < 8; ++g) {
		mean[g] = st[g] * (1.f / 512.f);
		const float var = fmaxf(st[8 + g] * (1.f / 512.f) - mean[g] * mean[g], 0.f);
		rstd[g] = 1.f / sqrtf(var + kGnEps);
	}
}

template <bool kProf>
__global__ void __launch_bounds__(kThreads, 1)
decode_tc_kernel(const DecoderMmaWeights w, const uint8_t* __restrict__ indices, int64_t n_leaves, float* __restrict__ voxels,
                 int tap_stage, float* __restrict__ tap_out) {
	extern __shared__ __align__(1024) uint8_t smem[];
	const uint32_t s_base = smem_u32(smem);
	const uint32_t ring = s_base + kOffRing;
	const uint32_t bars = s_base + kOffBar;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int64_t n_groups = (n_leaves + kLeavesPerCta - 1) / kLeavesPerCta;
	uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffTmemSlot);

	for (int i = threadIdx.x; i < 864; i += kThreads) reinterpret_cast<float*>(smem + kOffFinW)[i] = __ldg(w.fin_w + i);
	if (threadIdx.x == 0) {
		for (uint32_t s = 0; s < kStages; ++s) {
			mbar_init(bar_w_full(bars, s), 1);
			mbar_init(bar_w_empty(bars, s), kTiles);  // one tcgen05.commit per tile issuer
		}
		for (uint32_t t = 0; t < kTiles; ++t) {
			for (uint32_t b = 0; b < 2; ++b) {
				mbar_init(bar_a_full(bars, t, b), 4);   // one arrival per worker warp of the tile
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
		// ===================== control warps: one MMA issuer per tile (lane 0 of warp 16+t) =====================
		// Issuers are independent of each other, so tiles drift apart and one tile's epilogue overlaps the other
		// tiles' MMAs.  Lane 1 of the first control warp is the TMA producer for the ring all four tiles share.
		const uint32_t total = (uint32_t)(my_groups * kDecUnitsTotal);
		if (lane == 0) {
			const uint32_t t = warp - kWorkWarps;
			uint32_t unit = 0;
			long long tw = 0, ta = 0, ti = 0;
			const long long tstart = prof_clock<kProf>();
			for (int64_t g = 0; g < my_groups; ++g) {
#pragma unroll 1
				for (int u = 0; u < kDecUnitsTotal; ++u) {
					const uint32_t s = unit % kStages, buf = unit & 1u;
					// position of this unit inside its layer pass: stem = 54 units, then six passes of 27
					const int in_pass = u < 54 ? u : (u - 54) % 27;
					const bool last = u < 54 ? (u == 53) : (in_pass == 26);
					const long long c0 = prof_clock<kProf>();
					mbar_wait(bar_w_full(bars, s), (unit / kStages) & 1u);
					const long long c1 = prof_clock<kProf>();
					mbar_wait(bar_a_full(bars, t, buf), (unit >> 1) & 1u);
					tc_fence_after();
					const long long c2 = prof_clock<kProf>();
					const uint64_t bdesc = make_desc_sw128(ring + s * kUnitBytes);
#pragma unroll
					for (uint32_t kk = 0; kk < 4; ++kk)
						tc_mma_ts(tmem + kColD + t * 64, tmem + kColA + t * 64 + buf * 32 + kk * 8, bdesc + (uint64_t)(kk * 2),
						          (in_pass > 0 || kk > 0) ? 1u : 0u);
					tc_commit(bar_a_empty(bars, t, buf));
					if (last) tc_commit(bar_d_full(bars, t));
					tc_commit(bar_w_empty(bars, s));
					++unit;
					if (kProf) {
						tw += c1 - c0;
						ta += c2 - c1;
						ti += prof_clock<kProf>() - c2;
					}
				}
			}
			if (kProf && tap_out) {
				float* o = tap_out + ((size_t)blockIdx.x * kThreads + threadIdx.x) * 4;
				o[0] = (float)tw; o[1] = (float)ta; o[2] = (float)ti; o[3] = (float)(prof_clock<kProf>() - tstart);
			}
		} else if (lane == 1 && warp == kWorkWarps) {
#pragma unroll 1
			for (uint32_t issued = 0; issued < total; ++issued) {
				const uint32_t s = issued % kStages;
				mbar_wait(bar_w_empty(bars, s), ((issued / kStages) & 1u) ^ 1u);
				mbar_arrive_expect_tx(bar_w_full(bars, s), kUnitBytes);
				tma_load_1d(ring + s * kUnitBytes, w.units + (size_t)(issued % kDecUnitsTotal) * kUnitBytes, kUnitBytes, bar_w_full(bars, s));
			}
		}
		__syncwarp();
	} else {
		// ===================== worker warps: A staging + epilogues =====================
		Worker wk;
		wk.tile = warp >> 2;
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
		uint8_t* region = smem + kOffLeaf + wk.leaf_slot * kLeafBytes;
		const uint32_t a_base = s_base + kOffLeaf + wk.leaf_slot * kLeafBytes;  // Q [64][128] for the stem, then A [64][64]
		const uint32_t x_base = a_base + 8192;                                   // residual x [64][64] bf16; later the up_conv pass output
		float* scratch = reinterpret_cast<float*>(smem + kOffScratch) + wk.leaf_slot * kScratchFloats;
		float* exch = scratch;            // [2][32]
		float* s_mean = scratch + 64;     // [64] channel means, then [64] channel scales
		float* s_hid = scratch + 128;     // [16]
		uint32_t* s_idx = reinterpret_cast<uint32_t*>(scratch + 160);  // 16 words
		const float* s_finw = reinterpret_cast<const float*>(smem + kOffFinW);
		const int tl = wk.wil * 32 + lane;  // thread index within the leaf, 0..63 (== pos)
		long long ep[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // kProf only: cycles per epilogue section of this thread
		long long ep0 = prof_clock<kProf>();
		auto lap = [&](int slot) {
			if (kProf) {
				const long long c = prof_clock<kProf>();
				ep[slot] += c - ep0;
				ep0 = c;
			}
		};

		for (int64_t g = 0; g < my_groups; ++g) {
			const int64_t grp = blockIdx.x + g * gridDim.x;
			const int64_t leaf = grp * kLeavesPerCta + wk.leaf_slot;
			const bool leaf_ok = leaf < n_leaves;

			lap(7);
			// ---- gather: Q[pos][0..127] = codebook_bf16[idx[pos]] (64 threads per leaf; spare slots decode code 0) ----
			if (tl < 16) s_idx[tl] = leaf_ok ? __ldcs(reinterpret_cast<const uint32_t*>(indices + leaf * 64) + tl) : 0u;
			named_bar_sync(1 + wk.leaf_slot, 64);
			{
				const uint8_t* idx8 = reinterpret_cast<const uint8_t*>(s_idx);
#pragma unroll 4
				for (int i = tl; i < 64 * 16; i += 64) {
					const int pos = i >> 4, c = i & 15;
					const uint4 v = __ldg(reinterpret_cast<const uint4*>(w.emb_bf16 + (size_t)idx8[pos] * 128) + c);
					const uint32_t pc = (c & 8) | ((c & 7) ^ (pos & 7));
					*reinterpret_cast<uint4*>(region + pos * 256 + pc * 16) = v;
				}
			}
			named_bar_sync(1 + wk.leaf_slot, 64);

			float mean[8], rstd[8];
			lap(0);
			// ---- stem.0 (128->64) ; stem.1 GroupNorm + ReLU -> x ; gn1 + ReLU -> conv1 input ----
			stage_conv<2, kProf>(wk, a_base);
			wait_accumulator<kProf>(wk);  // all stem MMAs done => every read of Q is done too
			lap(6);
			gn_stats_from_tmem(wk, w.stem_b, exch, mean, rstd);
			{
				float st[16];
#pragma unroll
				for (int i = 0; i < 16; ++i) st[i] = 0.f;
#pragma unroll
				for (int half = 0; half < 2; ++half) {
					float v[32];
					tmem_ld32(wk.tmem_lane + kColD + wk.tile * 64 + half * 32, v);
#pragma unroll
					for (int j = 0; j < 32; ++j) {
						const int c = half * 32 + j, gi = c >> 3;
						const float x = fmaxf(((v[j] + __ldg(w.stem_b + c)) - mean[gi]) * rstd[gi] * __ldg(w.stem_gn_w + c) + __ldg(w.stem_gn_b + c), 0.f);
						v[j] = x;
						st[gi] += x;
						st[8 + gi] = fmaf(x, x, st[8 + gi]);
					}
					if (tap_stage == 0 && leaf_ok) {
#pragma unroll
						for (int j = 0; j < 32; ++j) tap_out[leaf * 4096 + (half * 32 + j) * 64 + wk.pos] = v[j];
					}
					store_row_half(x_base, wk.pos, half, v);  // residual x (thread-private row)
				}
				leaf_allreduce<16>(st, wk, exch);
#pragma unroll
				for (int gi = 0; gi < 8; ++gi) {
					mean[gi] = st[gi] * (1.f / 512.f);
					rstd[gi] = 1.f / sqrtf(fmaxf(st[8 + gi] * (1.f / 512.f) - mean[gi] * mean[gi], 0.f) + kGnEps);
				}
#pragma unroll
				for (int half = 0; half < 2; ++half) {
					float v[32];
					load_row_half(x_base, wk.pos, half, v);
#pragma unroll
					for (int j = 0; j < 32; ++j) {
						const int c = half * 32 + j, gi = c >> 3;
						v[j] = fmaxf((v[j] - mean[gi]) * rstd[gi] * __ldg(w.res.gn1_w + c) + __ldg(w.res.gn1_b + c), 0.f);
					}
					store_row_half(a_base, wk.pos, half, v);
				}
			}
			tc_fence_before();
			named_bar_sync(1 + wk.leaf_slot, 64);  // both warps' rows of the conv input are in place

			lap(1);
			// ---- res conv1 ; gn2 + ReLU -> conv2 input ----
			stage_conv<1, kProf>(wk, a_base);
			wait_accumulator<kProf>(wk);
			lap(6);
			gn_stats_from_tmem(wk, w.res.c1_b, exch, mean, rstd);
#pragma unroll
			for (int half = 0; half < 2; ++half) {
				float v[32];
				tmem_ld32(wk.tmem_lane + kColD + wk.tile * 64 + half * 32, v);
#pragma unroll
				for (int j = 0; j < 32; ++j) {
					const int c = half * 32 + j, gi = c >> 3;
					v[j] = fmaxf(((v[j] + __ldg(w.res.c1_b + c)) - mean[gi]) * rstd[gi] * __ldg(w.res.gn2_w + c) + __ldg(w.res.gn2_b + c), 0.f);
				}
				store_row_half(a_base, wk.pos, half, v);
			}
			tc_fence_before();
			named_bar_sync(1 + wk.leaf_slot, 64);

			lap(2);
			// ---- res conv2 ; x + 0.1 * (.) ; ChannelAttention(64) -> up_conv input ----
			stage_conv<1, kProf>(wk, a_base);
			wait_accumulator<kProf>(wk);
			lap(6);
			{
				// x' = x + 0.1 (acc + b), written back into this thread's x row
#pragma unroll
				for (int half = 0; half < 2; ++half) {
					float v[32], xr[32];
					tmem_ld32(wk.tmem_lane + kColD + wk.tile * 64 + half * 32, v);
					load_row_half(x_base, wk.pos, half, xr);
#pragma unroll
					for (int j = 0; j < 32; ++j) v[j] = xr[j] + kResScale * (v[j] + __ldg(w.res.c2_b + half * 32 + j));
					if (tap_stage == 1 && leaf_ok) {
#pragma unroll
						for (int j = 0; j < 32; ++j) tap_out[leaf * 4096 + (half * 32 + j) * 64 + wk.pos] = v[j];
					}
					store_row_half(x_base, wk.pos, half, v);
				}
				named_bar_sync(1 + wk.leaf_slot, 64);
				// channel means: thread tl sums column tl of the leaf's 64 rows
				{
					float s = 0.f;
					const uint32_t cchunk = (uint32_t)tl >> 3, cin = ((uint32_t)tl & 7u) * 2;
#pragma unroll 8
					for (int p = 0; p < 64; ++p) {
						uint16_t hv;
						asm volatile("ld.shared.u16 %0, [%1];" : "=h"(hv) : "r"(x_base + (uint32_t)p * 128 + ((cchunk ^ ((uint32_t)p & 7u)) << 4) + cin));
						s += __uint_as_float((uint32_t)hv << 16);
					}
					s_mean[tl] = s * (1.f / 64.f);
				}
				named_bar_sync(1 + wk.leaf_slot, 64);
				if (tl < 16) {
					float s = 0.f;
#pragma unroll 8
					for (int c = 0; c < 64; ++c) s = fmaf(__ldg(w.fc0 + tl * 64 + c), s_mean[c], s);
					s_hid[tl] = fmaxf(s, 0.f);
				}
				named_bar_sync(1 + wk.leaf_slot, 64);
				{
					float s = 0.f;
#pragma unroll
					for (int j = 0; j < 16; ++j) s = fmaf(__ldg(w.fc2 + tl * 16 + j), s_hid[j], s);
					named_bar_sync(1 + wk.leaf_slot, 64);  // everyone has read s_mean before it becomes the scale table
					s_mean[tl] = sigmoid_f(s);
				}
				named_bar_sync(1 + wk.leaf_slot, 64);
#pragma unroll
				for (int half = 0; half < 2; ++half) {
					float v[32];
					load_row_half(x_base, wk.pos, half, v);
#pragma unroll
					for (int j = 0; j < 32; ++j) v[j] *= s_mean[half * 32 + j];
					if (tap_stage == 2 && leaf_ok) {
#pragma unroll
						for (int j = 0; j < 32; ++j) tap_out[leaf * 4096 + (half * 32 + j) * 64 + wk.pos] = v[j];
					}
					store_row_half(a_base, wk.pos, half, v);
				}
			}
			tc_fence_before();
			named_bar_sync(1 + wk.leaf_slot, 64);

			// ---- up_conv in four 64-channel passes ; PixelShuffle3D on the store ; final conv accumulated on FFMA ----
			// thread tl owns output row R = tl of the 8^3 leaf (D = R>>3, H = R&7): 8 voxels along W.
			float out[8];
#pragma unroll
			for (int j = 0; j < 8; ++j) out[j] = 0.f;
			lap(3);
#pragma unroll 1
			for (int np = 0; np < 4; ++np) {
				stage_conv<1, kProf>(wk, a_base);
				wait_accumulator<kProf>(wk);
				lap(6);
				// channel c = np*64 + cc = oc*8 + rd*4 + rh*2 + rw  ->  oc_local = cc>>3, (rd, rh, rw) = bits of cc&7
				// P[oc_local][(2d+rd)][(2h+rh)][(2w+rw)] bf16 in the x region (8 KB per leaf)
#pragma unroll
				for (int half = 0; half < 2; ++half) {
					float v[32];
					tmem_ld32(wk.tmem_lane + kColD + wk.tile * 64 + half * 32, v);
#pragma unroll
					for (int j = 0; j < 32; j += 2) {
						const int cc = half * 32 + j, ocl = cc >> 3, rd = (cc >> 2) & 1, rh = (cc >> 1) & 1;
						const float b0 = __ldg(w.up_b + np * 64 + cc), b1 = __ldg(w.up_b + np * 64 + cc + 1);
						const uint32_t p0 = ocl * 512 + ((2 * wk.d + rd) * 8 + 2 * wk.h + rh) * 8 + 2 * wk.w;
						asm volatile("st.shared.b32 [%0], %1;" ::"r"(x_base + p0 * 2), "r"(pack_bf16(v[j] + b0, v[j + 1] + b1)));
					}
				}
				tc_fence_before();
				named_bar_sync(1 + wk.leaf_slot, 64);
				lap(4);
				{
					const int D = tl >> 3, H = tl & 7;
#pragma unroll 1
					for (int oc = 0; oc < 8; ++oc) {
						const float* wf = s_finw + (np * 8 + oc) * 27;
#pragma unroll
						for (int kd = 0; kd < 3; ++kd) {
							const int Dp = D + kd - 1;
							if ((unsigned)Dp >= 8u) continue;
#pragma unroll
							for (int kh = 0; kh < 3; ++kh) {
								const int Hp = H + kh - 1;
								if ((unsigned)Hp >= 8u) continue;
								uint4 raw;
								asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
								             : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w)
								             : "r"(x_base + (oc * 512 + (Dp * 8 + Hp) * 8) * 2));
								float xr[10];
								xr[0] = 0.f;
								xr[9] = 0.f;
								float2 f;
								f = unpack_bf16(raw.x); xr[1] = f.x; xr[2] = f.y;
								f = unpack_bf16(raw.y); xr[3] = f.x; xr[4] = f.y;
								f = unpack_bf16(raw.z); xr[5] = f.x; xr[6] = f.y;
								f = unpack_bf16(raw.w); xr[7] = f.x; xr[8] = f.y;
								const float wk0 = wf[(kd * 3 + kh) * 3], wk1 = wf[(kd * 3 + kh) * 3 + 1], wk2 = wf[(kd * 3 + kh) * 3 + 2];
#pragma unroll
								for (int j = 0; j < 8; ++j) {
									float o = out[j];
									if (j > 0) o = fmaf(xr[j], wk0, o);
									o = fmaf(xr[j + 1], wk1, o);
									if (j < 7) o = fmaf(xr[j + 2], wk2, o);
									out[j] = o;
								}
							}
						}
					}
				}
				named_bar_sync(1 + wk.leaf_slot, 64);  // all reads of this pass's P are done before the next pass overwrites it
				lap(5);
			}

			// ---- sigmoid + store: one 32-byte row segment per thread ----
			if (leaf_ok) {
				const float fb = __ldg(w.fin_b);
				float4 o0, o1;
				o0.x = sigmoid_f(out[0] + fb); o0.y = sigmoid_f(out[1] + fb); o0.z = sigmoid_f(out[2] + fb); o0.w = sigmoid_f(out[3] + fb);
				o1.x = sigmoid_f(out[4] + fb); o1.y = sigmoid_f(out[5] + fb); o1.z = sigmoid_f(out[6] + fb); o1.w = sigmoid_f(out[7] + fb);
				float4* dst = reinterpret_cast<float4*>(voxels + leaf * 512 + tl * 8);
				__stcs(dst, o0);
				__stcs(dst + 1, o1);
			}
		}
		if (kProf && tap_out) {
			float* o = tap_out + ((size_t)blockIdx.x * kThreads + threadIdx.x) * 4;
			o[0] = (float)wk.t_wait; o[1] = (float)wk.t_stage; o[2] = (float)wk.t_acc; o[3] = 0.f;
			if (threadIdx.x == 0) {  // epilogue sections of thread 0, behind the per-thread records
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
	if (e != cudaSuccess) return e;
	return cudaFuncSetAttribute(decode_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
}

cudaError_t launch_decode_tc(const DecoderMmaWeights& w, const uint8_t* dev_indices, int64_t n_leaves, float* dev_voxels,
                             int num_sms, cudaStream_t stream, int tap_stage, float* tap_out) {
	if (n_leaves <= 0) return cudaSuccess;
	const int64_t groups = (n_leaves + kLeavesPerCta - 1) / kLeavesPerCta;
	const int grid = (int)(groups < (int64_t)num_sms ? groups : (int64_t)num_sms);
	if (tap_stage == 100)  // timing instrumentation: tap_out receives 4 floats per thread (see tools/tc_pipeline_prof.py)
		decode_tc_kernel<true><<<grid, kThreads, kSmemBytes, stream>>>(w, dev_indices, n_leaves, dev_voxels, -1, tap_out);
	else
		decode_tc_kernel<false><<<grid, kThreads, kSmemBytes, stream>>>(w, dev_indices, n_leaves, dev_voxels, tap_stage, tap_out);
	return cudaGetLastError();
}

}  // namespace vqvdb
