// Back half of the tensor-core vec3 encoder (EncoderVec3, python/VQVAE_v2.py:278-299): the 4^3 stage
//     2 x ResidualBlock(128) -> ChannelAttention(128) -> proj (1x1, 128 -> 128) -> argmin over the 256 codes
// (InferenceVectorQuantizer.get_indices, python/save_for_inference.py:55-61), 128 x 64 fp32 in (the output of down1, written
// by encode_tc128_front.cu), 64 uint8 indices out.  BASELINE.json configs[3].
//
// The four 128 -> 128 convolutions are 99 % of the arithmetic and run on tcgen05 at fp32-level accuracy: operands are
// split as v = hi + lo / 2048 (two fp16 planes), three products per k-step go to two fp32 TMEM accumulators
// (hi.hi | lo.hi + hi.lo) that the epilogue combines.  The machine mapping is the vec3 decoder's (decode_tc128.cu):
//   * a GEMM tile is 128 rows = the 64 latent positions of two leaves; the tap-shifted A rows go through TMEM (TS-mode
//     MMA), copied from channels-last fp16 activation planes in shared memory by 8 stager warps;
//   * a conv is two passes of 64 output channels, a pass is 9 (kd, kh) tap pairs x 2 input-channel halves, the three kw
//     taps ride along N (N = 192) and are recombined — zero padding along w included — when the accumulator is read;
//   * weights stream through a 3 x 24 KB ring by 1-D TMA bulk copies, hi plane then lo plane per step (288 units per
//     pair of leaves, L2-resident);
//   * 8 epilogue warps own GroupNorm / residual / attention; the residual stream x stays in fp32 (global memory, the
//     leaf's own 32 KB of the input array, thread-private elements).
// proj and the distances are exact fp32 FMA chains in the oracle's order (oracle/vqvae_oracle.c conv3d / quantize),
// computed by all 16 worker warps after the last convolution; with the same z they give the same index as the fp32
// kernel (generic_model.cu) — a difference in z (summation order of the convolutions) can only move near-ties.
// Warp roles (576 threads): 0-7 epilogue (TMEM lane quadrant, channel half of the pass), 8-15 stagers (quadrant, channel
// half of the step), 16 MMA issuer (whole warp, one elected lane), 17 TMA producer.
#include "encode_tc128.cuh"
#include "leaf_ops.cuh"
#include "tc128_ops.cuh"

namespace vqvdb {

namespace {

using namespace tc128;

constexpr int kEpiWarps = 8, kStageWarps = 8, kWorkers = (kEpiWarps + kStageWarps) * 32;
constexpr int kIssuerWarp = kEpiWarps + kStageWarps, kProducerWarp = kIssuerWarp + 1;
constexpr int kThreads = (kProducerWarp + 1) * 32;  // 576
constexpr int kStages = 3;
constexpr uint32_t kUnitBytes = kEnc128UnitBytes;
constexpr int kSteps = kEnc128StepsPerPass;          // 18 per pass
constexpr int kPasses = kEnc128BackPasses;           // 8 per pair of leaves
constexpr int kConvs = kEnc128BackConvs;             // 4
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColHH = 0, kColMix = 192;        // accumulators
constexpr uint32_t kColA = 384;                      // A buffers: hi at 384 + buf*64, lo at 384 + buf*64 + 32
constexpr uint32_t kIdesc = idesc_f16(192);
// After the conv units of a pair of leaves the same ring carries the fp32 operands of proj and of the distance
// computation to the worker warps: proj.weight^T [128 c][128 d] as 4 chunks of 32 c rows, then the codebook transposed
// [128 d][256 k] as 8 chunks of 16 d rows (16 KB each).
constexpr uint32_t kVqChunkBytes = 16384;
constexpr int kProjChunks = 4, kEmbChunks = 8;
constexpr int kRingLoadsPerPair = kEnc128BackUnits + kProjChunks + kEmbChunks;  // 300

// shared memory map (bytes)
constexpr uint32_t kOffRing = 0;
constexpr uint32_t kPlaneBytes = 16384;              // [64 pos][128 ch] fp16, 256-byte rows, 16-byte chunks swizzled by pos & 7
constexpr uint32_t kBufBytes = 2 * kPlaneBytes;      // hi plane, lo plane; also one fp32 [128][64] array
constexpr uint32_t kLeafBytes = 2 * kBufBytes;       // buffers P and Q
constexpr uint32_t kOffLeaf = kOffRing + kStages * kUnitBytes;
constexpr uint32_t kOffZero = kOffLeaf + 2 * kLeafBytes;
constexpr uint32_t kOffBar = kOffZero + 256;
constexpr uint32_t kNumBars = 2 * kStages + 2 + 2 + 1 + 1 + 1;  // w_full, w_empty, a_full[2], a_empty[2], d_full, d_empty, in_ready
constexpr uint32_t kOffTmemSlot = kOffBar + kNumBars * 8;
constexpr uint32_t kOffPar = (kOffTmemSlot + 16 + 15) & ~15u;
constexpr uint32_t kOffScratch = kOffPar + par128e::total * 4;
// per-leaf scratch (floats): exch [2 slots][2 warps][2 halves][8], part [2 wil][128], scale [128], hid [32], best [16][64], best index [16][64]
constexpr uint32_t kScrExch = 0, kScrPart = 64, kScrScale = 320, kScrHid = 448, kScrBest = 480, kScrBi = 1504, kScratchFloats = 2528;
constexpr uint32_t kSmemBytes = kOffScratch + 2 * kScratchFloats * 4;
static_assert(kSmemBytes <= 227 * 1024, "encode_tc128 back smem budget");
static_assert(kOffBar % 8 == 0 && kOffPar % 16 == 0 && kOffScratch % 16 == 0, "alignment");

__device__ __forceinline__ uint32_t bar_w_full(uint32_t bars, uint32_t s) { return bars + s * 8; }
__device__ __forceinline__ uint32_t bar_w_empty(uint32_t bars, uint32_t s) { return bars + (kStages + s) * 8; }
__device__ __forceinline__ uint32_t bar_a_full(uint32_t bars, uint32_t b) { return bars + (2 * kStages + b) * 8; }
__device__ __forceinline__ uint32_t bar_a_empty(uint32_t bars, uint32_t b) { return bars + (2 * kStages + 2 + b) * 8; }
__device__ __forceinline__ uint32_t bar_d_full(uint32_t bars) { return bars + (2 * kStages + 4) * 8; }
__device__ __forceinline__ uint32_t bar_d_empty(uint32_t bars) { return bars + (2 * kStages + 5) * 8; }
__device__ __forceinline__ uint32_t bar_in_ready(uint32_t bars) { return bars + (2 * kStages + 6) * 8; }

// Physical byte offset, inside a [64 pos][128 ch] fp16 plane, of the 16-byte chunk holding channels 8*c16 .. 8*c16+7 of row pos.
__device__ __forceinline__ uint32_t chunk_off(int pos, int c16) {
	return (uint32_t)pos * 256u + ((uint32_t)((c16 & 8) | ((c16 & 7) ^ (pos & 7))) << 4);
}

// An epilogue thread: one GEMM row (latent position of one of the two leaves) x 32 of the 64 channels of a pass.
struct Epi {
	int quad, chalf, lane, row, leaf_slot, pos, w, wil;
	uint32_t bars, tmem_lane;
	uint32_t passes = 0;  // accumulator hand-overs so far
	uint32_t reds = 0;    // half-leaf reductions so far (alternates the exchange slot)
};
__device__ __forceinline__ void leaf_bar(const Epi& e) { named_bar_sync(1 + e.leaf_slot, 128); }
__device__ __forceinline__ void half_bar(const Epi& e) { named_bar_sync(3 + e.leaf_slot * 2 + e.chalf, 64); }
constexpr int kBarWorkers = 7;  // all 16 worker warps

// This thread's 32 output channels of the finished pass: hh + mix / 2048, the three kw partials combined across
// neighbouring rows.  The accumulators go back to the issuer as soon as they have been read.
__device__ __forceinline__ void take_accumulator(Epi& e, float (&v)[32]) {
	mbar_wait(bar_d_full(e.bars), e.passes & 1u);
	tc_fence_after();
	const uint32_t base = e.tmem_lane + e.chalf * 32;
	const bool has_lo = e.w > 0, has_hi = e.w < 3;
#pragma unroll
	for (int part = 0; part < 2; ++part) {
#pragma unroll
		for (int kw = 0; kw < 3; ++kw) {
			float h[16], m[16];
			tmem_ld16_nowait(base + kColHH + kw * 64 + part * 16, h);
			tmem_ld16_nowait(base + kColMix + kw * 64 + part * 16, m);
			tmem_wait_ld();
#pragma unroll
			for (int j = 0; j < 16; ++j) {
				const float t = fmaf(m[j], kLoInv, h[j]);
				if (kw == 0) {  // belongs to the row at w + 1
					const float lo = __shfl_up_sync(0xffffffffu, t, 1);
					v[part * 16 + j] = has_lo ? lo : 0.f;
				} else if (kw == 1) {
					v[part * 16 + j] += t;
				} else {  // belongs to the row at w - 1
					const float hi = __shfl_down_sync(0xffffffffu, t, 1);
					v[part * 16 + j] += has_hi ? hi : 0.f;
				}
			}
		}
	}
	tc_fence_before();
	__syncwarp();
	if (e.lane == 0) mbar_arrive(bar_d_empty(e.bars));
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

// GroupNorm(8, 128) + ReLU over the leaf for this thread's two groups of 16 channels, two-pass variance.
__device__ __forceinline__ void group_norm_relu(float (&v)[32], Epi& e, float* exch, const float* gamma, const float* beta) {
	float s[2] = {0.f, 0.f};
#pragma unroll
	for (int j = 0; j < 32; ++j) s[j >> 4] += v[j];
	half_allreduce<2>(s, e, exch);
	const float mean[2] = {s[0] * (1.f / 1024.f), s[1] * (1.f / 1024.f)};
	float q[2] = {0.f, 0.f};
#pragma unroll
	for (int j = 0; j < 32; ++j) {
		const float d = v[j] - mean[j >> 4];
		q[j >> 4] = fmaf(d, d, q[j >> 4]);
	}
	half_allreduce<2>(q, e, exch);
#pragma unroll
	for (int g = 0; g < 2; ++g) {
		const float rstd = 1.f / sqrtf(q[g] * (1.f / 1024.f) + kGnEps);
#pragma unroll
		for (int j = 0; j < 16; ++j) v[g * 16 + j] = fmaxf((v[g * 16 + j] - mean[g]) * rstd * gamma[g * 16 + j] + beta[g * 16 + j], 0.f);
	}
}

// this thread's 32 channels (first channel 8*c16_0) of row pos -> the fp16 hi and lo planes of a buffer
__device__ __forceinline__ void store_row32_split(uint32_t buf, int pos, int c16_0, const float (&v)[32]) {
#pragma unroll
	for (int q = 0; q < 4; ++q) {
		uint4 hi, lo;
		split8(v + 8 * q, hi, lo);
		const uint32_t off = chunk_off(pos, c16_0 + q);
		sts128(buf + off, hi);
		sts128(buf + kPlaneBytes + off, lo);
	}
}

// proj + distances + argmin for one leaf by its 256 worker threads.  A thread owns FOUR positions (4 pq .. 4 pq + 3, pq = t & 15)
// and one of 16 groups (grp = t >> 4) of 8 projected dims / 16 codes, so that every weight it fetches from shared memory
// feeds four FMAs (the phase is bound by shared-memory instruction issue otherwise).
// x: fp32 [128 c][64 pos] (attention output), z: fp32 [128 d][64 pos] scratch, both in shared memory.  The weights arrive
// through the ring (first_load = index of the pair's first fp32 chunk among all ring loads of this CTA).
__device__ __forceinline__ void project_and_quantize(const Encoder128BackWeights& w, const float* s_par, uint32_t bars, uint32_t ring, uint32_t first_load,
                                                     uint32_t x, uint32_t z, float* s_best, int* s_bi, int t, bool releaser, int64_t leaf, bool leaf_ok,
                                                     uint8_t* __restrict__ indices, int tap_stage, float* __restrict__ tap_out) {
	const int pq = t & 15, grp = t >> 4;
	uint32_t load = first_load;
	named_bar_sync(kBarWorkers, kWorkers);  // x is complete
	{
		// z[d][p] = b[d] + sum_c x[c][p] * W[d][c], c ascending from 0 (conv3d of the oracle with k = 1), d = 8 grp .. 8 grp + 7
		float acc[8][4];
#pragma unroll
		for (int j = 0; j < 8; ++j)
#pragma unroll
			for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
#pragma unroll 1
		for (int ch = 0; ch < kProjChunks; ++ch, ++load) {
			const uint32_t slot = load % kStages;
			mbar_wait(bar_w_full(bars, slot), (load / kStages) & 1u);
			const uint32_t wbase = ring + slot * kUnitBytes + (uint32_t)grp * 32;
#pragma unroll 4
			for (int cl = 0; cl < 32; ++cl) {
				const uint4 xr = lds128(x + (uint32_t)((ch * 32 + cl) * 64 + pq * 4) * 4);
				const float xv[4] = {__uint_as_float(xr.x), __uint_as_float(xr.y), __uint_as_float(xr.z), __uint_as_float(xr.w)};
#pragma unroll
				for (int j4 = 0; j4 < 2; ++j4) {
					const uint4 raw = lds128(wbase + (uint32_t)cl * 512 + j4 * 16);
					const float wv[4] = {__uint_as_float(raw.x), __uint_as_float(raw.y), __uint_as_float(raw.z), __uint_as_float(raw.w)};
#pragma unroll
					for (int jj = 0; jj < 4; ++jj)
#pragma unroll
						for (int i = 0; i < 4; ++i) acc[4 * j4 + jj][i] = fmaf(xv[i], wv[jj], acc[4 * j4 + jj][i]);
				}
			}
			named_bar_sync(kBarWorkers, kWorkers);  // every worker is done with the slot
			if (releaser) mbar_arrive(bar_w_empty(bars, slot));
		}
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			const int d = grp * 8 + j;
			const float b = s_par[par128e::proj_b + d];
			uint4 o;
			o.x = __float_as_uint(acc[j][0] + b);
			o.y = __float_as_uint(acc[j][1] + b);
			o.z = __float_as_uint(acc[j][2] + b);
			o.w = __float_as_uint(acc[j][3] + b);
			sts128(z + (uint32_t)(d * 64 + pq * 4) * 4, o);
			if (tap_stage == 3 && leaf_ok) {
#pragma unroll
				for (int i = 0; i < 4; ++i) tap_out[leaf * 8192 + d * 64 + pq * 4 + i] = acc[j][i] + b;
			}
		}
	}
	named_bar_sync(kBarWorkers, kWorkers);  // z is complete
	{
		// dist_k = (sum_d z_d^2 + |e_k|^2) - 2 * sum_d z_d e_kd, every sum sequential in d; first minimum wins.
		// This thread: codes 16 grp .. 16 grp + 15 for its four positions, all 64 dot products carried through the 8 chunks.
		float dot[16][4];
#pragma unroll
		for (int j = 0; j < 16; ++j)
#pragma unroll
			for (int i = 0; i < 4; ++i) dot[j][i] = 0.f;
		float zz[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
		for (int ch = 0; ch < kEmbChunks; ++ch, ++load) {
			const uint32_t slot = load % kStages;
			mbar_wait(bar_w_full(bars, slot), (load / kStages) & 1u);
			const uint32_t ebase = ring + slot * kUnitBytes + (uint32_t)grp * 64;
#pragma unroll 2
			for (int dl = 0; dl < 16; ++dl) {
				const uint4 zr = lds128(z + (uint32_t)((ch * 16 + dl) * 64 + pq * 4) * 4);
				const float zv[4] = {__uint_as_float(zr.x), __uint_as_float(zr.y), __uint_as_float(zr.z), __uint_as_float(zr.w)};
#pragma unroll
				for (int i = 0; i < 4; ++i) zz[i] = fmaf(zv[i], zv[i], zz[i]);
#pragma unroll
				for (int j4 = 0; j4 < 4; ++j4) {
					const uint4 raw = lds128(ebase + (uint32_t)dl * 1024 + j4 * 16);
					const float ev[4] = {__uint_as_float(raw.x), __uint_as_float(raw.y), __uint_as_float(raw.z), __uint_as_float(raw.w)};
#pragma unroll
					for (int jj = 0; jj < 4; ++jj)
#pragma unroll
						for (int i = 0; i < 4; ++i) dot[4 * j4 + jj][i] = fmaf(zv[i], ev[jj], dot[4 * j4 + jj][i]);
				}
			}
			named_bar_sync(kBarWorkers, kWorkers);
			if (releaser) mbar_arrive(bar_w_empty(bars, slot));
		}
		float best[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
		int bi[4] = {0, 0, 0, 0};
#pragma unroll
		for (int j = 0; j < 16; ++j) {
			const float esq = __ldg(w.emb_sq + grp * 16 + j);
#pragma unroll
			for (int i = 0; i < 4; ++i) {
				const float dist = (zz[i] + esq) - 2.f * dot[j][i];
				if (dist < best[i]) {
					best[i] = dist;
					bi[i] = grp * 16 + j;
				}
			}
		}
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			s_best[grp * 64 + pq * 4 + i] = best[i];
			s_bi[grp * 64 + pq * 4 + i] = bi[i];
		}
	}
	named_bar_sync(kBarWorkers, kWorkers);
	if (t < 64) {
		float best = s_best[t];
		int bi = s_bi[t];
#pragma unroll
		for (int gg = 1; gg < 16; ++gg)  // ascending code order, strict <: the first minimum wins
			if (s_best[gg * 64 + t] < best) {
				best = s_best[gg * 64 + t];
				bi = s_bi[gg * 64 + t];
			}
		if (leaf_ok) indices[leaf * 64 + t] = (uint8_t)bi;
	}
}

__global__ void __launch_bounds__(kThreads, 1)
encode_tc128_back_kernel(const Encoder128BackWeights w, float* __restrict__ y, int64_t n_leaves, uint8_t* __restrict__ indices, int tap_stage,
                         float* __restrict__ tap_out) {
	extern __shared__ __align__(1024) uint8_t smem[];
	const uint32_t s_base = smem_u32(smem);
	const uint32_t ring = s_base + kOffRing;
	const uint32_t bars = s_base + kOffBar;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int64_t n_groups = (n_leaves + 1) / 2;
	uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffTmemSlot);
	float* s_par = reinterpret_cast<float*>(smem + kOffPar);

	for (int i = threadIdx.x; i < par128e::total; i += kThreads) s_par[i] = __ldg(w.par + i);
	if (threadIdx.x < 64) reinterpret_cast<uint32_t*>(smem + kOffZero)[threadIdx.x] = 0u;
	if (threadIdx.x == 0) {
		for (uint32_t s = 0; s < kStages; ++s) {
			mbar_init(bar_w_full(bars, s), 1);
			mbar_init(bar_w_empty(bars, s), 1);
		}
		for (uint32_t b = 0; b < 2; ++b) {
			mbar_init(bar_a_full(bars, b), kStageWarps);
			mbar_init(bar_a_empty(bars, b), 1);
		}
		mbar_init(bar_d_full(bars), 1);
		mbar_init(bar_d_empty(bars), kEpiWarps);
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
		// ===================== TMA producer: one contiguous 24 KB unit per (pass, step, hi | lo) =====================
		if (lane == 0) {
			const uint32_t total = (uint32_t)(my_groups * kRingLoadsPerPair);
#pragma unroll 1
			for (uint32_t issued = 0; issued < total; ++issued) {
				const uint32_t s = issued % kStages, u = issued % kRingLoadsPerPair;
				const bool conv = u < (uint32_t)kEnc128BackUnits;
				const uint32_t bytes = conv ? kUnitBytes : kVqChunkBytes;
				const uint8_t* src = conv ? w.units + (size_t)u * kUnitBytes
				                          : reinterpret_cast<const uint8_t*>(w.vq_stream) + (size_t)(u - kEnc128BackUnits) * kVqChunkBytes;
				mbar_wait(bar_w_empty(bars, s), ((issued / kStages) & 1u) ^ 1u);
				mbar_arrive_expect_tx(bar_w_full(bars, s), bytes);
				tma_load_1d(ring + s * kUnitBytes, src, bytes, bar_w_full(bars, s));
			}
		}
		__syncwarp();
	} else if (warp == kIssuerWarp) {
		// ===================== MMA issuer (whole warp, one elected lane issues) =====================
		const bool leader = elect_one();
		uint32_t unit = 0, step = 0, pass = 0;
#pragma unroll 1
		for (int64_t g = 0; g < my_groups; ++g) {
#pragma unroll 1
			for (int p = 0; p < kPasses; ++p, ++pass) {
				// the epilogue warps have read the previous result out of the accumulators
				mbar_wait(bar_d_empty(bars), (pass & 1u) ^ 1u);
				tc_fence_after();
#pragma unroll 1
				for (int u = 0; u < kSteps; ++u, ++step) {
					const uint32_t ab = step & 1u;
					const uint32_t a_hi = tmem + kColA + ab * 64, a_lo = a_hi + 32;
					const uint32_t s_hi = unit % kStages, ph_hi = (unit / kStages) & 1u;
					++unit;
					const uint32_t s_lo = unit % kStages, ph_lo = (unit / kStages) & 1u;
					++unit;
					mbar_wait(bar_w_full(bars, s_hi), ph_hi);
					mbar_wait(bar_a_full(bars, ab), (step >> 1) & 1u);
					tc_fence_after();
					const uint64_t d_hi = make_desc_sw128(ring + s_hi * kUnitBytes);
					const uint32_t acc0 = u > 0 ? 1u : 0u;
#pragma unroll
					for (uint32_t kk = 0; kk < 4; ++kk)
						if (leader) tc_mma_ts(tmem + kColHH, a_hi + kk * 8, d_hi + (uint64_t)(kk * 2), kIdesc, kk > 0 ? 1u : acc0);
#pragma unroll
					for (uint32_t kk = 0; kk < 4; ++kk)
						if (leader) tc_mma_ts(tmem + kColMix, a_lo + kk * 8, d_hi + (uint64_t)(kk * 2), kIdesc, kk > 0 ? 1u : acc0);
					if (leader) tc_commit(bar_w_empty(bars, s_hi));
					mbar_wait(bar_w_full(bars, s_lo), ph_lo);
					tc_fence_after();
					const uint64_t d_lo = make_desc_sw128(ring + s_lo * kUnitBytes);
#pragma unroll
					for (uint32_t kk = 0; kk < 4; ++kk)
						if (leader) tc_mma_ts(tmem + kColMix, a_hi + kk * 8, d_lo + (uint64_t)(kk * 2), kIdesc, 1u);
					if (leader) tc_commit(bar_a_empty(bars, ab));
					if (leader) tc_commit(bar_w_empty(bars, s_lo));
					if (u == kSteps - 1 && leader) tc_commit(bar_d_full(bars));
				}
			}
			unit += kProjChunks + kEmbChunks;  // ring loads consumed by the worker warps
		}
		__syncwarp();
	} else {
		// ===================== worker warps =====================
		const bool is_stager = warp >= kEpiWarps;
		const int quad = warp & 3, chalf = (warp >> 2) & 1;
		const int row = quad * 32 + lane, leaf_slot = row >> 6, pos = row & 63;
		const uint32_t leaf_base = s_base + kOffLeaf + (uint32_t)leaf_slot * kLeafBytes;
		const uint32_t bufP = leaf_base, bufQ = leaf_base + kBufBytes;
		float* scratch = reinterpret_cast<float*>(smem + kOffScratch) + leaf_slot * kScratchFloats;
		const int t256 = (is_stager ? 128 : 0) + chalf * 64 + pos;  // thread index among the leaf's 256 workers
		const uint32_t tmem_lane = tmem + ((uint32_t)(quad * 32) << 16);

		// stager state: tap-shifted activation rows (hi and lo planes) -> TMEM A buffers
		const int pd = pos >> 4, ph = (pos >> 2) & 3;
		const uint32_t zero_row = s_base + kOffZero;
		uint32_t step = 0, layer = 0;
		// epilogue state
		Epi e;
		e.quad = quad;
		e.chalf = chalf;
		e.lane = lane;
		e.row = row;
		e.leaf_slot = leaf_slot;
		e.wil = (row >> 5) & 1;
		e.pos = pos;
		e.w = pos & 3;
		e.bars = bars;
		e.tmem_lane = tmem_lane;
		float* exch = scratch + kScrExch;
		float* s_part = scratch + kScrPart;
		float* s_scale = scratch + kScrScale;
		float* s_hid = scratch + kScrHid;
		const int tl = chalf * 64 + pos;  // thread index among the leaf's 128 epilogue threads
		auto signal_input_ready = [&]() {
			__syncwarp();
			if (lane == 0) mbar_arrive(bar_in_ready(bars));
		};

#pragma unroll 1
		for (int64_t g = 0; g < my_groups; ++g) {
			const int64_t leaf = (blockIdx.x + g * gridDim.x) * 2 + leaf_slot;
			const bool leaf_ok = leaf < n_leaves;
			const bool prof = tap_stage == 100 && threadIdx.x == 0 && blockIdx.x == 0 && g == 1;  // phase timestamps (tools/check_vec3_encode.py)
			const long long prof_t0 = prof ? clock64() : 0;
			int prof_n = 16;
			auto stamp = [&]() {
				if (prof) tap_out[prof_n++] = (float)(clock64() - prof_t0);
			};
			if (is_stager) {
#pragma unroll 1
				for (int l = 0; l < kConvs; ++l, ++layer) {
					// the layer's input (P for conv1, Q for conv2) is complete
					if (lane == 0) mbar_wait(bar_in_ready(bars), layer & 1u);
					__syncwarp();
					const uint32_t in_buf = (l & 1) ? bufQ : bufP;
#pragma unroll 1
					for (int pu = 0; pu < 2 * kSteps; ++pu, ++step) {
						const int u = pu % kSteps, t = u >> 1, khalf = u & 1;
						const int td = t / 3, th = t - td * 3;
						const bool ok = (unsigned)(pd + td - 1) < 4u && (unsigned)(ph + th - 1) < 4u;
						const int p2 = pos + (td - 1) * 16 + (th - 1) * 4;
						uint32_t rh[16], rl[16];
#pragma unroll
						for (int q = 0; q < 4; ++q) {
							const uint32_t a = ok ? in_buf + chunk_off(p2, khalf * 8 + chalf * 4 + q) : zero_row + q * 16;
							const uint4 vh = lds128(a);
							const uint4 vl = lds128(ok ? a + kPlaneBytes : a);
							rh[4 * q] = vh.x; rh[4 * q + 1] = vh.y; rh[4 * q + 2] = vh.z; rh[4 * q + 3] = vh.w;
							rl[4 * q] = vl.x; rl[4 * q + 1] = vl.y; rl[4 * q + 2] = vl.z; rl[4 * q + 3] = vl.w;
						}
						const uint32_t ab = step & 1u;
						if (lane == 0) mbar_wait(bar_a_empty(bars, ab), ((step >> 1) & 1u) ^ 1u);
						__syncwarp();
						tc_fence_after();
						tmem_st16(tmem_lane + kColA + ab * 64 + chalf * 16, rh);
						tmem_st16(tmem_lane + kColA + ab * 64 + 32 + chalf * 16, rl);
						tmem_wait_st();
						tc_fence_before();
						__syncwarp();
						if (lane == 0) mbar_arrive(bar_a_full(bars, ab));
					}
				}
			} else {
				float* xg = y + (leaf_ok ? leaf : 0) * 8192;  // the residual stream x, [128 ch][64 pos] fp32, updated in place

				float v[32];
				// ---- res_stack.0.gn1 + ReLU of the input -> P ----
#pragma unroll 1
				for (int hh = 0; hh < 2; ++hh) {
					const int c0 = hh * 64 + chalf * 32;
#pragma unroll
					for (int j = 0; j < 32; ++j) v[j] = leaf_ok ? __ldcs(xg + (c0 + j) * 64 + pos) : 0.f;
					group_norm_relu(v, e, exch, s_par + par128e::res0 + par128e::gn1_w + c0, s_par + par128e::res0 + par128e::gn1_b + c0);
					store_row32_split(bufP, pos, c0 >> 3, v);
				}
				signal_input_ready();  // the mbarrier's release/acquire orders the stores above before the stagers' loads
				stamp();

#pragma unroll 1
				for (int r = 0; r < 2; ++r) {
					const float* rp = s_par + par128e::res0 + r * par128e::res_stride;
					// conv1 -> + bias -> gn2 + ReLU -> Q
#pragma unroll 1
					for (int hh = 0; hh < 2; ++hh) {
						const int c0 = hh * 64 + chalf * 32;
						take_accumulator(e, v);
#pragma unroll
						for (int j = 0; j < 32; ++j) v[j] += rp[par128e::c1_b + c0 + j];
						group_norm_relu(v, e, exch, rp + par128e::gn2_w + c0, rp + par128e::gn2_b + c0);
						store_row32_split(bufQ, pos, c0 >> 3, v);
					}
					signal_input_ready();  // conv2's input
					stamp();
					// conv2 -> x' = x + 0.1 * (conv2 + bias)
#pragma unroll 1
					for (int hh = 0; hh < 2; ++hh) {
						const int c0 = hh * 64 + chalf * 32;
						take_accumulator(e, v);
#pragma unroll
						for (int j = 0; j < 32; ++j) {
							const float xo = leaf_ok ? xg[(c0 + j) * 64 + pos] : 0.f;
							v[j] = xo + kResScale * (v[j] + rp[par128e::c2_b + c0 + j]);
						}
						if (r == 0) {
							if (leaf_ok) {
#pragma unroll
								for (int j = 0; j < 32; ++j) xg[(c0 + j) * 64 + pos] = v[j];  // thread-private elements
							}
							if (tap_stage == 0 && leaf_ok) {
#pragma unroll
								for (int j = 0; j < 32; ++j) tap_out[leaf * 8192 + (c0 + j) * 64 + pos] = v[j];
							}
							group_norm_relu(v, e, exch, rp + par128e::res_stride + par128e::gn1_w + c0, rp + par128e::res_stride + par128e::gn1_b + c0);
							store_row32_split(bufP, pos, c0 >> 3, v);  // conv1 of this block has consumed P
						} else {
							if (tap_stage == 1 && leaf_ok) {
#pragma unroll
								for (int j = 0; j < 32; ++j) tap_out[leaf * 8192 + (c0 + j) * 64 + pos] = v[j];
							}
							// x'' as fp32 [c][pos] over P (res_stack.1.conv1 has consumed it) for the attention scale and proj
#pragma unroll
							for (int j = 0; j < 32; ++j) sts32(bufP + (uint32_t)((c0 + j) * 64 + pos) * 4, v[j]);
							const float cs = column_sums(v, lane);  // channel c0 + lane over this warp's 32 rows
							s_part[e.wil * 128 + c0 + lane] = cs;
						}
					}
					if (r == 0) signal_input_ready();  // res_stack.1.conv1's input
					stamp();
				}

				// ---- ChannelAttention(128): mean over the leaf -> 128 -> 32 -> 128 -> sigmoid ; x'' * scale in place ----
				leaf_bar(e);
				{
					const int unit = tl >> 2, part = tl & 3;
					float s = 0.f;
#pragma unroll
					for (int i = 0; i < 32; ++i) {  // fully unrolled: the 32 loads are in flight together
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
#pragma unroll
					for (int j = 0; j < 32; ++j) s = fmaf(__ldg(w.fc2 + tl * 32 + j), s_hid[j], s);
					s_scale[tl] = sigmoid_f(s);
				}
				leaf_bar(e);
#pragma unroll 1
				for (int hh = 0; hh < 2; ++hh) {
					const int c0 = hh * 64 + chalf * 32;
#pragma unroll
					for (int j = 0; j < 32; ++j) {
						const uint32_t a = bufP + (uint32_t)((c0 + j) * 64 + pos) * 4;
						const float xv = lds32(a) * s_scale[c0 + j];
						sts32(a, xv);
						if (tap_stage == 2 && leaf_ok) tap_out[leaf * 8192 + (c0 + j) * 64 + pos] = xv;
					}
				}
			}
			// every worker warp, one call site: proj + distances + argmin
			stamp();
			project_and_quantize(w, s_par, bars, ring, (uint32_t)g * kRingLoadsPerPair + kEnc128BackUnits, bufP, bufQ, scratch + kScrBest,
			                     reinterpret_cast<int*>(scratch + kScrBi), t256, threadIdx.x == 0, leaf, leaf_ok, indices, tap_stage, tap_out);
			stamp();
		}
	}

	// ---- teardown: everybody is done with TMEM before the owner frees it ----
	tc_fence_before();
	__syncthreads();
	if (warp == kIssuerWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
}

}  // namespace

cudaError_t configure_encode_tc128() {
	return cudaFuncSetAttribute(encode_tc128_back_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
}

cudaError_t launch_encode_tc128_back(const Encoder128BackWeights& w, float* dev_y, int64_t n_leaves, uint8_t* dev_indices, int num_sms,
                                     cudaStream_t stream, int tap_stage, float* tap_out) {
	if (n_leaves <= 0) return cudaSuccess;
	const int64_t groups = (n_leaves + 1) / 2;
	const int grid = (int)(groups < (int64_t)num_sms ? groups : (int64_t)num_sms);
	encode_tc128_back_kernel<<<grid, kThreads, kSmemBytes, stream>>>(w, dev_y, n_leaves, dev_indices, tap_stage, tap_out);
	return cudaGetLastError();
}

}  // namespace vqvdb
