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
// proj + codebook search: proj feeds nothing but the distances, so it is folded into the codebook on the host (M = E W,
// encode_tc128_stream.hpp) and the scores of all 256 codes come from ONE more split-fp16 GEMM on the attention output
// ([128 rows][128 c] x [128 c][256 k], SS mode: the A operand is written by the epilogue threads in the UMMA layout), with
// an error bound; a row whose two best scores are further apart than twice that bound is decided.  The others (near-ties,
// ~0.1 % of the rows) get z = W x + b and the distances of their shortlisted codes as exact fp32 FMA chains in the oracle's
// order (oracle/vqvae_oracle.c conv3d / quantize: sequential sums, first minimum wins), the index the all-fp32 path gives.
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
// After the conv units of a pair of leaves the same ring carries the folded codebook (encode_tc128_stream.hpp): for each
// input-channel half, for each half of the codes, M_hi then M_lo as [128 codes][64 c] fp16 units of 16 KB.
constexpr int kVqUnits = kEnc128VqUnits;
constexpr uint32_t kVqUnitBytes = kEnc128VqUnitBytes;
constexpr int kRingLoadsPerPair = kEnc128BackUnits + kVqUnits;  // 296
constexpr uint32_t kColVqHH = 0, kColVqMix = 256;    // score accumulators: x_hi.M_hi | x_lo.M_hi + x_hi.M_lo, 256 codes each (all of TMEM)
constexpr uint32_t kIdescVq = idesc_f16(128);
// Error bound of a score (see quantize_rows): kVqCb |x| max_k |M_k| for the split-fp16 GEMM, kVqCn (|W| |x| + |b| + max |e|)^2
// + kVqC0 for the fp32 evaluation noise of the exact formula itself (measured <= 8.6e-8 of that square on mixed fields).
constexpr float kVqCb = 1e-5f, kVqCn = 1e-6f, kVqC0 = 1e-4f;

// shared memory map (bytes)
constexpr uint32_t kOffRing = 0;
constexpr uint32_t kPlaneBytes = 16384;              // [64 pos][128 ch] fp16, 256-byte rows, 16-byte chunks swizzled by pos & 7
constexpr uint32_t kBufBytes = 2 * kPlaneBytes;      // hi plane, lo plane; also one fp32 [128][64] array
constexpr uint32_t kLeafBytes = 2 * kBufBytes;       // buffers P and Q
constexpr uint32_t kOffLeaf = kOffRing + kStages * kUnitBytes;
constexpr uint32_t kOffZero = kOffLeaf + 2 * kLeafBytes;
constexpr uint32_t kOffBar = kOffZero + 256;
constexpr uint32_t kNumBars = 2 * kStages + 2 + 2 + 1 + 1 + 1 + 1;  // w_full, w_empty, a_full[2], a_empty[2], d_full, d_empty, in_ready, x_ready
constexpr uint32_t kOffTmemSlot = kOffBar + kNumBars * 8;
constexpr uint32_t kOffPar = (kOffTmemSlot + 16 + 15) & ~15u;
constexpr uint32_t kOffScratch = kOffPar + par128e::total * 4;
// per-leaf scratch (floats): exch [2 slots][2 warps][2 halves][8], part [2 wil][128], scale [128], hid [32]
constexpr uint32_t kScrExch = 0, kScrPart = 64, kScrScale = 320, kScrHid = 448, kScratchFloats = 480;
// codebook-search scratch of the pair (floats): |x|^2 partials [2 channel halves][128 rows]; per code quarter and row the
// smallest / second-smallest score and the smallest's code [4][128] each; the re-scored minima and codes [4][128] each;
// the list of near-tie rows [128] and its length
constexpr uint32_t kVqXx = 0, kVqLo1 = 256, kVqLo2 = 768, kVqK1 = 1280, kVqBest = 1792, kVqBi = 2304, kVqAmb = 2816, kVqNamb = 2944, kVqFloats = 2948;
constexpr uint32_t kOffVq = kOffScratch + 2 * kScratchFloats * 4;
constexpr uint32_t kSmemBytes = kOffVq + kVqFloats * 4;
static_assert(kSmemBytes <= 227 * 1024, "encode_tc128 back smem budget");
static_assert(kOffBar % 8 == 0 && kOffPar % 16 == 0 && kOffScratch % 16 == 0, "alignment");

__device__ __forceinline__ uint32_t bar_w_full(uint32_t bars, uint32_t s) { return bars + s * 8; }
__device__ __forceinline__ uint32_t bar_w_empty(uint32_t bars, uint32_t s) { return bars + (kStages + s) * 8; }
__device__ __forceinline__ uint32_t bar_a_full(uint32_t bars, uint32_t b) { return bars + (2 * kStages + b) * 8; }
__device__ __forceinline__ uint32_t bar_a_empty(uint32_t bars, uint32_t b) { return bars + (2 * kStages + 2 + b) * 8; }
__device__ __forceinline__ uint32_t bar_d_full(uint32_t bars) { return bars + (2 * kStages + 4) * 8; }
__device__ __forceinline__ uint32_t bar_d_empty(uint32_t bars) { return bars + (2 * kStages + 5) * 8; }
__device__ __forceinline__ uint32_t bar_in_ready(uint32_t bars) { return bars + (2 * kStages + 6) * 8; }
__device__ __forceinline__ uint32_t bar_x_ready(uint32_t bars) { return bars + (2 * kStages + 7) * 8; }

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
		for (int j = 0; j < 16; ++j) v[g * 16 + j] = relu_f((v[g * 16 + j] - mean[g]) * rstd * gamma[g * 16 + j] + beta[g * 16 + j]);
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

// Codebook search for the pair's 128 GEMM rows by all 512 worker threads (InferenceVectorQuantizer.get_indices,
// python/save_for_inference.py:55-61: argmin_k (|z|^2 + |e_k|^2) - 2 z.e_k, first minimum wins; z = W x + b).
//  1. a_k = (|e_k|^2 - 2 b.e_k) - 2 (x_hi.M_hi + (x_lo.M_hi + x_hi.M_lo) / 2048) from the accumulators = dist_k - |z|^2 up to
//     |error| <= B = kVqCb |x| max_k |M_k| (operand split 3 * 2^-22, accumulator truncation over 8 k-steps; x 2)
//                  + kVqCn (|W| |x| + |b| + max |e|)^2 + kVqC0 (the fp32 noise of the reference formula itself).
//     A thread owns one row and a quarter of the codes (cq): one pass keeps its two smallest scores and the smallest's code.
//  2. every code that can be the fp32 arg-min has a_k <= min_j a_j + 2 B: a row whose second-smallest score exceeds that
//     is decided (~99.9 % of the rows; z is never formed).
//  3. the others: z = W x + b as an fp32 FMA chain (c ascending, bias last: conv3d of the oracle with k = 1), then every
//     code inside the window is re-scored with (sum_d z_d^2 + |e_k|^2) - 2 sum_d z_d e_kd, every sum sequential in d,
//     codes ascending, strict <: the index of the all-fp32 path.  tap_stage 3 sends every row this way (and taps z).
// x: the attention output as fp32 [128 c][64 pos] in each leaf's buffer P; z rows go over the (consumed) A operand in Q.
struct VqCtx {
	uint32_t s_base, tmem_lane;
	const float* s_par;
	float* s_vq;
	int t, row, cq, lane;
};
__device__ __forceinline__ uint32_t vq_zrow(uint32_t s_base, int r) {
	return s_base + kOffLeaf + (uint32_t)(r >> 6) * kLeafBytes + kBufBytes + (uint32_t)(r & 63) * 512u;
}
template <class Stamp>
__device__ __forceinline__ void quantize_rows(const Encoder128BackWeights& w, const VqCtx& q, int64_t pair_first_leaf, bool leaf_ok,
                                              uint8_t* __restrict__ indices, int tap_stage, float* __restrict__ tap_out, Stamp&& stamp) {
	float* s_xx = q.s_vq + kVqXx;
	float* s_lo1 = q.s_vq + kVqLo1;
	float* s_lo2 = q.s_vq + kVqLo2;
	int* s_k1 = reinterpret_cast<int*>(q.s_vq + kVqK1);
	float* s_best = q.s_vq + kVqBest;
	int* s_bi = reinterpret_cast<int*>(q.s_vq + kVqBi);
	int* s_amb = reinterpret_cast<int*>(q.s_vq + kVqAmb);
	int* s_namb = reinterpret_cast<int*>(q.s_vq + kVqNamb);
	const float* s_esq2 = q.s_par + par128e::vq_esq2;
	const int row = q.row, kb = q.cq * 64;
	const int64_t leaf = pair_first_leaf + (row >> 6);
	if (q.t == 0) *s_namb = 0;  // read again only after the barriers below

	const float xx = s_xx[row] + s_xx[128 + row];
	const float xn = sqrtf(xx);
	const float zb = fmaf(q.s_par[par128e::vq_const + 1], xn, q.s_par[par128e::vq_const + 2]);  // >= |z| + |e_k|
	const float bmax = fmaf(kVqCb * xn, q.s_par[par128e::vq_const], fmaf(kVqCn * zb, zb, kVqC0));
	float a1 = INFINITY, a2 = INFINITY;
	int k1 = 0;
#pragma unroll
	for (int b = 0; b < 4; ++b) {
		float hh[16], mx[16];
		tmem_ld16_nowait(q.tmem_lane + kColVqHH + kb + b * 16, hh);
		tmem_ld16_nowait(q.tmem_lane + kColVqMix + kb + b * 16, mx);
		tmem_wait_ld();
#pragma unroll
		for (int j = 0; j < 16; ++j) {
			const int k = kb + b * 16 + j;
			const float a = s_esq2[k] - 2.f * fmaf(mx[j], kLoInv, hh[j]);
			k1 = a < a1 ? k : k1;  // codes ascend: strict < keeps the lowest code among equal scores
			a2 = fminf(a2, fmaxf(a1, a));
			a1 = fminf(a1, a);
		}
	}
	s_lo1[q.cq * 128 + row] = a1;
	s_lo2[q.cq * 128 + row] = a2;
	s_k1[q.cq * 128 + row] = k1;
	tc_fence_before();
	named_bar_sync(kBarWorkers, kWorkers);
	float l1 = INFINITY, l2 = INFINITY;
	int kbest = 0;
#pragma unroll
	for (int o = 0; o < 4; ++o) {  // ascending code ranges
		const float b1 = s_lo1[o * 128 + row], b2 = s_lo2[o * 128 + row];
		if (b1 < l1) {
			l2 = l1;
			l1 = b1;
			kbest = s_k1[o * 128 + row];
		} else {
			l2 = fminf(l2, b1);
		}
		l2 = fminf(l2, b2);
	}
	// a non-finite row (inf / nan inputs; fminf would skip a nan score) takes the exact path; the spare slot of an odd call none
	const bool amb = leaf_ok && (!(l2 > l1 + 2.f * bmax) || !(xx < INFINITY) || tap_stage == 3);
	if (q.cq == 0) {
		if (amb) s_amb[atomicAdd(s_namb, 1)] = row;
		else if (leaf_ok) indices[leaf * 64 + (row & 63)] = (uint8_t)kbest;  // pos = (d*4+h)*4+w == view(B,4,4,4)
	}
	named_bar_sync(kBarWorkers, kWorkers);
	const int n_amb = *s_namb;
	stamp();
	if (n_amb == 0) return;  // uniform over the CTA

	// ---- z rows of the near-tie rows: thread = (one of four rows per round, projected dim d) ----
#pragma unroll 1
	for (int i0 = 0; i0 < n_amb; i0 += 4) {
		const int rs = i0 + (q.t >> 7);
		if (rs < n_amb) {
			const int r = s_amb[rs], d = q.t & 127;
			const uint32_t xb = q.s_base + kOffLeaf + (uint32_t)(r >> 6) * kLeafBytes + (uint32_t)(r & 63) * 4u;  // x[c][pos] of that leaf
			const float* wt = w.proj_t + d;
			float acc = 0.f;
#pragma unroll 1
			for (int c0 = 0; c0 < 128; c0 += 32) {  // 32 L2 loads in flight: the chain is latency-bound
				float wv[32];
#pragma unroll
				for (int c = 0; c < 32; ++c) wv[c] = __ldg(wt + (c0 + c) * 128);
#pragma unroll
				for (int c = 0; c < 32; ++c) acc = fmaf(lds32(xb + (uint32_t)(c0 + c) * 256u), wv[c], acc);
			}
			const float zv = acc + q.s_par[par128e::proj_b + d];
			sts32(vq_zrow(q.s_base, r) + (uint32_t)d * 4u, zv);
			if (tap_stage == 3) tap_out[(pair_first_leaf + (r >> 6)) * 8192 + d * 64 + (r & 63)] = zv;
		}
	}
	named_bar_sync(kBarWorkers, kWorkers);
	stamp();

	// ---- re-score this thread's codes inside the row's window (a second look at the accumulators: tcgen05.ld is
	//      warp-collective, so a warp with a near-tie row runs the loads with all its lanes) ----
	float best = INFINITY;
	int bi = 0x7fffffff;
	if (__any_sync(0xffffffffu, amb)) {
		const float near = l1 + 2.f * bmax;
		uint32_t mask[2] = {0u, 0u};
#pragma unroll
		for (int b = 0; b < 4; ++b) {
			float hh[16], mx[16];
			tmem_ld16_nowait(q.tmem_lane + kColVqHH + kb + b * 16, hh);
			tmem_ld16_nowait(q.tmem_lane + kColVqMix + kb + b * 16, mx);
			tmem_wait_ld();
#pragma unroll
			for (int j = 0; j < 16; ++j) {
				const float a = s_esq2[kb + b * 16 + j] - 2.f * fmaf(mx[j], kLoInv, hh[j]);
				if (!(a > near)) mask[b >> 1] |= 1u << ((b & 1) * 16 + j);
			}
		}
		if (amb) {
			const uint32_t zr = vq_zrow(q.s_base, row);
			float zz = 0.f;
#pragma unroll 4
			for (int d4 = 0; d4 < 32; ++d4) {
				const uint4 zq = lds128(zr + (uint32_t)d4 * 16u);
				zz = fmaf(__uint_as_float(zq.x), __uint_as_float(zq.x), zz);
				zz = fmaf(__uint_as_float(zq.y), __uint_as_float(zq.y), zz);
				zz = fmaf(__uint_as_float(zq.z), __uint_as_float(zq.z), zz);
				zz = fmaf(__uint_as_float(zq.w), __uint_as_float(zq.w), zz);
			}
#pragma unroll 1
			for (int half = 0; half < 2; ++half) {
				uint32_t m = mask[half];
				while (m) {
					const int code = kb + half * 32 + __ffs((int)m) - 1;
					m &= m - 1;
					const float4* e = reinterpret_cast<const float4*>(w.emb + (size_t)code * 128);
					float dot = 0.f;
#pragma unroll 1
					for (int d0 = 0; d0 < 32; d0 += 8) {  // 8 x 16 B of the code in flight (L2)
						float4 ev[8];
#pragma unroll
						for (int i = 0; i < 8; ++i) ev[i] = __ldg(e + d0 + i);
#pragma unroll
						for (int i = 0; i < 8; ++i) {
							const uint4 zq = lds128(zr + (uint32_t)(d0 + i) * 16u);
							dot = fmaf(__uint_as_float(zq.x), ev[i].x, dot);
							dot = fmaf(__uint_as_float(zq.y), ev[i].y, dot);
							dot = fmaf(__uint_as_float(zq.z), ev[i].z, dot);
							dot = fmaf(__uint_as_float(zq.w), ev[i].w, dot);
						}
					}
					const float dist = (zz + __ldg(w.emb_sq + code)) - 2.f * dot;
					if (dist < best) {  // codes ascend, so strict < keeps the first minimum
						best = dist;
						bi = code;
					}
				}
			}
		}
	}
	s_best[q.cq * 128 + row] = best;
	s_bi[q.cq * 128 + row] = bi;
	tc_fence_before();
	named_bar_sync(kBarWorkers, kWorkers);
	stamp();
	if (q.cq == 0 && amb) {
#pragma unroll
		for (int o = 1; o < 4; ++o) {  // ascending code ranges: strict < keeps the lowest code among equal distances
			const float ob = s_best[o * 128 + row];
			if (ob < best) {
				best = ob;
				bi = s_bi[o * 128 + row];
			}
		}
		if (bi == 0x7fffffff) bi = kbest;  // every distance was nan: keep the shortlist's choice
		indices[leaf * 64 + (row & 63)] = (uint8_t)bi;
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
		mbar_init(bar_x_ready(bars), kEpiWarps);
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
				const uint32_t bytes = conv ? kUnitBytes : kVqUnitBytes;
				const uint8_t* src = conv ? w.units + (size_t)u * kUnitBytes : w.vq_units + (size_t)(u - kEnc128BackUnits) * kVqUnitBytes;
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
			// ---- codebook scores: [128 rows][128 c] x [128 c][256 k], SS mode, three products into two accumulators ----
			mbar_wait(bar_d_empty(bars), (pass & 1u) ^ 1u);   // the last conv result has been read out of the accumulators
			mbar_wait(bar_x_ready(bars), (uint32_t)g & 1u);    // the A operand (attention output, hi / lo planes) is in shared memory
			tc_fence_after();
#pragma unroll 1
			for (uint32_t kh = 0; kh < 2; ++kh) {
				const uint64_t a_hi = make_desc_sw128(s_base + kOffLeaf + kBufBytes + kh * 16384u);
				const uint64_t a_lo = make_desc_sw128(s_base + kOffLeaf + kLeafBytes + kBufBytes + kh * 16384u);
#pragma unroll 1
				for (uint32_t nh = 0; nh < 2; ++nh) {
					const uint32_t s_hi = unit % kStages, ph_hi = (unit / kStages) & 1u;
					++unit;
					const uint32_t s_lo = unit % kStages, ph_lo = (unit / kStages) & 1u;
					++unit;
					const uint32_t d_hh = tmem + kColVqHH + nh * 128, d_mix = tmem + kColVqMix + nh * 128;
					mbar_wait(bar_w_full(bars, s_hi), ph_hi);
					tc_fence_after();
					const uint64_t b_hi = make_desc_sw128(ring + s_hi * kUnitBytes);
#pragma unroll
					for (uint32_t kk = 0; kk < 4; ++kk)
						if (leader) tc_mma_ss(d_hh, a_hi + (uint64_t)(kk * 2), b_hi + (uint64_t)(kk * 2), kIdescVq, (kh | kk) ? 1u : 0u);
#pragma unroll
					for (uint32_t kk = 0; kk < 4; ++kk)
						if (leader) tc_mma_ss(d_mix, a_lo + (uint64_t)(kk * 2), b_hi + (uint64_t)(kk * 2), kIdescVq, (kh | kk) ? 1u : 0u);
					if (leader) tc_commit(bar_w_empty(bars, s_hi));
					mbar_wait(bar_w_full(bars, s_lo), ph_lo);
					tc_fence_after();
					const uint64_t b_lo = make_desc_sw128(ring + s_lo * kUnitBytes);
#pragma unroll
					for (uint32_t kk = 0; kk < 4; ++kk)
						if (leader) tc_mma_ss(d_mix, a_hi + (uint64_t)(kk * 2), b_lo + (uint64_t)(kk * 2), kIdescVq, 1u);
					if (leader) tc_commit(bar_w_empty(bars, s_lo));
				}
			}
			if (leader) tc_commit(bar_d_full(bars));
			++pass;
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
		VqCtx vq;
		vq.s_base = s_base;
		vq.tmem_lane = tmem + ((uint32_t)(quad * 32) << 16);
		vq.s_par = s_par;
		vq.s_vq = reinterpret_cast<float*>(smem + kOffVq);
		vq.t = threadIdx.x;                       // 0..511 among the worker threads
		vq.row = row;
		vq.cq = (is_stager ? 2 : 0) + chalf;      // this thread's quarter of the codes
		vq.lane = lane;
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
					if (part == 0) s_hid[unit] = relu_f(s);
				}
				leaf_bar(e);
				stamp();
				{
					float s = 0.f;
#pragma unroll
					for (int j = 0; j < 32; ++j) s = fmaf(__ldg(w.fc2_t + j * 128 + tl), s_hid[j], s);  // [hidden][channel]: a warp reads one line
					s_scale[tl] = sigmoid_f(s);
				}
				leaf_bar(e);
				stamp();
				// x'' * scale: kept as fp32 [c][pos] in P (near-tie rows), split into the score GEMM's A operand — SWIZZLE_128B
				// K-major tiles [128 rows][64 c] over the two leaves' (consumed) Q buffers: hi planes in leaf slot 0's, lo planes
				// in leaf slot 1's, one 16 KB tile per input-channel half — and its squared norm for the error bound
				float xxp = 0.f;
#pragma unroll 1
				for (int hh = 0; hh < 2; ++hh) {
					const int c0 = hh * 64 + chalf * 32;
					float xv[32];
#pragma unroll
					for (int j = 0; j < 32; ++j) xv[j] = lds32(bufP + (uint32_t)((c0 + j) * 64 + pos) * 4);  // 32 loads in flight
#pragma unroll
					for (int j = 0; j < 32; ++j) {
						xv[j] *= s_scale[c0 + j];
						sts32(bufP + (uint32_t)((c0 + j) * 64 + pos) * 4, xv[j]);
						xxp = fmaf(xv[j], xv[j], xxp);
						if (tap_stage == 2 && leaf_ok) tap_out[leaf * 8192 + (c0 + j) * 64 + pos] = xv[j];
					}
					const uint32_t a_hi = s_base + kOffLeaf + kBufBytes + (uint32_t)hh * 16384u, a_lo = a_hi + kLeafBytes;
#pragma unroll
					for (int q = 0; q < 4; ++q) {
						uint4 hi, lo;
						split8(xv + 8 * q, hi, lo);
						const uint32_t off = (uint32_t)row * 128u + ((uint32_t)((chalf * 4 + q) ^ (row & 7)) << 4);
						sts128(a_hi + off, hi);
						sts128(a_lo + off, lo);
					}
				}
				vq.s_vq[kVqXx + chalf * 128 + row] = xxp;
				fence_proxy_async_smem();
				__syncwarp();
				if (lane == 0) mbar_arrive(bar_x_ready(bars));
				stamp();
				// the scores of the pair's 128 rows are in the accumulators
				mbar_wait(bar_d_full(bars), e.passes & 1u);
				tc_fence_after();
				tc_fence_before();
			}
			// every worker warp, one call site: scores -> indices
			stamp();
			named_bar_sync(kBarWorkers, kWorkers);  // the epilogue warps have seen the score GEMM complete; |x|^2 partials are in place
			tc_fence_after();
			quantize_rows(w, vq, (blockIdx.x + g * gridDim.x) * 2, leaf_ok, indices, tap_stage, tap_out, stamp);
			if (!is_stager) {  // every worker passed a barrier after its last look at the accumulators: hand them back
				tc_fence_before();
				__syncwarp();
				if (lane == 0) mbar_arrive(bar_d_empty(bars));
				++e.passes;
			}
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
