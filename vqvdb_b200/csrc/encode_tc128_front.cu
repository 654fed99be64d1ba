// Front half of the tensor-core vec3 encoder (EncoderVec3, python/VQVAE_v2.py:278-299): the 8^3 stage
//     pre.0 (3 -> 64) -> GroupNorm(8, 64) -> ReLU -> ResidualBlock(64) -> down1 (64 -> 128, 3x3x3, stride 2)
// 3 x 512 voxels in, 128 x 64 fp32 out (continued by encode_tc128.cu).  BASELINE.json configs[3].
//
// One leaf per CTA pass.  pre.0 is 2.5 % of the arithmetic with raw, unbounded voxel values as its operand: it stays on
// the fp32 pipes, bit-identical to the oracle's conv3d (oracle/vqvae_oracle.c) — as packed FFMA2 with the weights read
// straight from the constant bank (they travel as a kernel parameter), and, from a CTA's second leaf on, computed for
// the NEXT leaf by the epilogue warps inside the MMA phases of the current one (they have the lowest issue priority of
// the CTA's warps and are idle there for ~80 % of the time).  The two 64 -> 64 convolutions and the stride-2 conv run on
// tcgen05 in the split-fp16 scheme (encode_tc128.cuh):
//   * the conv input lives in shared memory as channels-last fp16 planes [512 pos][64 ch] (hi, lo; 128-byte rows,
//     16-byte chunks XOR-swizzled by pos & 7); a GEMM tile is 128 positions (two d slices); the tap-shifted rows go
//     through TMEM (TS-mode MMA), copied by 8 stager warps;
//   * a 64 -> 64 conv is 4 tiles x 9 (kd, kh) steps with the kw taps along N (N = 192), recombined when the
//     accumulators are read; the stride-2 conv is one tile (64 output positions, rows 64..127 idle) x 27 taps, N = 128,
//     its hi.hi products alternating between two accumulators and drained after 14 of the 27 taps (the tensor core
//     truncates its fp32 accumulator after every MMA: four chains of <= 7 taps keep that bias at the level of the
//     other layers);
//   * GroupNorm needs the whole leaf: conv outputs and the residual stream go through a per-CTA fp32 scratch in global
//     memory (L2-resident, every element private to one thread between barriers) and are normalised / split into the
//     planes once all four tiles are done.
// Warp roles (576 threads): 0-7 epilogue, 8-15 stagers, 16 MMA issuer, 17 TMA producer; a CTA's first pre.0 is computed
// by all 16 worker warps.
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
constexpr uint32_t kSlotBytes = kEnc128UnitBytes;    // ring slot: one 24 KB conv unit or one 16 KB down unit
constexpr uint32_t kDownUnitBytes = 16384;           // [128 n][64 k] fp16
constexpr int kConvSteps = 9, kTiles = 4, kDownSteps = 27;
constexpr int kDownSplit = 14;                       // down1's taps are accumulated in two halves (0..13, 14..26), drained separately
constexpr int kPassesPerLeaf = 2 * kTiles + 2;       // accumulator hand-overs per leaf
constexpr int kUnitsPerLeaf = 2 * kTiles * kConvSteps * 2 + kDownSteps * 2;  // 198 ring loads per leaf
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColHH = 0, kColMix = 192;        // 64 -> 64 convs: N = 192
constexpr uint32_t kColDownHH = 0, kColDownMix = 256;  // down: two hh chains of N = 128, then mix
constexpr uint32_t kColA = 384;                      // A buffers: hi at 384 + buf*64, lo at + 32
constexpr uint32_t kIdescConv = idesc_f16(192), kIdescDown = idesc_f16(128);

// shared memory map (bytes)
constexpr uint32_t kOffRing = 0;
constexpr uint32_t kPlaneBytes = 512 * 128;          // [512 pos][64 ch] fp16
constexpr uint32_t kOffPlanes = kOffRing + kStages * kSlotBytes;  // hi plane, lo plane
constexpr uint32_t kOffIn = kOffPlanes + 2 * kPlaneBytes;         // the input leaf with a zero halo: [3][10][10][10] fp32
constexpr uint32_t kInFloats = 3000;
constexpr uint32_t kOffZero = kOffIn + kInFloats * 4;
constexpr uint32_t kOffBar = kOffZero + 128;
constexpr uint32_t kNumBars = 2 * kStages + 2 + 2 + 1 + 1 + 1;
constexpr uint32_t kOffTmemSlot = kOffBar + kNumBars * 8;
constexpr uint32_t kOffPar = (kOffTmemSlot + 16 + 15) & ~15u;
constexpr uint32_t kOffRed = kOffPar + par128f::total * 4;        // [2 slots][16 warps][8] block reductions, then [2 slots][8 warps][4]
constexpr uint32_t kSmemBytes = kOffRed + (2 * 16 * 8 + 2 * 8 * 4) * 4;
static_assert(kSmemBytes <= 227 * 1024, "encode_tc128 front smem budget");
static_assert(kOffBar % 8 == 0 && kOffPar % 16 == 0 && kOffIn % 16 == 0, "alignment");

__device__ __forceinline__ uint32_t bar_w_full(uint32_t bars, uint32_t s) { return bars + s * 8; }
__device__ __forceinline__ uint32_t bar_w_empty(uint32_t bars, uint32_t s) { return bars + (kStages + s) * 8; }
__device__ __forceinline__ uint32_t bar_a_full(uint32_t bars, uint32_t b) { return bars + (2 * kStages + b) * 8; }
__device__ __forceinline__ uint32_t bar_a_empty(uint32_t bars, uint32_t b) { return bars + (2 * kStages + 2 + b) * 8; }
__device__ __forceinline__ uint32_t bar_d_full(uint32_t bars) { return bars + (2 * kStages + 4) * 8; }
__device__ __forceinline__ uint32_t bar_d_empty(uint32_t bars) { return bars + (2 * kStages + 5) * 8; }
__device__ __forceinline__ uint32_t bar_in_ready(uint32_t bars) { return bars + (2 * kStages + 6) * 8; }
constexpr int kBarWorkers = 1;   // all 16 worker warps
constexpr int kBarEpiHalf = 2;   // + chalf: the 4 epilogue warps of one channel half
constexpr int kBarEpi = 4;       // the 8 epilogue warps

// byte offset of the 16-byte chunk with channels 8*c8 .. 8*c8+7 of row pos inside a [512][64] fp16 plane
__device__ __forceinline__ uint32_t chunk_off(int pos, int c8) { return (uint32_t)pos * 128u + ((uint32_t)(c8 ^ (pos & 7)) << 4); }

// Sum of 8 per-thread values over the 512 worker threads (every one of them calls this).
__device__ __forceinline__ void workers_allreduce8(float (&v)[8], float* red, uint32_t& count, int warp, int lane) {
#pragma unroll
	for (int i = 0; i < 8; ++i) v[i] = warp_sum(v[i]);
	float* x = red + (count & 1u) * 128;
	if (lane == 0) {
#pragma unroll
		for (int i = 0; i < 8; ++i) x[warp * 8 + i] = v[i];
	}
	named_bar_sync(kBarWorkers, kWorkers);
	float s = 0.f;  // lane i (mod 8) adds up column i, in the same order in every warp
#pragma unroll
	for (int wi = 0; wi < 16; ++wi) s += x[wi * 8 + (lane & 7)];
#pragma unroll
	for (int i = 0; i < 8; ++i) v[i] = __shfl_sync(0xffffffffu, s, i);
	++count;  // the other slot next time: this one is rewritten only after one more barrier has been passed
}

// Sum of 4 per-thread values over the 128 epilogue threads of one channel half (4 warps).
__device__ __forceinline__ void epi_allreduce4(float (&v)[4], float* red, uint32_t& count, int quad, int chalf, int lane) {
#pragma unroll
	for (int i = 0; i < 4; ++i) v[i] = warp_sum(v[i]);
	float* x = red + 256 + (count & 1u) * 32;
	if (lane == 0) {
#pragma unroll
		for (int i = 0; i < 4; ++i) x[(quad * 2 + chalf) * 4 + i] = v[i];
	}
	named_bar_sync(kBarEpiHalf + chalf, 128);
#pragma unroll
	for (int i = 0; i < 4; ++i) v[i] = x[chalf * 4 + i] + x[(2 + chalf) * 4 + i] + x[(4 + chalf) * 4 + i] + x[(6 + chalf) * 4 + i];
	++count;
}

// pre.0 for four w-adjacent positions x eight output channels c0 .. c0 + 7 of the leaf staged (with a zero halo) at s_in:
// out[c] = b[c] + sum_{ic, kd, kh, kw} in * w, ascending, as conv3d of the oracle (halo taps add an exact 0).  Packed fp32
// pairs (FFMA2, each half an IEEE fma) over channel pairs; every input value feeds up to three taps, every weight — read
// from the constant bank — four positions.  pq = the thread's position quad: (d, h, half of the w row).
// Input channels [ic0, ic1): a call with ic0 > 0 continues the partial sums a previous call left in xs (same chain of
// FMAs, same order); the call that ends with the last input channel adds the bias.
__device__ __forceinline__ void pre0_chunk(const Encoder128PreWeights& pw, uint32_t s_in_addr, const float* s_par, float* __restrict__ xs, int pq, int c0,
                                           int ic0, int ic1) {
	const int pd = pq >> 4, ph = (pq >> 1) & 7, w0 = (pq & 1) * 4;
	const uint32_t in0 = s_in_addr + (uint32_t)(pd * 100 + ph * 10 + w0) * 4;
	const int pos0 = pd * 64 + ph * 8 + w0;
	uint64_t acc2[4][4];  // [channel pair][position]
#pragma unroll
	for (int jp = 0; jp < 4; ++jp) {
		float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
		if (ic0 > 0) {
			lo = *reinterpret_cast<const float4*>(xs + (c0 + 2 * jp) * 512 + pos0);
			hi = *reinterpret_cast<const float4*>(xs + (c0 + 2 * jp + 1) * 512 + pos0);
		}
		acc2[jp][0] = pack_f32x2(lo.x, hi.x);
		acc2[jp][1] = pack_f32x2(lo.y, hi.y);
		acc2[jp][2] = pack_f32x2(lo.z, hi.z);
		acc2[jp][3] = pack_f32x2(lo.w, hi.w);
	}
#pragma unroll 1
	for (int ic = ic0; ic < ic1; ++ic) {
#pragma unroll 1
		for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
			for (int kh = 0; kh < 3; ++kh) {
				const uint32_t ia = in0 + (uint32_t)(ic * 1000 + kd * 100 + kh * 10) * 4;
				uint64_t in2[6];  // the input value in both halves
#pragma unroll
				for (int i2 = 0; i2 < 3; ++i2) {
					float2 v2;
					asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v2.x), "=f"(v2.y) : "r"(ia + i2 * 8));
					in2[2 * i2] = pack_f32x2(v2.x, v2.x);
					in2[2 * i2 + 1] = pack_f32x2(v2.y, v2.y);
				}
#pragma unroll
				for (int kw = 0; kw < 3; ++kw) {
					const float4* wp = reinterpret_cast<const float4*>(pw.w + (ic * 27 + (kd * 3 + kh) * 3 + kw) * 64 + c0);
					const float4 wa = wp[0], wb = wp[1];
					const uint64_t w2[4] = {pack_f32x2(wa.x, wa.y), pack_f32x2(wa.z, wa.w), pack_f32x2(wb.x, wb.y), pack_f32x2(wb.z, wb.w)};
#pragma unroll
					for (int jp = 0; jp < 4; ++jp)
#pragma unroll
						for (int i = 0; i < 4; ++i) acc2[jp][i] = fma_f32x2(in2[i + kw], w2[jp], acc2[jp][i]);
				}
			}
		}
	}
#pragma unroll
	for (int jp = 0; jp < 4; ++jp) {
		float lo[4], hi[4];
#pragma unroll
		for (int i = 0; i < 4; ++i) unpack_f32x2(acc2[jp][i], lo[i], hi[i]);
		if (ic1 == 3) {  // (a partial sum is stored as it is: adding a zero would turn a -0 into +0)
			const float b0 = s_par[par128f::pre_b + c0 + 2 * jp], b1 = s_par[par128f::pre_b + c0 + 2 * jp + 1];
#pragma unroll
			for (int i = 0; i < 4; ++i) {
				lo[i] += b0;
				hi[i] += b1;
			}
		}
		*reinterpret_cast<float4*>(xs + (c0 + 2 * jp) * 512 + pos0) = make_float4(lo[0], lo[1], lo[2], lo[3]);
		*reinterpret_cast<float4*>(xs + (c0 + 2 * jp + 1) * 512 + pos0) = make_float4(hi[0], hi[1], hi[2], hi[3]);
	}
}
// The leaf's 3 x 512 voxels into the haloed staging buffer, by `nthreads` threads (t = 0 .. nthreads - 1).
__device__ __forceinline__ void stage_leaf(float* s_in, const float* __restrict__ src, int t, int nthreads) {
	for (int e = t; e < 1536; e += nthreads) {
		const int c = e >> 9, p = e & 511;
		s_in[c * 1000 + ((p >> 6) + 1) * 100 + (((p >> 3) & 7) + 1) * 10 + (p & 7) + 1] = __ldcs(src + e);
	}
}

__global__ void __launch_bounds__(kThreads, 1)
encode_tc128_front_kernel(const Encoder128FrontWeights w, const __grid_constant__ Encoder128PreWeights pw, const float* __restrict__ leaves,
                          int64_t n_leaves, float* __restrict__ y, float* __restrict__ scratch, int tap_stage, float* __restrict__ tap_out) {
	extern __shared__ __align__(1024) uint8_t smem[];
	const uint32_t s_base = smem_u32(smem);
	const uint32_t ring = s_base + kOffRing;
	const uint32_t bars = s_base + kOffBar;
	const uint32_t plane_hi = s_base + kOffPlanes, plane_lo = plane_hi + kPlaneBytes;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffTmemSlot);
	float* s_par = reinterpret_cast<float*>(smem + kOffPar);
	float* s_in = reinterpret_cast<float*>(smem + kOffIn);
	float* s_red = reinterpret_cast<float*>(smem + kOffRed);

	for (int i = threadIdx.x; i < par128f::total; i += kThreads) s_par[i] = __ldg(w.par + i);
	for (int i = threadIdx.x; i < (int)kInFloats; i += kThreads) s_in[i] = 0.f;  // the halo stays zero
	if (threadIdx.x < 32) reinterpret_cast<uint32_t*>(smem + kOffZero)[threadIdx.x] = 0u;
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
	const int64_t my_leaves = blockIdx.x < n_leaves ? (n_leaves - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

	if (warp == kProducerWarp) {
		// ===================== TMA producer =====================
		// per leaf: conv1's 18 units for each of the 4 tiles, conv2's likewise, then the 54 units of down
		if (lane == 0) {
			uint32_t issued = 0;
#pragma unroll 1
			for (int64_t g = 0; g < my_leaves; ++g) {
#pragma unroll 1
				for (int i = 0; i < kUnitsPerLeaf; ++i, ++issued) {
					const uint32_t s = issued % kStages;
					uint32_t u, bytes = kSlotBytes;
					if (i < 2 * kTiles * 18) u = (uint32_t)(i / (kTiles * 18)) * 18 + (uint32_t)(i % 18);
					else {
						u = 36 + (uint32_t)(i - 2 * kTiles * 18);
						bytes = kDownUnitBytes;
					}
					mbar_wait(bar_w_empty(bars, s), ((issued / kStages) & 1u) ^ 1u);
					mbar_arrive_expect_tx(bar_w_full(bars, s), bytes);
					tma_load_1d(ring + s * kSlotBytes, w.units + (size_t)u * kSlotBytes, bytes, bar_w_full(bars, s));
				}
			}
		}
		__syncwarp();
	} else if (warp == kIssuerWarp) {
		// ===================== MMA issuer (whole warp, one elected lane issues) =====================
		const bool leader = elect_one();
		uint32_t unit = 0, step = 0, pass = 0;
#pragma unroll 1
		for (int64_t g = 0; g < my_leaves; ++g) {
#pragma unroll 1
			for (int p = 0; p < kPassesPerLeaf; ++p, ++pass) {
				const bool down = p >= 2 * kTiles;
				const int steps = !down ? kConvSteps : p == 2 * kTiles ? kDownSplit : kDownSteps - kDownSplit;
				const uint32_t idesc = down ? kIdescDown : kIdescConv;
				mbar_wait(bar_d_empty(bars), (pass & 1u) ^ 1u);
				tc_fence_after();
#pragma unroll 1
				for (int u = 0; u < steps; ++u, ++step) {
					const uint32_t ab = step & 1u;
					const uint32_t a_hi = tmem + kColA + ab * 64, a_lo = a_hi + 32;
					const uint32_t s_hi = unit % kStages, ph_hi = (unit / kStages) & 1u;
					++unit;
					const uint32_t s_lo = unit % kStages, ph_lo = (unit / kStages) & 1u;
					++unit;
					// down: the hi.hi products alternate between two accumulators
					const uint32_t d_hh = tmem + (down ? kColDownHH + (uint32_t)(u & 1) * 128 : kColHH);
					const uint32_t d_mix = tmem + (down ? kColDownMix : kColMix);
					const uint32_t acc_hh = (down ? u > 1 : u > 0) ? 1u : 0u, acc_mix = u > 0 ? 1u : 0u;
					mbar_wait(bar_w_full(bars, s_hi), ph_hi);
					mbar_wait(bar_a_full(bars, ab), (step >> 1) & 1u);
					tc_fence_after();
					const uint64_t b_hi = make_desc_sw128(ring + s_hi * kSlotBytes);
#pragma unroll
					for (uint32_t kk = 0; kk < 4; ++kk)
						if (leader) tc_mma_ts(d_hh, a_hi + kk * 8, b_hi + (uint64_t)(kk * 2), idesc, kk > 0 ? 1u : acc_hh);
#pragma unroll
					for (uint32_t kk = 0; kk < 4; ++kk)
						if (leader) tc_mma_ts(d_mix, a_lo + kk * 8, b_hi + (uint64_t)(kk * 2), idesc, kk > 0 ? 1u : acc_mix);
					if (leader) tc_commit(bar_w_empty(bars, s_hi));
					mbar_wait(bar_w_full(bars, s_lo), ph_lo);
					tc_fence_after();
					const uint64_t b_lo = make_desc_sw128(ring + s_lo * kSlotBytes);
#pragma unroll
					for (uint32_t kk = 0; kk < 4; ++kk)
						if (leader) tc_mma_ts(d_mix, a_hi + kk * 8, b_lo + (uint64_t)(kk * 2), idesc, 1u);
					if (leader) tc_commit(bar_a_empty(bars, ab));
					if (leader) tc_commit(bar_w_empty(bars, s_lo));
					if (u == steps - 1 && leader) tc_commit(bar_d_full(bars));
				}
			}
		}
		__syncwarp();
	} else {
		// ===================== worker warps =====================
		const bool is_stager = warp >= kEpiWarps;
		const int quad = warp & 3, chalf = (warp >> 2) & 1;
		const int row = quad * 32 + lane;
		const int wt = threadIdx.x;  // 0..511: this thread's position in the pre phase
		const uint32_t tmem_lane = tmem + ((uint32_t)(quad * 32) << 16);
		const uint32_t zero_row = s_base + kOffZero;
		// per-CTA scratch: the residual stream x [64 ch][512 pos] fp32 of the current leaf and of the next one (whose pre.0
		// is computed ahead), and conv1's output before GroupNorm
		float* xs = scratch + (size_t)blockIdx.x * (3 * 64 * 512);
		float* xs_next = xs + 64 * 512;
		float* cs = xs + 2 * 64 * 512;
		const uint32_t s_in_addr = s_base + kOffIn;
		uint32_t n_red = 0, n_ered = 0, step = 0, layer = 0, passes = 0;

#pragma unroll 1
		for (int64_t g = 0; g < my_leaves; ++g) {
			const int64_t leaf = blockIdx.x + g * gridDim.x;
			const bool prof = tap_stage == 100 && threadIdx.x == 0 && blockIdx.x == 0 && g == 1;  // phase timestamps (tools/check_vec3_encode.py)
			const long long prof_t0 = prof ? clock64() : 0;
			int prof_n = 0;
			auto stamp = [&]() {
				if (prof) tap_out[prof_n++] = (float)(clock64() - prof_t0);
			};

			// ---------- pre phase, all workers: pre.0 + GroupNorm + ReLU -> x ; res.gn1 + ReLU -> planes ----------
			const bool has_next = g + 1 < my_leaves;
			{
				if (g == 0) {
					// the CTA's first leaf: pre.0 by all 512 workers, two chunks of 8 channels each (later leaves: see the epilogue warps)
					stage_leaf(s_in, leaves + leaf * 1536, wt, kWorkers);
					named_bar_sync(kBarWorkers, kWorkers);
#pragma unroll 1
					for (int r = 0; r < 2; ++r) pre0_chunk(pw, s_in_addr, s_par, xs, wt & 127, ((wt >> 7) * 2 + r) * 8, 0, 3);
				}
				named_bar_sync(kBarWorkers, kWorkers);  // the conv output is complete in the scratch; nobody reads the previous leaf's planes any more
				stamp();
				float v[64], gsum[8];
#pragma unroll
				for (int c = 0; c < 64; ++c) v[c] = xs[c * 512 + wt];
#pragma unroll
				for (int gI = 0; gI < 8; ++gI) {
					float t = 0.f;
#pragma unroll
					for (int j = 0; j < 8; ++j) t += v[gI * 8 + j];
					gsum[gI] = t;  // GroupNorm(8, 64): group = 8 consecutive channels
				}
				// pre.1 GroupNorm + ReLU (two-pass variance over 8 channels x 512 positions)
				workers_allreduce8(gsum, s_red, n_red, warp, lane);
				float mean[8], q[8];
#pragma unroll
				for (int gI = 0; gI < 8; ++gI) {
					mean[gI] = gsum[gI] * (1.f / 4096.f);
					float t = 0.f;
#pragma unroll
					for (int j = 0; j < 8; ++j) {
						const float d = v[gI * 8 + j] - mean[gI];
						t = fmaf(d, d, t);
					}
					q[gI] = t;
				}
				workers_allreduce8(q, s_red, n_red, warp, lane);
#pragma unroll
				for (int gI = 0; gI < 8; ++gI) {
					const float rstd = 1.f / sqrtf(q[gI] * (1.f / 4096.f) + kGnEps);
					float t = 0.f;
#pragma unroll
					for (int j = 0; j < 8; ++j) {
						const int c = gI * 8 + j;
						v[c] = relu_f((v[c] - mean[gI]) * rstd * s_par[par128f::pre_gn_w + c] + s_par[par128f::pre_gn_b + c]);
						xs[c * 512 + wt] = v[c];
						t += v[c];
					}
					gsum[gI] = t;
				}
				if (tap_stage == 0) {
#pragma unroll
					for (int c = 0; c < 64; ++c) tap_out[leaf * 32768 + c * 512 + wt] = v[c];
				}
				// res.gn1 + ReLU -> conv1's input
				workers_allreduce8(gsum, s_red, n_red, warp, lane);
#pragma unroll
				for (int gI = 0; gI < 8; ++gI) {
					mean[gI] = gsum[gI] * (1.f / 4096.f);
					float t = 0.f;
#pragma unroll
					for (int j = 0; j < 8; ++j) {
						const float d = v[gI * 8 + j] - mean[gI];
						t = fmaf(d, d, t);
					}
					q[gI] = t;
				}
				workers_allreduce8(q, s_red, n_red, warp, lane);
#pragma unroll
				for (int gI = 0; gI < 8; ++gI) {
					const float rstd = 1.f / sqrtf(q[gI] * (1.f / 4096.f) + kGnEps);
					float a[8];
#pragma unroll
					for (int j = 0; j < 8; ++j) {
						const int c = gI * 8 + j;
						a[j] = relu_f((v[c] - mean[gI]) * rstd * s_par[par128f::gn1_w + c] + s_par[par128f::gn1_b + c]);
					}
					uint4 hi, lo;
					split8(a, hi, lo);
					const uint32_t off = chunk_off(wt, gI);
					sts128(plane_hi + off, hi);
					sts128(plane_lo + off, lo);
				}
				named_bar_sync(kBarWorkers, kWorkers);  // planes and x complete (x is read by other threads from here on)
				stamp();
			}

			if (is_stager) {
				// ---------- stagers: tap-shifted rows of the planes -> TMEM A buffers ----------
#pragma unroll 1
				for (int l = 0; l < 3; ++l, ++layer) {
					if (lane == 0) mbar_wait(bar_in_ready(bars), layer & 1u);
					__syncwarp();
					const int n_steps = l < 2 ? kTiles * kConvSteps : kDownSteps;
#pragma unroll 1
					for (int su = 0; su < n_steps; ++su, ++step) {
						bool ok;
						int p2;
						if (l < 2) {
							const int tile = su / kConvSteps, t = su - tile * kConvSteps, td = t / 3, th = t - td * 3;
							const int pos = tile * 128 + row, pd = pos >> 6, ph = (pos >> 3) & 7;
							ok = (unsigned)(pd + td - 1) < 8u && (unsigned)(ph + th - 1) < 8u;
							p2 = pos + (td - 1) * 64 + (th - 1) * 8;
						} else {
							const int td = su / 9, th = (su / 3) % 3, tw = su % 3;
							const int id = 2 * ((row >> 4) & 3) - 1 + td, ih = 2 * ((row >> 2) & 3) - 1 + th, iw = 2 * (row & 3) - 1 + tw;
							ok = row < 64 && (unsigned)id < 8u && (unsigned)ih < 8u && (unsigned)iw < 8u;
							p2 = id * 64 + ih * 8 + iw;
						}
						uint32_t rh[16], rl[16];
#pragma unroll
						for (int q = 0; q < 4; ++q) {
							const uint32_t off = ok ? chunk_off(p2, chalf * 4 + q) : 0u;
							const uint4 vh = lds128(ok ? plane_hi + off : zero_row + q * 16);
							const uint4 vl = lds128(ok ? plane_lo + off : zero_row + q * 16);
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
				// ---------- epilogue warps ----------
				const int w8 = row & 7;  // w coordinate of this thread's row in every tile
				const bool has_lo = w8 > 0, has_hi = w8 < 7;
				const int c0 = chalf * 32;
				auto signal_input_ready = [&]() {
					__syncwarp();
					if (lane == 0) mbar_arrive(bar_in_ready(bars));
				};
				// this thread's 32 output channels of a finished N = 192 tile: hh + mix / 2048, kw partials combined
				auto take_conv_tile = [&](float (&v)[32]) {
					mbar_wait(bar_d_full(bars), passes & 1u);
					tc_fence_after();
					const uint32_t base = tmem_lane + c0;
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
								if (kw == 0) {
									const float lo = __shfl_up_sync(0xffffffffu, t, 1);
									v[part * 16 + j] = has_lo ? lo : 0.f;
								} else if (kw == 1) {
									v[part * 16 + j] += t;
								} else {
									const float hi = __shfl_down_sync(0xffffffffu, t, 1);
									v[part * 16 + j] += has_hi ? hi : 0.f;
								}
							}
						}
					}
					tc_fence_before();
					__syncwarp();
					if (lane == 0) mbar_arrive(bar_d_empty(bars));
					++passes;
				};
				signal_input_ready();  // layer 0: the planes were completed before the workers' barrier above
				// The next leaf's pre.0 runs in the shadow of this leaf's MMAs: its voxels go into the staging buffer now (this
				// leaf's pre.0 is long done with it) and its 12 sub-chunks (four channel-slice pairs x three input channels, ~4 k
				// cycles each) follow the epilogues of conv1's and conv2's tiles 0..2 and fill the two halves of down1 (three each), partial
				// sums parked in the other x buffer.  (Measured, cycles per leaf: serial pre.0 199 k; whole 81-tap chunks after the
				// tile epilogues or under down1 188 k — they delay the next accumulator hand-over; this placement 172 k; two
				// sub-chunks per tile epilogue and one per half of down1 185 k.)
				if (has_next) {
					stage_leaf(s_in, leaves + (leaf + gridDim.x) * 1536, wt, kEpiWarps * 32);
					named_bar_sync(kBarEpi, kEpiWarps * 32);
				}
				// sub-chunk sc = 0 .. 11 of the next leaf's pre.0: channel slice pair sc / 3 (8 channels per 128 threads), input channel sc % 3
				auto next_pre0 = [&](int sc) {
					if (has_next) pre0_chunk(pw, s_in_addr, s_par, xs_next, wt & 127, ((sc / 3) * 2 + (wt >> 7)) * 8, sc % 3, sc % 3 + 1);
				};

				float v[32];
				// ---- conv1 + bias -> scratch, group sums ; after the 4 tiles: gn2 + ReLU -> planes ----
				float gs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
				for (int tile = 0; tile < kTiles; ++tile) {
					take_conv_tile(v);
					const int pos = tile * 128 + row;
#pragma unroll
					for (int j = 0; j < 32; ++j) {
						v[j] += s_par[par128f::c1_b + c0 + j];
						cs[(c0 + j) * 512 + pos] = v[j];
						gs[j >> 3] += v[j];
					}
					if (tile < 3) next_pre0(tile);  // the last tile's epilogue is followed by the GroupNorm sweeps: nothing in their way
				}
				stamp();
				epi_allreduce4(gs, s_red, n_ered, quad, chalf, lane);
				float mean[4], q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
				for (int i = 0; i < 4; ++i) mean[i] = gs[i] * (1.f / 4096.f);
#pragma unroll 1
				for (int tile = 0; tile < kTiles; ++tile) {
					const int pos = tile * 128 + row;
#pragma unroll
					for (int j = 0; j < 32; ++j) {
						const float d = cs[(c0 + j) * 512 + pos] - mean[j >> 3];
						q[j >> 3] = fmaf(d, d, q[j >> 3]);
					}
				}
				epi_allreduce4(q, s_red, n_ered, quad, chalf, lane);
#pragma unroll
				for (int i = 0; i < 4; ++i) q[i] = 1.f / sqrtf(q[i] * (1.f / 4096.f) + kGnEps);
#pragma unroll 1
				for (int tile = 0; tile < kTiles; ++tile) {
					const int pos = tile * 128 + row;
					float cv[32];  // the tile's 32 loads in flight together (the shared-memory stores below are compiler barriers)
#pragma unroll
					for (int j = 0; j < 32; ++j) cv[j] = cs[(c0 + j) * 512 + pos];
#pragma unroll
					for (int c8 = 0; c8 < 4; ++c8) {
						float a[8];
#pragma unroll
						for (int j = 0; j < 8; ++j) {
							const int c = c0 + c8 * 8 + j;
							a[j] = relu_f((cv[c8 * 8 + j] - mean[c8]) * q[c8] * s_par[par128f::gn2_w + c] + s_par[par128f::gn2_b + c]);
						}
						uint4 hi, lo;
						split8(a, hi, lo);
						const uint32_t off = chunk_off(pos, chalf * 4 + c8);
						sts128(plane_hi + off, hi);
						sts128(plane_lo + off, lo);
					}
				}
				signal_input_ready();  // layer 1: conv2's input
				stamp();

				// ---- conv2: x' = x + 0.1 * (conv2 + bias), in place in the scratch ; after the 4 tiles: x' split -> planes ----
#pragma unroll 1
				for (int tile = 0; tile < kTiles; ++tile) {
					take_conv_tile(v);
					const int pos = tile * 128 + row;
					float xo[32];  // all 32 loads before the first store (the compiler cannot prove the stores do not alias them)
#pragma unroll
					for (int j = 0; j < 32; ++j) xo[j] = xs[(c0 + j) * 512 + pos];
#pragma unroll
					for (int j = 0; j < 32; ++j) {
						const float xn = xo[j] + kResScale * (v[j] + s_par[par128f::c2_b + c0 + j]);
						xs[(c0 + j) * 512 + pos] = xn;
						if (tap_stage == 1) tap_out[leaf * 32768 + (c0 + j) * 512 + pos] = xn;
					}
					if (tile < 3) next_pre0(3 + tile);
				}
				stamp();
#pragma unroll 1
				for (int tile = 0; tile < kTiles; ++tile) {
					const int pos = tile * 128 + row;
					float xv[32];  // as above: all loads of the tile before its first shared-memory store
#pragma unroll
					for (int j = 0; j < 32; ++j) xv[j] = xs[(c0 + j) * 512 + pos];
#pragma unroll
					for (int c8 = 0; c8 < 4; ++c8) {
						const float* a = xv + c8 * 8;
						uint4 hi, lo;
						split8(a, hi, lo);
						const uint32_t off = chunk_off(pos, chalf * 4 + c8);
						sts128(plane_hi + off, hi);
						sts128(plane_lo + off, lo);
					}
				}
				signal_input_ready();  // layer 2: down1's input
				stamp();

				// ---- down1: rows 0..63 are the 4^3 output positions; this thread: output channels 64 chalf .. + 63.  The taps
				//      arrive as two separately drained halves (four accumulation chains of <= 7 taps in all) ----
#pragma unroll 1
				for (int half = 0; half < 2; ++half) {
					// three more sub-chunks of the next leaf's pre.0 under each half of down1's MMAs (~15 k cycles each)
#pragma unroll 1
					for (int k = 0; k < 3; ++k) next_pre0(6 + half * 3 + k);
					mbar_wait(bar_d_full(bars), passes & 1u);
					tc_fence_after();
					if (row < 64) {
#pragma unroll
						for (int part = 0; part < 4; ++part) {
							float h0[16], h1[16], m[16];
							const uint32_t col = tmem_lane + chalf * 64 + part * 16;
							tmem_ld16_nowait(col + kColDownHH, h0);
							tmem_ld16_nowait(col + kColDownHH + 128, h1);
							tmem_ld16_nowait(col + kColDownMix, m);
							tmem_wait_ld();
#pragma unroll
							for (int j = 0; j < 16; ++j) {
								const int c = chalf * 64 + part * 16 + j;
								float* dst = y + leaf * 8192 + c * 64 + row;
								const float part_sum = fmaf(m[j], kLoInv, h0[j] + h1[j]);
								*dst = half == 0 ? part_sum : (*dst + part_sum) + s_par[par128f::down_b + c];  // thread-private element
							}
						}
					}
					tc_fence_before();
					__syncwarp();
					if (lane == 0) mbar_arrive(bar_d_empty(bars));
					++passes;
				}
				stamp();
			}
			{  // the next leaf's x (pre.0 output) was written into the other buffer
				float* t = xs;
				xs = xs_next;
				xs_next = t;
			}
		}
	}

	// ---- teardown: everybody is done with TMEM before the owner frees it ----
	tc_fence_before();
	__syncthreads();
	if (warp == kIssuerWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
}

}  // namespace

cudaError_t configure_encode_tc128_front() {
	return cudaFuncSetAttribute(encode_tc128_front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
}

size_t encode_tc128_front_scratch_floats(int num_sms) { return (size_t)num_sms * 3 * 64 * 512; }

cudaError_t launch_encode_tc128_front(const Encoder128FrontWeights& w, const Encoder128PreWeights& pw, const float* dev_leaves, int64_t n_leaves,
                                      float* dev_y, float* dev_scratch, int num_sms, cudaStream_t stream, int tap_stage, float* tap_out) {
	if (n_leaves <= 0) return cudaSuccess;
	const int grid = (int)(n_leaves < (int64_t)num_sms ? n_leaves : (int64_t)num_sms);
	encode_tc128_front_kernel<<<grid, kThreads, kSmemBytes, stream>>>(w, pw, dev_leaves, n_leaves, dev_y, dev_scratch, tap_stage, tap_out);
	return cudaGetLastError();
}

}  // namespace vqvdb
