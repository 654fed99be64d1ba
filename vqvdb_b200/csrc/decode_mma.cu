// Tensor-core decoder, warp-level MMA (bf16 operands, fp32 accumulate): 64 index bytes in, 512 voxels out.
//
// Same dataflow as decode_fp32.cu (reference: python/save_for_inference.py:91-104 +
// python/VQVAE_v2.py:253-275), but the four big convolutions (stem 128->64, the two 64->64 of the
// residual block, up_conv 64->256: 56.6 of the decoder's 57.1 M MAC/leaf) run as implicit GEMMs on the
// tensor cores.  SURVEY §7.4(3): bf16 conv operands change PSNR(x, recon) by < 0.005 dB.
//
// Work decomposition
//   * a CTA holds 8 leaves; consumer warp w owns leaf w END TO END: its 64 latent positions are the
//     M = 64 rows of every GEMM, so GroupNorm / channel-attention reductions are warp shuffles and no
//     CTA-wide barrier exists on the data path.
//   * im2col is never materialised: an ldmatrix row address IS the gather.  Activations live in shared
//     memory channels-last ([64 pos][C] bf16, 16-byte chunks XOR-swizzled by pos&7); for a tap the lane
//     points ldmatrix at the shifted position's row, or at a zero chunk when the tap falls outside the leaf.
//   * weights never fit in shared memory (1.77 MB bf16), so they stream from L2 as 216 "units" of
//     [64 n][64 k] bf16 (8 KB, pre-swizzled on the host) through an 8-stage ring filled
//     with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx) issued by one elected lane;
//     all 8 leaves share each unit.
//   * PixelShuffle3D is a store-address permutation; the final 32->1 conv (0.44 M MAC) runs on FFMA,
//     accumulated over the four 64-channel passes of up_conv so only 8 KB of its input is ever live.
#include <cuda_bf16.h>

#include "decode_mma.cuh"
#include "leaf_ops.cuh"
#include "ptx_utils.cuh"

namespace vqvdb {

namespace {

constexpr int kLeavesPerCta = 8;
constexpr int kThreads = kLeavesPerCta * 32;  // 8 warps = 2 per SM sub-partition, so each may use up to 255 registers;
                                              // lane 0 of warp 0 doubles as the TMA producer (a 9th warp would cap regs at 168)
constexpr int kStages = 8;
constexpr uint32_t kUnitBytes = 8192;

// shared memory map (bytes)
constexpr uint32_t kOffRing = 0;                                    // kStages x 8 KB, 1 KB aligned
constexpr uint32_t kOffLeaf = kOffRing + kStages * kUnitBytes;      // 8 x 16 KB per-leaf regions
constexpr uint32_t kLeafBytes = 16384;
constexpr uint32_t kOffZero = kOffLeaf + kLeavesPerCta * kLeafBytes;  // 128 B of zeros
constexpr uint32_t kOffBar = kOffZero + 128;                        // full[kStages], empty[kStages]
constexpr uint32_t kOffFinW = kOffBar + 2 * kStages * 8;            // final conv weights, 864 floats
constexpr uint32_t kOffScratch = kOffFinW + 864 * 4;                // per warp: 64 means + 16 hidden + 16 idx words
constexpr uint32_t kScratchBytes = 96 * 4;
constexpr uint32_t kSmemBytes = kOffScratch + kLeavesPerCta * kScratchBytes;
static_assert(kSmemBytes <= 227 * 1024, "decode_mma smem budget");

// Position in the weight-unit stream.  Every warp consumes the same sequence; `unit` counts units since
// kernel start, so stage = unit % kStages and the mbarrier parity = (unit / kStages) & 1.
struct Pipe {
	uint32_t unit = 0;      // next unit this warp consumes
	uint32_t issued = 0;    // producer only (warp 0): units whose TMA copy has been issued
	uint32_t total = 0;     // producer only: units this CTA will consume in the whole launch
	const uint8_t* src = nullptr;
	uint32_t ring = 0, bars = 0;
	bool producer = false;  // warp 0, lane 0
	__device__ __forceinline__ uint32_t stage() const { return unit % kStages; }
	__device__ __forceinline__ uint32_t phase() const { return (unit / kStages) & 1u; }
	// Producer duty, run before consuming unit `unit`: that unit MUST be in flight (blocking on its stage's
	// empty barrier if necessary); up to kStages-1 further units are issued only if their stage is already free.
	__device__ __forceinline__ void produce() {
		if (!producer) return;
		while (issued < total && issued < unit + kStages) {
			const uint32_t s = issued % kStages, par = ((issued / kStages) & 1u) ^ 1u;
			const uint32_t empty = bars + (kStages + s) * 8, full = bars + s * 8;
			if (issued <= unit) mbar_wait(empty, par);
			else if (!mbar_test(empty, par)) break;
			mbar_arrive_expect_tx(full, kUnitBytes);
			tma_load_1d(ring + s * kUnitBytes, src + (size_t)(issued % kDecUnitsTotal) * kUnitBytes, kUnitBytes, full);
			++issued;
		}
	}
};

// One 3x3x3 convolution over the warp's leaf as an implicit GEMM: acc[64 pos][64 n] += A[64 pos][27*CIN] W.
// UNITS_PER_TAP = CIN / 64; activation rows are ROW_BYTES = CIN*2 bytes, swizzled per 128-byte half.
template <int UNITS_PER_TAP>
__device__ __forceinline__ void conv_mma(float (&acc)[4][8][4], uint32_t a_base, uint32_t zero_addr, uint32_t ring,
                                         uint32_t bars, Pipe& pipe, int lane) {
	constexpr uint32_t ROW_BYTES = UNITS_PER_TAP * 128;
#pragma unroll
	for (int mt = 0; mt < 4; ++mt)
#pragma unroll
		for (int nt = 0; nt < 8; ++nt)
#pragma unroll
			for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;

	const int r = lane & 15, h = r >> 2, w = r & 3;
	const uint32_t khalf = lane >> 4;  // which 8-wide k chunk of the 16-wide step this lane addresses
	// B-fragment addressing: lane -> row n = (lane>>4)*8 + (lane&7) of a 16-row pair of n-tiles, chunk parity (lane>>3)&1
	const uint32_t bn = ((lane >> 4) << 3) + (lane & 7);
	const uint32_t bpar = (lane >> 3) & 1;

#pragma unroll 1
	for (int tap = 0; tap < 27; ++tap) {
		const int td = tap / 9, th = (tap / 3) % 3, tw = tap % 3;
		const bool hw_ok = (unsigned)(h + th - 1) < 4u && (unsigned)(w + tw - 1) < 4u;
		const int shift = (td - 1) * 16 + (th - 1) * 4 + (tw - 1);
		uint32_t row_addr[4], swz[4];
		bool ok[4];
#pragma unroll
		for (int mt = 0; mt < 4; ++mt) {
			const int pos = mt * 16 + r + shift;
			ok[mt] = hw_ok && (unsigned)(mt + td - 1) < 4u;
			row_addr[mt] = a_base + (uint32_t)pos * ROW_BYTES;
			swz[mt] = (uint32_t)pos & 7u;
		}
#pragma unroll
		for (int u = 0; u < UNITS_PER_TAP; ++u) {
			pipe.produce();
			mbar_wait(bars + pipe.stage() * 8, pipe.phase());
			const uint32_t wbase = ring + pipe.stage() * kUnitBytes;
#pragma unroll
			for (int kk = 0; kk < 4; ++kk) {
				uint32_t a[4][4];
				const uint32_t chunk = kk * 2 + khalf;
#pragma unroll
				for (int mt = 0; mt < 4; ++mt) {
					const uint32_t addr = ok[mt] ? row_addr[mt] + u * 128 + ((chunk ^ swz[mt]) << 4) : zero_addr;
					ldmatrix_x4(addr, a[mt][0], a[mt][1], a[mt][2], a[mt][3]);
				}
#pragma unroll
				for (int j = 0; j < 4; ++j) {
					const uint32_t n = j * 16 + bn;
					const uint32_t bchunk = kk * 2 + bpar;
					uint32_t b0, b1, b2, b3;
					ldmatrix_x4(wbase + n * 128 + ((bchunk ^ (n & 7u)) << 4), b0, b1, b2, b3);
#pragma unroll
					for (int mt = 0; mt < 4; ++mt) {
						mma_bf16(acc[mt][2 * j], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b0, b1);
						mma_bf16(acc[mt][2 * j + 1], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b2, b3);
					}
				}
			}
			__syncwarp();
			if (lane == 0) mbar_arrive(bars + (kStages + pipe.stage()) * 8);
			++pipe.unit;
		}
	}
}

// Consumes the unit stream without computing (warps whose leaf slot is past the end of the batch).
__device__ __forceinline__ void skip_units(int n_units, uint32_t bars, Pipe& pipe, int lane) {
	for (int i = 0; i < n_units; ++i) {
		pipe.produce();
		mbar_wait(bars + pipe.stage() * 8, pipe.phase());
		__syncwarp();
		if (lane == 0) mbar_arrive(bars + (kStages + pipe.stage()) * 8);
		++pipe.unit;
	}
}

// acc += bias (per output channel), C-fragment layout: e=0,1 -> ch nt*8+2t+{0,1} at row g ; e=2,3 same ch at row g+8
__device__ __forceinline__ void add_bias(float (&acc)[4][8][4], const float* __restrict__ bias, int t) {
#pragma unroll
	for (int nt = 0; nt < 8; ++nt) {
		const float2 b = __ldg(reinterpret_cast<const float2*>(bias + nt * 8 + 2 * t));
#pragma unroll
		for (int mt = 0; mt < 4; ++mt) {
			acc[mt][nt][0] += b.x;
			acc[mt][nt][1] += b.y;
			acc[mt][nt][2] += b.x;
			acc[mt][nt][3] += b.y;
		}
	}
}

// GroupNorm(8, 64) + ReLU in registers.  Group nt = the 8 channels of n-tile nt over all 64 rows (512 values
// spread over the 32 lanes, 16 per lane).  Two-pass variance, fp32.
__device__ __forceinline__ void group_norm_relu(float (&v)[4][8][4], const float* __restrict__ gamma,
                                                const float* __restrict__ beta, int t) {
#pragma unroll
	for (int nt = 0; nt < 8; ++nt) {
		float s = 0.f;
#pragma unroll
		for (int mt = 0; mt < 4; ++mt) s += (v[mt][nt][0] + v[mt][nt][1]) + (v[mt][nt][2] + v[mt][nt][3]);
		const float mean = warp_sum(s) * (1.f / 512.f);
		float q = 0.f;
#pragma unroll
		for (int mt = 0; mt < 4; ++mt)
#pragma unroll
			for (int e = 0; e < 4; ++e) {
				const float d = v[mt][nt][e] - mean;
				q = fmaf(d, d, q);
			}
		const float rstd = 1.f / sqrtf(warp_sum(q) * (1.f / 512.f) + kGnEps);
		const float2 ga = __ldg(reinterpret_cast<const float2*>(gamma + nt * 8 + 2 * t));
		const float2 be = __ldg(reinterpret_cast<const float2*>(beta + nt * 8 + 2 * t));
#pragma unroll
		for (int mt = 0; mt < 4; ++mt) {
			v[mt][nt][0] = fmaxf((v[mt][nt][0] - mean) * rstd * ga.x + be.x, 0.f);
			v[mt][nt][1] = fmaxf((v[mt][nt][1] - mean) * rstd * ga.y + be.y, 0.f);
			v[mt][nt][2] = fmaxf((v[mt][nt][2] - mean) * rstd * ga.x + be.x, 0.f);
			v[mt][nt][3] = fmaxf((v[mt][nt][3] - mean) * rstd * ga.y + be.y, 0.f);
		}
	}
}

// Registers (C-fragment layout) -> the warp's [64 pos][64 ch] bf16 conv-input buffer (128-byte rows, swizzled).
__device__ __forceinline__ void store_conv_input(const float (&v)[4][8][4], uint32_t a_base, int g, int t) {
#pragma unroll
	for (int mt = 0; mt < 4; ++mt) {
		const uint32_t row0 = mt * 16 + g;  // row0 + 8 has the same (row & 7)
#pragma unroll
		for (int nt = 0; nt < 8; ++nt) {
			const uint32_t off = (((uint32_t)nt ^ (row0 & 7u)) << 4) + t * 4;
			asm volatile("st.shared.b32 [%0], %1;" ::"r"(a_base + row0 * 128 + off), "r"(pack_bf16(v[mt][nt][0], v[mt][nt][1])));
			asm volatile("st.shared.b32 [%0], %1;" ::"r"(a_base + (row0 + 8) * 128 + off), "r"(pack_bf16(v[mt][nt][2], v[mt][nt][3])));
		}
	}
}

__device__ __forceinline__ void dump_tap(const float (&v)[4][8][4], float* dst, int g, int t) {
	// dst: [64 ch][64 pos] fp32 for this leaf (the C oracle's tap layout)
#pragma unroll
	for (int mt = 0; mt < 4; ++mt)
#pragma unroll
		for (int nt = 0; nt < 8; ++nt) {
			const int ch = nt * 8 + 2 * t, row = mt * 16 + g;
			dst[ch * 64 + row] = v[mt][nt][0];
			dst[(ch + 1) * 64 + row] = v[mt][nt][1];
			dst[ch * 64 + row + 8] = v[mt][nt][2];
			dst[(ch + 1) * 64 + row + 8] = v[mt][nt][3];
		}
}

__global__ void __launch_bounds__(kThreads, 1)
decode_mma_kernel(const DecoderMmaWeights w, const uint8_t* __restrict__ indices, int64_t n_leaves,
                  float* __restrict__ voxels, int tap_stage, float* __restrict__ tap_out) {
	extern __shared__ __align__(1024) uint8_t smem[];
	const uint32_t s_base = smem_u32(smem);
	const uint32_t ring = s_base + kOffRing;
	const uint32_t bars = s_base + kOffBar;
	const uint32_t zero_addr = s_base + kOffZero;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int64_t n_groups = (n_leaves + kLeavesPerCta - 1) / kLeavesPerCta;

	if (threadIdx.x < 32) reinterpret_cast<uint32_t*>(smem + kOffZero)[threadIdx.x] = 0u;
	for (int i = threadIdx.x; i < 864; i += kThreads) reinterpret_cast<float*>(smem + kOffFinW)[i] = __ldg(w.fin_w + i);
	if (threadIdx.x == 0) {
		for (int s = 0; s < kStages; ++s) {
			mbar_init(bars + s * 8, 1);                          // full: one expect_tx arrival by the producer
			mbar_init(bars + (kStages + s) * 8, kLeavesPerCta);  // empty: one arrival per consumer warp
		}
		mbar_fence_init();
	}
	__syncthreads();

	// ===== consumer warps: one leaf each =====
	const int g = lane >> 2, t = lane & 3;
	uint8_t* region = smem + kOffLeaf + warp * kLeafBytes;
	const uint32_t a_base = s_base + kOffLeaf + warp * kLeafBytes;  // conv input [64][64] bf16 (or Q [64][128] for the stem)
	const uint32_t x_base = a_base + 8192;                           // residual (thread-private words), later up_conv pass output
	float* scratch = reinterpret_cast<float*>(smem + kOffScratch + warp * kScratchBytes);
	const float* s_finw = reinterpret_cast<const float*>(smem + kOffFinW);
	Pipe pipe;
	pipe.ring = ring;
	pipe.bars = bars;
	pipe.src = w.units;
	pipe.producer = (threadIdx.x == 0);
	{
		const int64_t my_groups = blockIdx.x < n_groups ? (n_groups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
		pipe.total = (uint32_t)(my_groups * kDecUnitsTotal);
	}
	float acc[4][8][4];

	for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
		const int64_t leaf = grp * kLeavesPerCta + warp;
		if (leaf >= n_leaves) {  // ragged tail: keep the barrier protocol alive
			skip_units(kDecUnitsTotal, bars, pipe, lane);
			continue;
		}

		// ---- gather: Q[pos][0..127] = codebook_bf16[idx[pos]]  (F.embedding + permute, save_for_inference.py:91-101) ----
		{
			uint32_t* s_idx = reinterpret_cast<uint32_t*>(scratch + 80);
			if (lane < 16) s_idx[lane] = __ldcs(reinterpret_cast<const uint32_t*>(indices + leaf * 64) + lane);
			__syncwarp();
			const uint8_t* idx8 = reinterpret_cast<const uint8_t*>(s_idx);
#pragma unroll 4
			for (int i = lane; i < 64 * 16; i += 32) {
				const int pos = i >> 4, c = i & 15;  // 16-byte chunk c of the 256-byte row
				const uint4 v = __ldg(reinterpret_cast<const uint4*>(w.emb_bf16 + (size_t)idx8[pos] * 128) + c);
				const uint32_t pc = (c & 8) | ((c & 7) ^ (pos & 7));
				*reinterpret_cast<uint4*>(region + pos * 256 + pc * 16) = v;
			}
			__syncwarp();
		}

		// ---- stem.0 (128->64) + stem.1 GroupNorm + ReLU -> x ----
		conv_mma<2>(acc, a_base, zero_addr, ring, bars, pipe, lane);
		__syncwarp();  // every lane is done reading Q before the region is reused
		add_bias(acc, w.stem_b, t);
		group_norm_relu(acc, w.stem_gn_w, w.stem_gn_b, t);
		if (tap_stage == 0) dump_tap(acc, tap_out + leaf * 4096, g, t);
		// residual x -> thread-private bf16 words
#pragma unroll
		for (int mt = 0; mt < 4; ++mt)
#pragma unroll
			for (int nt = 0; nt < 8; ++nt) {
				const int j = (mt * 8 + nt) * 2;
				asm volatile("st.shared.b32 [%0], %1;" ::"r"(x_base + (j * 32 + lane) * 4), "r"(pack_bf16(acc[mt][nt][0], acc[mt][nt][1])));
				asm volatile("st.shared.b32 [%0], %1;" ::"r"(x_base + ((j + 1) * 32 + lane) * 4), "r"(pack_bf16(acc[mt][nt][2], acc[mt][nt][3])));
			}

		// ---- ResidualBlock(64): x + 0.1 * conv2(relu(gn2(conv1(relu(gn1(x)))))) ----
		group_norm_relu(acc, w.res.gn1_w, w.res.gn1_b, t);
		store_conv_input(acc, a_base, g, t);
		__syncwarp();
		conv_mma<1>(acc, a_base, zero_addr, ring, bars, pipe, lane);
		__syncwarp();
		add_bias(acc, w.res.c1_b, t);
		group_norm_relu(acc, w.res.gn2_w, w.res.gn2_b, t);
		store_conv_input(acc, a_base, g, t);
		__syncwarp();
		conv_mma<1>(acc, a_base, zero_addr, ring, bars, pipe, lane);
		__syncwarp();
		add_bias(acc, w.res.c2_b, t);
#pragma unroll
		for (int mt = 0; mt < 4; ++mt)
#pragma unroll
			for (int nt = 0; nt < 8; ++nt) {
				const int j = (mt * 8 + nt) * 2;
				uint32_t w0, w1;
				asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w0) : "r"(x_base + (j * 32 + lane) * 4));
				asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w1) : "r"(x_base + ((j + 1) * 32 + lane) * 4));
				const float2 x0 = unpack_bf16(w0), x1 = unpack_bf16(w1);
				acc[mt][nt][0] = x0.x + kResScale * acc[mt][nt][0];
				acc[mt][nt][1] = x0.y + kResScale * acc[mt][nt][1];
				acc[mt][nt][2] = x1.x + kResScale * acc[mt][nt][2];
				acc[mt][nt][3] = x1.y + kResScale * acc[mt][nt][3];
			}
		if (tap_stage == 1) dump_tap(acc, tap_out + leaf * 4096, g, t);

		// ---- ChannelAttention(64): x *= sigmoid(W2 relu(W1 mean(x))) ----
		{
#pragma unroll
			for (int nt = 0; nt < 8; ++nt) {
				float s0 = 0.f, s1 = 0.f;
#pragma unroll
				for (int mt = 0; mt < 4; ++mt) {
					s0 += acc[mt][nt][0] + acc[mt][nt][2];
					s1 += acc[mt][nt][1] + acc[mt][nt][3];
				}
#pragma unroll
				for (int o = 4; o < 32; o <<= 1) {
					s0 += __shfl_xor_sync(0xffffffffu, s0, o);
					s1 += __shfl_xor_sync(0xffffffffu, s1, o);
				}
				if (g == 0) {
					scratch[nt * 8 + 2 * t] = s0 * (1.f / 64.f);
					scratch[nt * 8 + 2 * t + 1] = s1 * (1.f / 64.f);
				}
			}
			__syncwarp();
			{
				const int j = lane & 15;
				float s = 0.f;
#pragma unroll 8
				for (int c = 0; c < 64; ++c) s = fmaf(__ldg(w.fc0 + j * 64 + c), scratch[c], s);
				__syncwarp();
				if (lane < 16) scratch[64 + j] = fmaxf(s, 0.f);
			}
			__syncwarp();
#pragma unroll
			for (int nt = 0; nt < 8; ++nt) {
				float y[2];
#pragma unroll
				for (int e = 0; e < 2; ++e) {
					const int ch = nt * 8 + 2 * t + e;
					float s = 0.f;
#pragma unroll
					for (int j = 0; j < 16; ++j) s = fmaf(__ldg(w.fc2 + ch * 16 + j), scratch[64 + j], s);
					y[e] = sigmoid_f(s);
				}
#pragma unroll
				for (int mt = 0; mt < 4; ++mt) {
					acc[mt][nt][0] *= y[0];
					acc[mt][nt][1] *= y[1];
					acc[mt][nt][2] *= y[0];
					acc[mt][nt][3] *= y[1];
				}
			}
			__syncwarp();
		}
		if (tap_stage == 2) dump_tap(acc, tap_out + leaf * 4096, g, t);
		store_conv_input(acc, a_base, g, t);
		__syncwarp();

		// ---- up_conv (64->256) in four 64-channel passes, PixelShuffle3D on the store, final conv accumulated ----
		// Each lane owns two rows of the 8^3 output: R0 = lane and R1 = lane + 32, R = D*8 + H, 8 voxels along W each.
		float out[2][8];
#pragma unroll
		for (int rr = 0; rr < 2; ++rr)
#pragma unroll
			for (int j = 0; j < 8; ++j) out[rr][j] = 0.f;

#pragma unroll 1
		for (int np = 0; np < 4; ++np) {
			conv_mma<1>(acc, a_base, zero_addr, ring, bars, pipe, lane);
			add_bias(acc, w.up_b + np * 64, t);
			// channel c = np*64 + nt*8 + 2t + e = oc*8 + rd*4 + rh*2 + rw  ->  oc = np*8 + nt, rd = t>>1, rh = t&1, rw = e
			// P[oc_local = nt][(2d+rd)][(2h+rh)][(2w+rw)] bf16, d = mt, (h, w) from the fragment row.
			{
				const int rd = t >> 1, rh = t & 1;
				const int h0 = g >> 2, w0 = g & 3;
#pragma unroll
				for (int mt = 0; mt < 4; ++mt)
#pragma unroll
					for (int nt = 0; nt < 8; ++nt) {
						const uint32_t p0 = nt * 512 + ((2 * mt + rd) * 8 + 2 * h0 + rh) * 8 + 2 * w0;
						asm volatile("st.shared.b32 [%0], %1;" ::"r"(x_base + p0 * 2), "r"(pack_bf16(acc[mt][nt][0], acc[mt][nt][1])));
						asm volatile("st.shared.b32 [%0], %1;" ::"r"(x_base + (p0 + 32) * 2), "r"(pack_bf16(acc[mt][nt][2], acc[mt][nt][3])));
					}
			}
			__syncwarp();
			// final conv, partial over the 8 input channels of this pass (decoder.final, 32->1 k3 p1)
#pragma unroll
			for (int rr = 0; rr < 2; ++rr) {
				const int R = lane + rr * 32, D = R >> 3, H = R & 7;
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
							for (int j = 0; j < 8; ++j) {  // xr[0] and xr[9] are the exact-zero w halo: their taps are skipped
								float o = out[rr][j];
								if (j > 0) o = fmaf(xr[j], wk0, o);
								o = fmaf(xr[j + 1], wk1, o);
								if (j < 7) o = fmaf(xr[j + 2], wk2, o);
								out[rr][j] = o;
							}
						}
					}
				}
			}
			__syncwarp();  // all lanes finished reading this pass's P before the next pass overwrites it
		}

		// ---- sigmoid + store: each lane writes two 32-byte row segments ----
		{
			const float fb = __ldg(w.fin_b);
#pragma unroll
			for (int rr = 0; rr < 2; ++rr) {
				const int R = lane + rr * 32;
				float4 o0, o1;
				o0.x = sigmoid_f(out[rr][0] + fb); o0.y = sigmoid_f(out[rr][1] + fb);
				o0.z = sigmoid_f(out[rr][2] + fb); o0.w = sigmoid_f(out[rr][3] + fb);
				o1.x = sigmoid_f(out[rr][4] + fb); o1.y = sigmoid_f(out[rr][5] + fb);
				o1.z = sigmoid_f(out[rr][6] + fb); o1.w = sigmoid_f(out[rr][7] + fb);
				float4* dst = reinterpret_cast<float4*>(voxels + leaf * 512 + R * 8);
				__stcs(dst, o0);
				__stcs(dst + 1, o1);
			}
		}
		__syncwarp();
	}
}

}  // namespace

cudaError_t configure_decode_mma() {
	return cudaFuncSetAttribute(decode_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
}

cudaError_t launch_decode_mma(const DecoderMmaWeights& w, const uint8_t* dev_indices, int64_t n_leaves,
                              float* dev_voxels, int num_sms, cudaStream_t stream, int tap_stage, float* tap_out) {
	if (n_leaves <= 0) return cudaSuccess;
	const int64_t groups = (n_leaves + kLeavesPerCta - 1) / kLeavesPerCta;
	const int grid = (int)(groups < (int64_t)num_sms ? groups : (int64_t)num_sms);
	decode_mma_kernel<<<grid, kThreads, kSmemBytes, stream>>>(w, dev_indices, n_leaves, dev_voxels, tap_stage, tap_out);
	return cudaGetLastError();
}

}  // namespace vqvdb
