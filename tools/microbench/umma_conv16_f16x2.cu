// Bring-up of the tensor-core encoder's 8^3 convolution in its final form, in isolation:
//   Conv3d(16 -> 16, k3, p1) over one 8^3 leaf on tcgen05.mma (SS mode, kind::f16) with
//   * fp32-level accuracy from a 2-way fp16 split of both operands, a = a_hi + 2^-11 a_lo, w = w_hi + 2^-11 w_lo:
//     three products  a_hi w_hi | a_hi w_lo + a_lo w_hi  in two accumulator groups (the dropped a_lo w_lo term is 2^-22);
//   * the flattened zero-halo activation layout q = d*72 + h*8 + w (h has a ninth, all-zero row block, d a zero slab
//     before and after; NO halo along w), channels-last in two 8-channel planes per precision — the canonical
//     no-swizzle K-major UMMA layout, so a (kd, kh) filter offset is a shifted descriptor start address;
//   * the three kw taps concatenated along N (one A read serves three taps): per (kd, kh)
//       MMA1: A_hi x [w_hi(kw0..2) ; w_lo(kw0..2)]  (N = 96)  -> columns [0, 96)
//       MMA2: A_lo x [w_hi(kw0..2)]                 (N = 48)  -> columns [48, 96)  (accumulates onto MMA1's w_lo group)
//     and the kw shift applied in the epilogue:  out[q] = P0[q-1] (w > 0) + P1[q] + P2[q+1] (w < 7), a lane shuffle
//     that never crosses a warp because w = lane & 7.
// Prints the error against an fp64 convolution next to a plain fp32 FMA chain's error, and the cycle counts.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o umma_conv16_f16x2 umma_conv16_f16x2.cu && ./umma_conv16_f16x2
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int kMargin = 80;
constexpr int kRows = kMargin + 640 + 80;          // 800
constexpr uint32_t kPlane = kRows * 16;            // one 8-channel plane
constexpr uint32_t kPrec = 2 * kPlane;             // 16 channels of one precision
constexpr uint32_t kABytes = 2 * kPrec;            // 51 200
constexpr uint32_t kBUnit = 2 * 96 * 16;           // per (kd,kh): [k-chunk][n = 96][16 B]
constexpr uint32_t kBBytes = 9 * kBUnit;           // 27 648
constexpr float kLoScale = 2048.f, kLoInv = 1.f / 2048.f;

__host__ __device__ constexpr uint32_t idesc_f16(uint32_t n) { return (1u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24); }
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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
	asm volatile(
	    "{\n.reg .pred p;\nLAB_WAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n@p bra LAB_DONE_%=;\nbra LAB_WAIT_%=;\nLAB_DONE_%=:\n}\n" ::"r"(bar),
	    "r"(parity), "r"(0x989680)
	    : "memory");
}
// fp32 -> (hi, lo) fp16 pair with v ~= hi + lo / 2048
__device__ __forceinline__ void split_f16(float v, __half& hi, __half& lo) {
	hi = __float2half_rn(v);
	lo = __float2half_rn((v - __half2float(hi)) * kLoScale);
}

// x: [16][512] fp32 (channel-major leaf), wB: prepared B units, out: [640][16] fp32 (flattened q rows), cyc: clocks
__global__ void __launch_bounds__(160, 1)
conv_test(const float* __restrict__ x, const uint8_t* __restrict__ wB, float* __restrict__ out, long long* __restrict__ cyc, int reps) {
	extern __shared__ __align__(1024) uint8_t smem[];
	uint8_t* sA = smem;
	uint8_t* sB = smem + kABytes;
	uint64_t* bar = reinterpret_cast<uint64_t*>(sB + kBBytes);
	uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

	for (uint32_t i = tid; i < kABytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
	for (uint32_t i = tid; i < kBBytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(wB)[i];
	__syncthreads();
	for (int i = tid; i < 16 * 512; i += blockDim.x) {
		const int c = i >> 9, p = i & 511, d = p >> 6, h = (p >> 3) & 7, w = p & 7;
		__half hi, lo;
		split_f16(x[i], hi, lo);
		const uint32_t off = (c >> 3) * kPlane + (uint32_t)(kMargin + d * 72 + h * 8 + w) * 16 + (c & 7) * 2;
		*reinterpret_cast<__half*>(sA + off) = hi;
		*reinterpret_cast<__half*>(sA + kPrec + off) = lo;
	}
	if (tid == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 4) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem = *tmem_slot;

	if (tid == 128) {
		const long long t0 = clock64();
		const uint64_t a_d0 = make_desc(smem_u32(sA) + kMargin * 16, kPlane, 128);
		const uint64_t b_d0 = make_desc(smem_u32(sB), 96 * 16, 128);
		constexpr uint64_t kPrec16 = kPrec >> 4;
		for (int r = 0; r < reps; ++r) {
#pragma unroll 1
			for (int t = 0; t < 5; ++t) {
				const uint32_t dcol = tmem + t * 96;
				const uint64_t at = a_d0 + (uint64_t)(128 * t);
#pragma unroll
				for (int kk = 0; kk < 9; ++kk) {
					const int s = (kk / 3 - 1) * 72 + (kk % 3 - 1) * 8;
					const uint64_t ad = at + (uint64_t)(int64_t)s, bd = b_d0 + (uint64_t)(kk * (kBUnit >> 4));
					mma_ss(dcol, ad, bd, idesc_f16(96), (kk > 0 || r > 0) ? 1u : 0u);
					mma_ss(dcol + 48, ad + kPrec16, bd, idesc_f16(48), 1u);
				}
			}
		}
		const long long t1 = clock64();
		asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
		mbar_wait(smem_u32(bar), 0);
		const long long t2 = clock64();
		if (blockIdx.x == 0) {
			cyc[0] = t1 - t0;
			cyc[1] = t2 - t0;
		}
	} else if (tid < 128) {
		mbar_wait(smem_u32(bar), 0);
		asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
		const long long c0 = clock64();
		const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
		const int w = lane & 7;
		for (int t = 0; t < 5; ++t) {
			float hh[32], mix[32], hl[32];  // columns [0,32) = hh kw0, kw1 ; [32,64) = hh kw2, hl kw0 ; [64,96) = hl kw1, kw2
			tmem_ld32(lane_addr + t * 96, hh);
			tmem_ld32(lane_addr + t * 96 + 32, mix);
			tmem_ld32(lane_addr + t * 96 + 64, hl);
			float o[16];
#pragma unroll
			for (int c = 0; c < 16; ++c) {
				const float p0 = fmaf(mix[16 + c], kLoInv, hh[c]);
				const float p1 = fmaf(hl[c], kLoInv, hh[16 + c]);
				const float p2 = fmaf(hl[16 + c], kLoInv, mix[c]);
				const float up = __shfl_up_sync(0xffffffffu, p0, 1), dn = __shfl_down_sync(0xffffffffu, p2, 1);
				o[c] = ((w > 0 ? up : 0.f) + p1) + (w < 7 ? dn : 0.f);
			}
			if (blockIdx.x == 0) {
#pragma unroll
				for (int c = 0; c < 16; ++c) out[(size_t)(t * 128 + tid) * 16 + c] = o[c];
			}
		}
		if (tid == 0 && blockIdx.x == 0) cyc[2] = clock64() - c0;
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

static uint16_t f2h(float f) {
	__half h = __float2half_rn(f);
	uint16_t u;
	memcpy(&u, &h, 2);
	return u;
}
static float h2f(uint16_t u) {
	__half h;
	memcpy(&h, &u, 2);
	return __half2float(h);
}

int main() {
	std::mt19937 rng(7);
	std::normal_distribution<float> nd(0.f, 1.f);
	std::vector<float> x(16 * 512), w(16 * 16 * 27);  // w[cout][cin][tap]
	for (auto& v : x) v = std::max(0.f, nd(rng) * 1.3f + 0.2f);
	for (auto& v : w) v = nd(rng) * 0.06f;
	// B units: per (kd,kh): [k-chunk j][n][8 cin], n = kw*16 + cout for w_hi, 48 + kw*16 + cout for w_lo
	std::vector<uint8_t> B(kBBytes);
	for (int tap = 0; tap < 27; ++tap)
		for (int co = 0; co < 16; ++co)
			for (int ci = 0; ci < 16; ++ci) {
				const float v = w[(co * 16 + ci) * 27 + tap];
				const uint16_t hi = f2h(v), lo = f2h((v - h2f(hi)) * 2048.f);
				const int kk = tap / 3, kw = tap % 3;
				const size_t base = (size_t)kk * kBUnit + (size_t)(ci >> 3) * 96 * 16 + (ci & 7) * 2;
				memcpy(&B[base + (size_t)(kw * 16 + co) * 16], &hi, 2);
				memcpy(&B[base + (size_t)(48 + kw * 16 + co) * 16], &lo, 2);
			}
	std::vector<double> ref(512 * 16);
	std::vector<float> ref32(512 * 16);
	for (int p = 0; p < 512; ++p) {
		const int d = p >> 6, h = (p >> 3) & 7, ww = p & 7;
		for (int co = 0; co < 16; ++co) {
			double s = 0;
			float s32 = 0.f;
			for (int ci = 0; ci < 16; ++ci)
				for (int tap = 0; tap < 27; ++tap) {
					const int dd = d + tap / 9 - 1, hh = h + (tap / 3) % 3 - 1, w2 = ww + tap % 3 - 1;
					if (dd < 0 || dd > 7 || hh < 0 || hh > 7 || w2 < 0 || w2 > 7) continue;
					const float a = x[ci * 512 + dd * 64 + hh * 8 + w2], b = w[(co * 16 + ci) * 27 + tap];
					s += (double)a * (double)b;
					s32 = fmaf(a, b, s32);
				}
			ref[p * 16 + co] = s;
			ref32[p * 16 + co] = s32;
		}
	}
	float *dx, *dout;
	uint8_t* dB;
	long long* dcyc;
	cudaMalloc(&dx, x.size() * 4);
	cudaMalloc(&dout, 640 * 16 * 4);
	cudaMalloc(&dB, kBBytes);
	cudaMalloc(&dcyc, 32);
	cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice);
	cudaMemcpy(dB, B.data(), kBBytes, cudaMemcpyHostToDevice);
	const int smem = kABytes + kBBytes + 64;
	cudaFuncSetAttribute(conv_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	int rc = 0;
	cudaMemset(dout, 0xff, 640 * 16 * 4);
	conv_test<<<1, 160, smem>>>(dx, dB, dout, dcyc, 1);
	cudaError_t e = cudaDeviceSynchronize();
	std::vector<float> out(640 * 16);
	cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
	double maxerr = 0, maxerr32 = 0, sumsq = 0, sumsq32 = 0, maxref = 0;
	int nbad = 0;
	for (int q = 0; q < 576; ++q) {
		const int d = q / 72, h = (q / 8) % 9, ww = q % 8;
		if (h > 7) continue;
		const int p = d * 64 + h * 8 + ww;
		for (int co = 0; co < 16; ++co) {
			const double o = out[q * 16 + co];
			const double er = fabs(o - ref[p * 16 + co]), er32 = fabs((double)ref32[p * 16 + co] - ref[p * 16 + co]);
			if (!(er <= maxerr)) maxerr = er;
			if (!(er <= 2e-5)) ++nbad;
			maxerr32 = std::max(maxerr32, er32);
			sumsq += er * er;
			sumsq32 += er32 * er32;
			maxref = std::max(maxref, fabs(ref[p * 16 + co]));
		}
	}
	printf("status %s | max|ref| %.3f | tcgen05 fp16x2 kw-concat: max err %.3e rms %.3e, %d of 8192 > 2e-5 | fp32 FMA chain: max err %.3e rms %.3e\n",
	       cudaGetErrorString(e), maxref, maxerr, sqrt(sumsq / 8192), nbad, maxerr32, sqrt(sumsq32 / 8192));
	if (e != cudaSuccess || !(maxerr < 1e-4)) rc = 1;
	long long cyc[3];
	cudaMemcpy(cyc, dcyc, 24, cudaMemcpyDeviceToHost);
	printf("single pass: MMA issue %lld cyc, complete %lld cyc; epilogue (TMEM load + combine + shuffle + global store) %lld cyc\n", cyc[0], cyc[1], cyc[2]);
	for (int grid : {1, 148}) {
		const int reps = 40;
		conv_test<<<grid, 160, smem>>>(dx, dB, dout, dcyc, reps);
		e = cudaDeviceSynchronize();
		cudaMemcpy(cyc, dcyc, 24, cudaMemcpyDeviceToHost);
		printf("grid %3d: %s, per (tile, kd, kh) [2 MMAs N=96/48]: issue %.1f cyc, complete %.1f cyc; per leaf-conv %.0f cyc\n", grid,
		       cudaGetErrorString(e), cyc[0] / (reps * 45.0), cyc[1] / (reps * 45.0), (double)cyc[1] / reps);
		if (e != cudaSuccess) rc = 1;
	}
	return rc;
}
