// tcgen05.mma issue/complete rate vs N for the two operand modes the kernels use (M = 128, K = 16, kind::f16):
//   TS: A from TMEM (8 columns), B from shared memory (no-swizzle K-major [2][N][16 B])   — the decoder's mode
//   SS: A and B from shared memory                                                          — the encoder's mode
// Prints cycles per MMA next to the floor 128*N/256.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o umma_rate umma_rate.cu && ./umma_rate
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__host__ __device__ constexpr uint32_t idesc(uint32_t n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
	return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
	asm volatile("{\n.reg .pred p;\nLAB_WAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n@p bra LAB_DONE_%=;\nbra LAB_WAIT_%=;\nLAB_DONE_%=:\n}\n" ::"r"(bar), "r"(parity), "r"(0x989680) : "memory");
}
template <int N, bool TS>
__global__ void __launch_bounds__(64, 1) rate(long long* out, int reps) {
	extern __shared__ __align__(1024) uint8_t smem[];
	uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 64 * 1024);
	uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
	for (uint32_t i = threadIdx.x; i < 64 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (threadIdx.x < 32) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem = *slot;
	if (threadIdx.x == 32) {
		const uint64_t adesc = make_desc(smem_u32(smem), 2048, 128);             // A: [2][128 rows][16 B]
		const uint64_t bdesc = make_desc(smem_u32(smem) + 8192, N * 16, 128);     // B: [2][N][16 B]
		const long long t0 = clock64();
		for (int r = 0; r < reps; ++r) {
			const uint32_t d = tmem + (N == 256 ? 0 : (r & 1) * N);                // alternate accumulators when they fit beside A
			const uint32_t acc = 1u;
			if (TS) {
				asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(tmem + 504), "l"(bdesc), "r"(idesc(N)), "r"(acc) : "memory");
			} else {
				asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc(N)), "r"(acc) : "memory");
			}
		}
		const long long t1 = clock64();
		asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
		mbar_wait(smem_u32(bar), 0);
		const long long t2 = clock64();
		if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
template <int N, bool TS> void run(long long* d) {
	const int smem = 64 * 1024 + 64, reps = 4000;
	cudaFuncSetAttribute(rate<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	for (int grid : {1, 148}) {
		rate<N, TS><<<grid, 64, smem>>>(d, reps);
		cudaError_t e = cudaDeviceSynchronize();
		long long h[2];
		cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
		printf("%s N=%3d grid %3d: %s  issue %.1f cyc/MMA, complete %.1f cyc/MMA (floor %d)\n", TS ? "TS" : "SS", N, grid, cudaGetErrorString(e),
		       (double)h[0] / reps, (double)h[1] / reps, 128 * N / 256);
	}
}
int main() {
	long long* d;
	cudaMalloc(&d, 64);
	run<16, true>(d); run<64, true>(d); run<128, true>(d); run<256, true>(d);
	run<16, false>(d); run<64, false>(d); run<128, false>(d); run<256, false>(d);
	return 0;
}
