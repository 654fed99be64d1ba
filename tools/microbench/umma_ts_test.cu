// Bring-up test for the tcgen05 pieces the tensor-core decoder needs, in isolation:
//   D[128 x 64] (fp32, TMEM) = A[128 x 64] (bf16) * B[64 x 64]^T (bf16, shared memory, SWIZZLE_128B K-major)
//   mode 0 (SS): A from shared memory through a UMMA descriptor
//   mode 1 (TS): A written into TMEM by the threads themselves with tcgen05.st (lane = row), the form an
//                implicit-GEMM convolution needs because the gather cannot be expressed by a descriptor.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o umma_ts_test umma_ts_test.cu && ./umma_ts_test
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
	// K-major, SWIZZLE_128B: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused for swizzled K-major.
	uint64_t d = 0;
	d |= (uint64_t)((saddr >> 4) & 0x3FFF);        // start address, bits [0,14)
	d |= (uint64_t)1 << 16;                        // leading byte offset (ignored), bits [16,30)
	d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset, bits [32,46)
	d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
	d |= (uint64_t)2 << 61;                        // layout type SWIZZLE_128B
	return d;
}

constexpr uint32_t kIdesc = (1u << 4)      // D format F32
                            | (1u << 7)    // A format BF16
                            | (1u << 10)   // B format BF16
                            | ((64u >> 3) << 17)    // N = 64
                            | ((128u >> 4) << 24);  // M = 128

__global__ void __launch_bounds__(128, 1)
umma_test(const __nv_bfloat16* __restrict__ A, const uint8_t* __restrict__ Bunit, float* __restrict__ D, int mode) {
	extern __shared__ __align__(1024) uint8_t smem[];
	uint8_t* sB = smem;              // 8 KB
	uint8_t* sA = smem + 8192;       // 16 KB (SS mode)
	uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8192 + 16384);
	uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 8192 + 16384 + 16);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
	}
	if (tid == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	for (int i = tid; i < 8192 / 16; i += 128) reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(Bunit)[i];
	// A tile in the same canonical layout: row r at r*128 B, 16-B chunk c at (c ^ (r & 7))
	for (int i = tid; i < 128 * 8; i += 128) {
		const int r = i >> 3, c = i & 7;
		*reinterpret_cast<uint4*>(sA + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(A + r * 64 + c * 8);
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core (async proxy)
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem = *tmem_slot;
	const uint32_t lane_base = (uint32_t)(warp * 32) << 16;

	if (mode == 1) {
		// each thread owns row tid: 64 bf16 = 32 packed words -> TMEM columns [64, 96)
		uint32_t r[32];
		const uint4* src = reinterpret_cast<const uint4*>(A + tid * 64);
#pragma unroll
		for (int q = 0; q < 8; ++q) {
			const uint4 v = src[q];
			r[4 * q] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
		}
		asm volatile(
		    "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
		    ::"r"(tmem + lane_base + 64), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
		    "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
		    "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]),
		    "r"(r[29]), "r"(r[30]), "r"(r[31])
		    : "memory");
		asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
		asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
		__syncthreads();
	}

	if (tid == 0) {
		asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
		const uint64_t bdesc = make_desc_sw128(smem_u32(sB));
		const uint64_t adesc = make_desc_sw128(smem_u32(sA));
#pragma unroll
		for (int kk = 0; kk < 4; ++kk) {
			const uint32_t acc = kk > 0 ? 1u : 0u;
			if (mode == 1) {
				asm volatile(
				    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
				    "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem),
				    "r"(tmem + 64 + kk * 8), "l"(bdesc + (uint64_t)(kk * 2)), "r"(kIdesc), "r"(acc)
				    : "memory");
			} else {
				asm volatile(
				    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
				    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
				    "l"(adesc + (uint64_t)(kk * 2)), "l"(bdesc + (uint64_t)(kk * 2)), "r"(kIdesc), "r"(acc)
				    : "memory");
			}
		}
		asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
	}
	// everyone waits for the MMAs
	asm volatile(
	    "{\n.reg .pred p;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra LAB_DONE;\nbra LAB_WAIT;\nLAB_DONE:\n}\n" ::"r"(smem_u32(bar)),
	    "r"(0)
	    : "memory");
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	uint32_t o[64];
	asm volatile(
	    "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
	    "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
	    "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
	    : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7]), "=r"(o[8]), "=r"(o[9]),
	      "=r"(o[10]), "=r"(o[11]), "=r"(o[12]), "=r"(o[13]), "=r"(o[14]), "=r"(o[15]), "=r"(o[16]), "=r"(o[17]), "=r"(o[18]), "=r"(o[19]),
	      "=r"(o[20]), "=r"(o[21]), "=r"(o[22]), "=r"(o[23]), "=r"(o[24]), "=r"(o[25]), "=r"(o[26]), "=r"(o[27]), "=r"(o[28]), "=r"(o[29]),
	      "=r"(o[30]), "=r"(o[31]), "=r"(o[32]), "=r"(o[33]), "=r"(o[34]), "=r"(o[35]), "=r"(o[36]), "=r"(o[37]), "=r"(o[38]), "=r"(o[39]),
	      "=r"(o[40]), "=r"(o[41]), "=r"(o[42]), "=r"(o[43]), "=r"(o[44]), "=r"(o[45]), "=r"(o[46]), "=r"(o[47]), "=r"(o[48]), "=r"(o[49]),
	      "=r"(o[50]), "=r"(o[51]), "=r"(o[52]), "=r"(o[53]), "=r"(o[54]), "=r"(o[55]), "=r"(o[56]), "=r"(o[57]), "=r"(o[58]), "=r"(o[59]),
	      "=r"(o[60]), "=r"(o[61]), "=r"(o[62]), "=r"(o[63])
	    : "r"(tmem + lane_base));
	asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
	for (int j = 0; j < 64; ++j) D[tid * 64 + j] = __uint_as_float(o[j]);
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

static uint16_t f2bf(float f) {
	uint32_t u;
	memcpy(&u, &f, 4);
	u += 0x7fffu + ((u >> 16) & 1u);
	return (uint16_t)(u >> 16);
}
static float bf2f(uint16_t b) {
	uint32_t u = (uint32_t)b << 16;
	float f;
	memcpy(&f, &u, 4);
	return f;
}

int main() {
	std::vector<uint16_t> A(128 * 64), B(64 * 64);
	srand(1);
	for (auto& v : A) v = f2bf((rand() % 2001 - 1000) / 500.f);
	for (auto& v : B) v = f2bf((rand() % 2001 - 1000) / 700.f);
	std::vector<uint8_t> unit(8192);
	for (int n = 0; n < 64; ++n)
		for (int k = 0; k < 64; ++k) memcpy(&unit[n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2], &B[n * 64 + k], 2);
	std::vector<float> ref(128 * 64);
	for (int m = 0; m < 128; ++m)
		for (int n = 0; n < 64; ++n) {
			float s = 0;
			for (int k = 0; k < 64; ++k) s += bf2f(A[m * 64 + k]) * bf2f(B[n * 64 + k]);
			ref[m * 64 + n] = s;
		}
	__nv_bfloat16* dA; uint8_t* dB; float* dD;
	cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, 8192); cudaMalloc(&dD, ref.size() * 4);
	cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
	cudaMemcpy(dB, unit.data(), 8192, cudaMemcpyHostToDevice);
	const int smem = 8192 + 16384 + 64;
	cudaFuncSetAttribute(umma_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	int rc = 0;
	for (int mode = 0; mode < 2; ++mode) {
		cudaMemset(dD, 0xff, ref.size() * 4);
		umma_test<<<1, 128, smem>>>(dA, dB, dD, mode);
		cudaError_t e = cudaDeviceSynchronize();
		std::vector<float> out(ref.size());
		cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
		double maxerr = 0;
		int bad = 0;
		for (size_t i = 0; i < out.size(); ++i) {
			const double d = fabs((double)out[i] - ref[i]);
			if (!(d <= 1e-2)) ++bad;
			if (d > maxerr || d != d) maxerr = d;
		}
		printf("mode %s: status %s, max |err| %.3g, %d of %zu outside 1e-2; D[0][0..3] = %g %g %g %g (ref %g %g %g %g); D[77][5]=%g (ref %g)\n",
		       mode ? "TS (A in TMEM)" : "SS (A in smem)", cudaGetErrorString(e), maxerr, bad, out.size(), out[0], out[1], out[2], out[3],
		       ref[0], ref[1], ref[2], ref[3], out[77 * 64 + 5], ref[77 * 64 + 5]);
		if (bad || e != cudaSuccess) rc = 1;
	}
	return rc;
}
