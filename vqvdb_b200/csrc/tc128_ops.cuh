// Device helpers shared by the vec3 model's tensor-core kernels (decode_tc128.cu, encode_tc128.cu, encode_tc128_front.cu):
// tcgen05 / TMEM wrappers for TS-mode MMAs (A operand in tensor memory, B through a SWIZZLE_128B shared-memory
// descriptor), fp16 hi/lo splitting, small shared-memory accessors.
#pragma once

#include <cuda_fp16.h>

#include <cstdint>

#include "ptx_utils.cuh"

namespace vqvdb {
namespace tc128 {

constexpr float kLoScale = 2048.f, kLoInv = 1.f / 2048.f;

// instruction descriptor: D = f32, A = B = fp16, both K-major, M = 128
__host__ __device__ constexpr uint32_t idesc_f16(uint32_t n) { return (1u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24); }

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem_d] (+)= A[tmem_a] (128 x 16 fp16, TMEM) * B[desc] (N x 16 fp16, shared)^T
__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
	asm volatile(
	    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
	    "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
	    "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
	    : "memory");
}
// D[tmem_d] (+)= A[desc] (128 x 16 fp16, shared) * B[desc] (N x 16 fp16, shared)^T
__device__ __forceinline__ void tc_mma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
	asm volatile(
	    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
	    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
	    "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
	    : "memory");
}
// generic-proxy shared-memory writes -> visible to the tensor core's (async proxy) operand reads
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// K-major SWIZZLE_128B operand: 128-byte rows, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
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
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
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
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
	asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float lds32(uint32_t a) {
	float v;
	asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
	return v;
}
__device__ __forceinline__ void sts32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }

// two fp32 values -> packed fp16 hi pair and fp16 lo pair (v ~= hi + lo / 2048)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
	const __half2 hh = __floats2half2_rn(a, b);
	const float2 hf = __half22float2(hh);
	const __half2 ll = __floats2half2_rn((a - hf.x) * kLoScale, (b - hf.y) * kLoScale);
	hi = *reinterpret_cast<const uint32_t*>(&hh);
	lo = *reinterpret_cast<const uint32_t*>(&ll);
}
// eight consecutive channels -> one 16-byte chunk of the hi plane and one of the lo plane
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
	split2(v[0], v[1], hi.x, lo.x);
	split2(v[2], v[3], hi.y, lo.y);
	split2(v[4], v[5], hi.z, lo.z);
	split2(v[6], v[7], hi.w, lo.w);
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

}  // namespace tc128
}  // namespace vqvdb
