// Thin wrappers over the sm_90+/sm_100 PTX the kernels use: mbarrier, 1-D TMA bulk copies, ldmatrix, warp MMA.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace vqvdb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Blocking wait.  The suspend-time hint lets the hardware park the thread until the phase completes (or the hint
// expires) instead of re-issuing the probe: spinning waiters otherwise eat a third of the issue slots of a
// warp-specialised kernel.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "LAB_WAIT_%=:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
	    "@p bra LAB_DONE_%=;\n"
	    "bra LAB_WAIT_%=;\n"
	    "LAB_DONE_%=:\n"
	    "}\n" ::"r"(bar),
	    "r"(parity), "r"(0x989680)
	    : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
	uint32_t ok;
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
	    "selp.u32 %0, 1, 0, p;\n"
	    "}\n"
	    : "=r"(ok)
	    : "r"(bar), "r"(parity)
	    : "memory");
	return ok != 0;
}
// 1-D bulk copy global -> shared through the TMA unit; completion is signalled on `bar` (complete_tx).
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
	             "l"(src), "r"(bytes), "r"(bar)
	             : "memory");
}
// Packed fp32 pairs (sm_100: FFMA2): d = a * b + c on both halves, each an IEEE fma — bit-identical to two fmaf().  When
// both halves of `a` hold the same value ptxas uses the instruction's scalar-broadcast operand form.
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
	uint64_t r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
	return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
	uint64_t d;
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
	return d;
}
// ReLU that propagates NaN, as torch.relu does (fmaxf(v, 0) returns 0 for a NaN): a leaf with a non-finite voxel must come
// out of the encoder as NaN everywhere — the reference's argmin then yields code 0 for all of its latents.
__device__ __forceinline__ float relu_f(float v) {
	float r;
	asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(r) : "f"(v));
	return r;
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
	asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
	             : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
	             : "r"(addr));
}
// D(16x8, f32) += A(16x16, bf16, row) * B(16x8, bf16, col)
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
	asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
	             : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
	             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
	__nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
	return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t w) {
	return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w));
}

}  // namespace vqvdb
