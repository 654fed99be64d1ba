// Micro-benchmark: peak issue rates of the pipes the codec kernels can use on this GPU.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_rates pipe_rates.cu && ./pipe_rates
// Reports FFMA TFLOP/s and legacy mma.sync (HMMA bf16 m16n8k16, TF32 m16n8k8) TFLOP/s with register operands only.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__global__ void ffma_kernel(float* out, int iters) {
	float a[16];
	for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
	float x = 1.0001f, y = 0.9999f;
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x, y);
	}
	float s = 0;
	for (int i = 0; i < 16; ++i) s += a[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void hmma_kernel(float* out, int iters) {
	float c[8][4];
	for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
	uint32_t a0 = 0x3f803f80u, a1 = a0, a2 = a0, a3 = a0, b0 = 0x3c003c00u, b1 = b0;
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int i = 0; i < 8; ++i)
			asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
			             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
	}
	float s = 0;
	for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void tf32_kernel(float* out, int iters) {
	float c[8][4];
	for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
	uint32_t a0 = 0x3f800000u, a1 = a0, a2 = a0, a3 = a0, b0 = 0x3c000000u, b1 = b0;
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int i = 0; i < 8; ++i)
			asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
			             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
	}
	float s = 0;
	for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
float time_ms(K k, int grid, int block, float* out, int iters) {
	cudaEvent_t a, b;
	cudaEventCreate(&a); cudaEventCreate(&b);
	k<<<grid, block>>>(out, iters);
	cudaDeviceSynchronize();
	cudaEventRecord(a);
	k<<<grid, block>>>(out, iters);
	cudaEventRecord(b);
	cudaEventSynchronize(b);
	float ms; cudaEventElapsedTime(&ms, a, b);
	return ms;
}

int main() {
	cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
	const int sms = p.multiProcessorCount;
	float* out; cudaMalloc(&out, sizeof(float) * sms * 8 * 1024);
	for (int warps : {4, 8, 16, 32}) {
		const int block = warps * 32, grid = sms * (warps <= 8 ? 2 : 1), iters = 20000;
		const int wps = (warps <= 8 ? 2 : 1) * warps;
		float ms = time_ms(ffma_kernel, grid, block, out, iters);
		double ffma = 2.0 * 16 * iters * (double)grid * block / (ms * 1e-3) / 1e12;
		ms = time_ms(hmma_kernel, grid, block, out, iters);
		double hmma = 2.0 * 16 * 8 * 16 * 8 * iters * (double)grid * warps / (ms * 1e-3) / 1e12;
		ms = time_ms(tf32_kernel, grid, block, out, iters);
		double tf32 = 2.0 * 16 * 8 * 8 * 8 * iters * (double)grid * warps / (ms * 1e-3) / 1e12;
		printf("%s sms=%d warps/SM=%d  FFMA %.1f TFLOP/s  HMMA.bf16 %.1f TFLOP/s  MMA.tf32 %.1f TFLOP/s\n", p.name, sms, wps, ffma, hmma, tf32);
	}
	cudaError_t e = cudaDeviceSynchronize();
	printf("status %s\n", cudaGetErrorString(e));
	return 0;
}
