// Per-leaf building blocks shared by the encode and decode kernels (fp32 CUDA-core path).
//
// A CTA owns one whole 8^3 leaf at a time because GroupNorm and the channel attention reduce
// over the entire leaf (python/VQVAE_v2.py:190-228).  All inter-layer activations of a leaf stay
// in shared memory; HBM sees only the leaf voxels and the 64 index bytes.
//
// Shared-memory activation layouts
//   plain : [C][S*S*S]                      residual stream, conv outputs before GroupNorm
//   halo  : [C][S+2][S+2][S+2], zero border the zero-padded conv input (padding=1 everywhere)
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "ptx_utils.cuh"

namespace vqvdb {

constexpr float kGnEps = 1e-5f;        // nn.GroupNorm default eps (VQVAE_v2.py:196,198,236,258)
constexpr float kResScale = 0.1f;      // ResidualBlock scale (VQVAE_v2.py:193,210)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

__device__ __forceinline__ float sigmoid_f(float v) { return 1.f / (1.f + expf(-v)); }

// ---------------------------------------------------------------------------------------------
// Direct convolution, register-tiled: one work item = one output row (SO contiguous w outputs at
// fixed d,h) x TN output channels.  Weights are the transposed [cin][k][k][k][cout] table in
// global memory; a warp reads the same addresses (L1 broadcast).  Accumulation order: cin, kd, kh,
// kw ascending — the same order as the C oracle's conv3d().
//   in_halo : shared, [CIN][HP][HP][HP] with HP = SI + 2 (SI = input extent)
//   epi(oc0, od, oh, acc) receives acc[TN][SO] WITHOUT bias.
// ---------------------------------------------------------------------------------------------
template <int CIN, int COUT, int SO, int KS, int STRIDE, int TN, class Epi>
__device__ __forceinline__ void conv_rows(const float* in_halo, const float* __restrict__ wt, Epi epi) {
	constexpr int SI = SO * STRIDE;
	constexpr int HP = SI + 2;
	constexpr int NIN = (SO - 1) * STRIDE + KS;
	constexpr int ROWS = SO * SO;
	constexpr int NOG = COUT / TN;
	static_assert(COUT % TN == 0, "TN must divide COUT");
	static_assert(NIN <= HP, "row read stays inside the halo row");
	static_assert(TN == 1 || TN == 2 || TN % 4 == 0, "TN in {1,2,4k}");
	for (int item = threadIdx.x; item < ROWS * NOG; item += blockDim.x) {
		const int row = item % ROWS, og = item / ROWS;
		const int od = row / SO, oh = row % SO;
		float acc[TN][SO];
#pragma unroll
		for (int n = 0; n < TN; ++n)
#pragma unroll
			for (int j = 0; j < SO; ++j) acc[n][j] = 0.f;
		const float* wbase = wt + og * TN;
#pragma unroll 1
		for (int ic = 0; ic < CIN; ++ic) {
#pragma unroll 1
			for (int kd = 0; kd < KS; ++kd) {
#pragma unroll
				for (int kh = 0; kh < KS; ++kh) {
					const float* ip = in_halo + ((ic * HP + od * STRIDE + kd) * HP + oh * STRIDE + kh) * HP;
					float x[HP];
#pragma unroll
					for (int j = 0; j < HP / 2; ++j) {
						const float2 v = *reinterpret_cast<const float2*>(ip + 2 * j);
						x[2 * j] = v.x;
						x[2 * j + 1] = v.y;
					}
					const float* wp = wbase + (size_t)(((ic * KS + kd) * KS + kh) * KS) * COUT;
#pragma unroll
					for (int kw = 0; kw < KS; ++kw) {
						float w[TN];
						if constexpr (TN % 4 == 0) {
#pragma unroll
							for (int q = 0; q < TN / 4; ++q) {
								const float4 v = __ldg(reinterpret_cast<const float4*>(wp + kw * COUT) + q);
								w[4 * q] = v.x;
								w[4 * q + 1] = v.y;
								w[4 * q + 2] = v.z;
								w[4 * q + 3] = v.w;
							}
						} else if constexpr (TN == 2) {
							const float2 v = __ldg(reinterpret_cast<const float2*>(wp + kw * COUT));
							w[0] = v.x;
							w[1] = v.y;
						} else {
							w[0] = __ldg(wp + kw * COUT);
						}
#pragma unroll
						for (int j = 0; j < SO; ++j)
#pragma unroll
							for (int n = 0; n < TN; ++n) acc[n][j] = fmaf(x[j * STRIDE + kw], w[n], acc[n][j]);
					}
				}
			}
		}
		epi(og * TN, od, oh, acc);
	}
}

// ---------------------------------------------------------------------------------------------
// GroupNorm statistics over a plain [C][NSP] buffer: biased variance over (C/G x NSP), two-pass.
// One warp per group, groups strided over the CTA's warps.  Results go to s_mean/s_rstd[G].
// Caller must __syncthreads() before using them.
// ---------------------------------------------------------------------------------------------
template <int C, int G, int NSP>
__device__ __forceinline__ void gn_stats(const float* buf, float* s_mean, float* s_rstd) {
	constexpr int CNT = (C / G) * NSP;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
	for (int g = warp; g < G; g += nwarps) {
		const float* p = buf + g * CNT;
		float s = 0.f;
		for (int i = lane; i < CNT; i += 32) s += p[i];
		const float mean = warp_sum(s) * (1.f / CNT);
		float v = 0.f;
		for (int i = lane; i < CNT; i += 32) {
			const float d = p[i] - mean;
			v = fmaf(d, d, v);
		}
		const float var = warp_sum(v) * (1.f / CNT);
		if (lane == 0) {
			s_mean[g] = mean;
			s_rstd[g] = 1.f / sqrtf(var + kGnEps);
		}
	}
}

// relu(gn(x)) from a plain [C][S^3] buffer into a halo buffer [C][S+2]^3 (interior only).
template <int C, int G, int S>
__device__ __forceinline__ void gn_relu_to_halo(const float* src, float* dst_halo, const float* s_mean,
                                                const float* s_rstd, const float* __restrict__ gamma,
                                                const float* __restrict__ beta) {
	constexpr int NSP = S * S * S, HP = S + 2, CG = C / G;
	for (int i = threadIdx.x; i < C * NSP; i += blockDim.x) {
		const int c = i / NSP, p = i % NSP;
		const int d = p / (S * S), h = (p / S) % S, w = p % S;
		const int g = c / CG;
		float v = (src[i] - s_mean[g]) * s_rstd[g] * __ldg(gamma + c) + __ldg(beta + c);
		dst_halo[((c * HP + d + 1) * HP + h + 1) * HP + w + 1] = relu_f(v);
	}
}

// relu(gn(x)) in place on a plain buffer.
template <int C, int G, int NSP>
__device__ __forceinline__ void gn_relu_inplace(float* buf, const float* s_mean, const float* s_rstd,
                                                const float* __restrict__ gamma, const float* __restrict__ beta) {
	constexpr int CG = C / G;
	for (int i = threadIdx.x; i < C * NSP; i += blockDim.x) {
		const int c = i / NSP, g = c / CG;
		const float v = (buf[i] - s_mean[g]) * s_rstd[g] * __ldg(gamma + c) + __ldg(beta + c);
		buf[i] = relu_f(v);
	}
}

// plain [C][S^3] -> halo interior, no transform.
template <int C, int S>
__device__ __forceinline__ void copy_to_halo(const float* src, float* dst_halo) {
	constexpr int NSP = S * S * S, HP = S + 2;
	for (int i = threadIdx.x; i < C * NSP; i += blockDim.x) {
		const int c = i / NSP, p = i % NSP;
		const int d = p / (S * S), h = (p / S) % S, w = p % S;
		dst_halo[((c * HP + d + 1) * HP + h + 1) * HP + w + 1] = src[i];
	}
}

template <int N>
__device__ __forceinline__ void zero_smem(float* p) {
	for (int i = threadIdx.x; i < N; i += blockDim.x) p[i] = 0.f;
}

// ---------------------------------------------------------------------------------------------
// ChannelAttention (VQVAE_v2.py:213-228): y = sigmoid(W2 relu(W1 mean_dhw(x))), x *= y.
// x is plain [C][NSP]; fc0 is [R][C], fc2 is [C][R] (nn.Linear layout, no bias).
// s_tmp needs C + R floats.  Contains its own barriers; all threads must call it.
// ---------------------------------------------------------------------------------------------
template <int C, int R, int NSP>
__device__ __forceinline__ void channel_attention(float* x, const float* __restrict__ fc0,
                                                  const float* __restrict__ fc2, float* s_tmp) {
	float* s_mean = s_tmp;
	float* s_hid = s_tmp + C;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
	for (int c = warp; c < C; c += nwarps) {
		float s = 0.f;
		for (int i = lane; i < NSP; i += 32) s += x[c * NSP + i];
		s = warp_sum(s);
		if (lane == 0) s_mean[c] = s * (1.f / NSP);
	}
	__syncthreads();
	for (int j = warp; j < R; j += nwarps) {
		float s = 0.f;
		for (int c = lane; c < C; c += 32) s = fmaf(__ldg(fc0 + j * C + c), s_mean[c], s);
		s = warp_sum(s);
		if (lane == 0) s_hid[j] = relu_f(s);
	}
	__syncthreads();
	for (int i = threadIdx.x; i < C * NSP; i += blockDim.x) {
		const int c = i / NSP;
		float s = 0.f;
#pragma unroll
		for (int j = 0; j < R; ++j) s = fmaf(__ldg(fc2 + c * R + j), s_hid[j], s);
		x[i] *= sigmoid_f(s);
	}
	__syncthreads();
}

}  // namespace vqvdb
