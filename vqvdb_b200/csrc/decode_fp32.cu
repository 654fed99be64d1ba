// Fused fp32 decoder: 64 uint8 indices in, 512 voxels out, one kernel, CUDA cores only.
//
// This is the bring-up / checking path (VQVDB_B200_DECODE_FP32): same dataflow as the tensor-core
// decoder but every multiply-add is an fp32 FFMA, so its output is comparable to the reference's
// CPU fp32 result to ~1e-6.  Replaces what TorchBackend::decode runs on the device
// (/root/reference/src/backends/torch/TorchBackend.cpp:179-181): F.embedding + permute
// (python/save_for_inference.py:91-101) and DecoderFloat.forward (python/VQVAE_v2.py:253-275)
// including PixelShuffle3D (:172-187), which here is only a change of store address.
#include "leaf_ops.cuh"
#include "model.cuh"

namespace vqvdb {

namespace {

constexpr int kDecThreads = 256;

struct DecSmem {
	static constexpr int r0 = 0;                      // Q halo [128][6^3] during the stem, then P halo [32][10^3]
	static constexpr int x64 = r0 + 32 * 1000;        // [64][64] residual stream
	static constexpr int t64 = x64 + 64 * 64;         // [64][64] conv1 output; split-K partials of the final conv
	static constexpr int h64 = t64 + 64 * 64;         // [64][6^3] conv input
	static constexpr int stats = h64 + 64 * 216;      // mean[8] rstd[8] tmp[80]
	static constexpr int idx = stats + 96;            // 64 bytes
	static constexpr int total = idx + 16;
};
static_assert(DecSmem::total * 4 <= 227 * 1024, "decoder smem budget");

__global__ void __launch_bounds__(kDecThreads, 1)
decode_fp32_kernel(const DecoderWeights w, const uint8_t* __restrict__ indices, int64_t n_leaves,
                   float* __restrict__ voxels) {
	extern __shared__ __align__(16) float smem[];
	float* r0 = smem + DecSmem::r0;
	float* x64 = smem + DecSmem::x64;
	float* t64 = smem + DecSmem::t64;
	float* h64 = smem + DecSmem::h64;
	float* s_mean = smem + DecSmem::stats;
	float* s_rstd = s_mean + 8;
	float* s_tmp = s_mean + 16;
	uint8_t* s_idx = reinterpret_cast<uint8_t*>(smem + DecSmem::idx);
	const int tid = threadIdx.x;

	zero_smem<64 * 216>(h64);
	__syncthreads();

	for (int64_t leaf = blockIdx.x; leaf < n_leaves; leaf += gridDim.x) {
		if (tid < 16)
			reinterpret_cast<uint32_t*>(s_idx)[tid] = __ldcs(reinterpret_cast<const uint32_t*>(indices + leaf * 64) + tid);
		zero_smem<128 * 216>(r0);
		__syncthreads();

		// ---- codebook gather: q[d][p] = embedding[idx[p]][d], into the haloed stem input ----
		for (int i = tid; i < 64 * 128; i += kDecThreads) {
			const int p = i >> 7, d = i & 127;
			const int pd = p >> 4, ph = (p >> 2) & 3, pw = p & 3;
			r0[((d * 6 + pd + 1) * 6 + ph + 1) * 6 + pw + 1] = __ldg(w.emb + (int)s_idx[p] * 128 + d);
		}
		__syncthreads();

		// ---- stem.0: Conv3d(128,64,k3) -> t64 ; stem.1 GroupNorm(8,64) + ReLU -> x64 ----
		{
			const float* sb = w.stem_b;
			conv_rows<128, 64, 4, 3, 1, 4>(r0, w.stem_w, [&](int oc0, int od, int oh, float (&acc)[4][4]) {
#pragma unroll
				for (int n = 0; n < 4; ++n) {
					const float b = __ldg(sb + oc0 + n);
#pragma unroll
					for (int j = 0; j < 4; ++j) t64[(oc0 + n) * 64 + (od * 4 + oh) * 4 + j] = acc[n][j] + b;
				}
			});
		}
		__syncthreads();
		gn_stats<64, 8, 64>(t64, s_mean, s_rstd);
		__syncthreads();
		for (int i = tid; i < 64 * 64; i += kDecThreads) {
			const int c = i >> 6, g = c >> 3;
			const float v = (t64[i] - s_mean[g]) * s_rstd[g] * __ldg(w.stem_gn_w + c) + __ldg(w.stem_gn_b + c);
			x64[i] = fmaxf(v, 0.f);
		}
		__syncthreads();

		// ---- res_stack.0: ResidualBlock(64) ----
		{
			const ResWeights& rw = w.res64;
			gn_stats<64, 8, 64>(x64, s_mean, s_rstd);
			__syncthreads();
			gn_relu_to_halo<64, 8, 4>(x64, h64, s_mean, s_rstd, rw.gn1_w, rw.gn1_b);
			__syncthreads();
			const float* c1b = rw.c1_b;
			conv_rows<64, 64, 4, 3, 1, 4>(h64, rw.c1_w, [&](int oc0, int od, int oh, float (&acc)[4][4]) {
#pragma unroll
				for (int n = 0; n < 4; ++n) {
					const float b = __ldg(c1b + oc0 + n);
#pragma unroll
					for (int j = 0; j < 4; ++j) t64[(oc0 + n) * 64 + (od * 4 + oh) * 4 + j] = acc[n][j] + b;
				}
			});
			__syncthreads();
			gn_stats<64, 8, 64>(t64, s_mean, s_rstd);
			__syncthreads();
			gn_relu_to_halo<64, 8, 4>(t64, h64, s_mean, s_rstd, rw.gn2_w, rw.gn2_b);
			__syncthreads();
			const float* c2b = rw.c2_b;
			conv_rows<64, 64, 4, 3, 1, 4>(h64, rw.c2_w, [&](int oc0, int od, int oh, float (&acc)[4][4]) {
#pragma unroll
				for (int n = 0; n < 4; ++n) {
					const float b = __ldg(c2b + oc0 + n);
#pragma unroll
					for (int j = 0; j < 4; ++j) {
						float* px = x64 + (oc0 + n) * 64 + (od * 4 + oh) * 4 + j;
						*px = *px + kResScale * (acc[n][j] + b);
					}
				}
			});
			__syncthreads();
		}

		// ---- attn: ChannelAttention(64) ----
		channel_attention<64, 16, 64>(x64, w.fc0, w.fc2, s_tmp);

		// ---- up_conv: Conv3d(64,256,k3) + PixelShuffle3D(2) -> P halo [32][10^3] ----
		copy_to_halo<64, 4>(x64, h64);
		zero_smem<32 * 1000>(r0);
		__syncthreads();
		{
			const float* ub = w.up_b;
			conv_rows<64, 256, 4, 3, 1, 4>(h64, w.up_w, [&](int oc0, int od, int oh, float (&acc)[4][4]) {
#pragma unroll
				for (int n = 0; n < 4; ++n) {
					const int c = oc0 + n;  // = oc*8 + rd*4 + rh*2 + rw   (VQVAE_v2.py:184-186)
					const int oc = c >> 3, rd = (c >> 2) & 1, rh = (c >> 1) & 1, rwb = c & 1;
					const float b = __ldg(ub + c);
					float* row = r0 + ((oc * 10 + 2 * od + rd + 1) * 10 + 2 * oh + rh + 1) * 10 + rwb + 1;
#pragma unroll
					for (int j = 0; j < 4; ++j) row[2 * j] = acc[n][j] + b;
				}
			});
		}
		__syncthreads();

		// ---- final: Conv3d(32,1,k3) + sigmoid.  Split-K: each half of the CTA takes 16 input channels. ----
		{
			const int half = tid >> 7, t = tid & 127;
			const int row = t >> 1, w0 = (t & 1) * 4;
			const int od = row >> 3, oh = row & 7;
			float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
			for (int ic = half * 16; ic < half * 16 + 16; ++ic) {
#pragma unroll
				for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
					for (int kh = 0; kh < 3; ++kh) {
						const float* ip = r0 + ((ic * 10 + od + kd) * 10 + oh + kh) * 10 + w0;
						float x[6];
#pragma unroll
						for (int j = 0; j < 3; ++j) {
							const float2 v = *reinterpret_cast<const float2*>(ip + 2 * j);
							x[2 * j] = v.x;
							x[2 * j + 1] = v.y;
						}
						const float* wp = w.fin_w + ((ic * 3 + kd) * 3 + kh) * 3;
#pragma unroll
						for (int kw = 0; kw < 3; ++kw) {
							const float wv = __ldg(wp + kw);
#pragma unroll
							for (int j = 0; j < 4; ++j) acc[j] = fmaf(x[j + kw], wv, acc[j]);
						}
					}
				}
			}
			if (half == 1) {
#pragma unroll
				for (int j = 0; j < 4; ++j) t64[row * 8 + w0 + j] = acc[j];
			}
			__syncthreads();
			if (half == 0) {
				const float b = __ldg(w.fin_b);
				float4 o;
				o.x = sigmoid_f(acc[0] + t64[row * 8 + w0 + 0] + b);
				o.y = sigmoid_f(acc[1] + t64[row * 8 + w0 + 1] + b);
				o.z = sigmoid_f(acc[2] + t64[row * 8 + w0 + 2] + b);
				o.w = sigmoid_f(acc[3] + t64[row * 8 + w0 + 3] + b);
				__stcs(reinterpret_cast<float4*>(voxels + leaf * 512) + t, o);  // 128 x 16 B = the whole leaf, coalesced
			}
		}
		__syncthreads();
	}
}

}  // namespace

cudaError_t configure_decode_fp32() {
	return cudaFuncSetAttribute(decode_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
	                            DecSmem::total * (int)sizeof(float));
}

cudaError_t launch_decode_fp32(const DecoderWeights& w, const uint8_t* dev_indices, int64_t n_leaves,
                               float* dev_voxels, int num_sms, cudaStream_t stream) {
	if (n_leaves <= 0) return cudaSuccess;
	const int grid = (int)(n_leaves < (int64_t)num_sms ? n_leaves : (int64_t)num_sms);
	decode_fp32_kernel<<<grid, kDecThreads, DecSmem::total * sizeof(float), stream>>>(w, dev_indices, n_leaves,
	                                                                                  dev_voxels);
	return cudaGetLastError();
}

}  // namespace vqvdb
