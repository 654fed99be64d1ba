// Fused fp32 encoder + codebook argmin: leaf voxels in, 64 uint8 indices out, one kernel.
//
// Replaces, for one batch, everything TorchBackend::encode runs on the device
// (/root/reference/src/backends/torch/TorchBackend.cpp:148-150): EncoderFloat.forward
// (python/VQVAE_v2.py:231-250) followed by InferenceVectorQuantizer.get_indices
// (python/save_for_inference.py:55-61) and the int64->uint8 cast.  The reference issues ~25
// library launches with an HBM round trip between each; here a persistent CTA
// keeps every activation of its leaf in shared memory and HBM sees 2048 B in + 64 B out per leaf.
//
// Arithmetic is fp32 FFMA throughout — index parity with the reference is a bit-exactness
// requirement and reduced-precision operands flip 0.2-2 % of indices.
#include "leaf_ops.cuh"
#include "model.cuh"

namespace vqvdb {

namespace {

constexpr int kEncThreads = 256;

struct EncSmem {
	// offsets in floats
	static constexpr int in_halo = 0;                  // [1][10][10][10]
	static constexpr int x16 = in_halo + 1000;         // [16][512]   residual stream at 8^3
	static constexpr int h16 = x16 + 16 * 512;         // [16][10^3]  conv input at 8^3
	static constexpr int t16 = h16 + 16 * 1000;        // [16][512]   conv1 output / z (aliased later)
	static constexpr int x32 = t16 + 16 * 512;         // [32][64]    residual stream at 4^3
	static constexpr int h32 = x32 + 32 * 64;          // [32][6^3]
	static constexpr int t32 = h32 + 32 * 216;         // [32][64]
	static constexpr int stats = t32 + 32 * 64;        // mean[8] rstd[8] tmp[48]
	static constexpr int vq_dist = stats + 64;         // [4][64]
	static constexpr int vq_idx = vq_dist + 256;       // [4][64] (int)
	static constexpr int total = vq_idx + 256;
	static constexpr int z = t16;                      // [128][64] latent, reuses t16 once the 8^3 stage is done
};
static_assert(EncSmem::total * 4 <= 227 * 1024, "encoder smem budget");

// ResidualBlock at spatial S with C channels: x += 0.1 * conv2(relu(gn2(conv1(relu(gn1(x))))))
template <int C, int S, int TN>
__device__ __forceinline__ void res_block(float* x, float* halo, float* tmp, float* s_mean, float* s_rstd,
                                          const ResWeights& w) {
	constexpr int NSP = S * S * S;
	gn_stats<C, 8, NSP>(x, s_mean, s_rstd);
	__syncthreads();
	gn_relu_to_halo<C, 8, S>(x, halo, s_mean, s_rstd, w.gn1_w, w.gn1_b);
	__syncthreads();
	const float* c1b = w.c1_b;
	conv_rows<C, C, S, 3, 1, TN>(halo, w.c1_w, [&](int oc0, int od, int oh, float (&acc)[TN][S]) {
#pragma unroll
		for (int n = 0; n < TN; ++n) {
			const float b = __ldg(c1b + oc0 + n);
#pragma unroll
			for (int j = 0; j < S; ++j) tmp[(oc0 + n) * NSP + (od * S + oh) * S + j] = acc[n][j] + b;
		}
	});
	__syncthreads();
	gn_stats<C, 8, NSP>(tmp, s_mean, s_rstd);
	__syncthreads();
	gn_relu_to_halo<C, 8, S>(tmp, halo, s_mean, s_rstd, w.gn2_w, w.gn2_b);
	__syncthreads();
	const float* c2b = w.c2_b;
	conv_rows<C, C, S, 3, 1, TN>(halo, w.c2_w, [&](int oc0, int od, int oh, float (&acc)[TN][S]) {
#pragma unroll
		for (int n = 0; n < TN; ++n) {
			const float b = __ldg(c2b + oc0 + n);
#pragma unroll
			for (int j = 0; j < S; ++j) {
				float* px = x + (oc0 + n) * NSP + (od * S + oh) * S + j;
				*px = *px + kResScale * (acc[n][j] + b);
			}
		}
	});
	__syncthreads();
}

__global__ void __launch_bounds__(kEncThreads, 1)
encode_fp32_kernel(const EncoderWeights w, const float* __restrict__ leaves, int64_t n_leaves,
                   uint8_t* __restrict__ indices) {
	extern __shared__ __align__(16) float smem[];
	float* s_in = smem + EncSmem::in_halo;
	float* x16 = smem + EncSmem::x16;
	float* h16 = smem + EncSmem::h16;
	float* t16 = smem + EncSmem::t16;
	float* x32 = smem + EncSmem::x32;
	float* h32 = smem + EncSmem::h32;
	float* t32 = smem + EncSmem::t32;
	float* s_mean = smem + EncSmem::stats;
	float* s_rstd = s_mean + 8;
	float* s_tmp = s_mean + 16;
	float* vq_dist = smem + EncSmem::vq_dist;
	int* vq_idx = reinterpret_cast<int*>(smem + EncSmem::vq_idx);
	float* zbuf = smem + EncSmem::z;
	const int tid = threadIdx.x;

	// Halo borders are written once and stay zero: every later write touches interiors only.
	zero_smem<1000>(s_in);
	zero_smem<16 * 1000>(h16);
	zero_smem<32 * 216>(h32);
	__syncthreads();

	for (int64_t leaf = blockIdx.x; leaf < n_leaves; leaf += gridDim.x) {
		// ---- stage the leaf: 2048 B, 128-bit coalesced loads, into the haloed input ----
		if (tid < 128) {
			const float4 v = __ldcs(reinterpret_cast<const float4*>(leaves + leaf * 512) + tid);
			const int p = tid * 4, d = p >> 6, h = (p >> 3) & 7, wq = p & 7;
			float* dst = s_in + ((d + 1) * 10 + h + 1) * 10 + wq + 1;
			dst[0] = v.x;
			dst[1] = v.y;
			dst[2] = v.z;
			dst[3] = v.w;
		}
		__syncthreads();

		// ---- pre.0: Conv3d(1,16,k3) -> t16 ; pre.1: GroupNorm(4,16) + ReLU -> x16 ----
		{
			const float* pb = w.pre_b;
			conv_rows<1, 16, 8, 3, 1, 4>(s_in, w.pre_w, [&](int oc0, int od, int oh, float (&acc)[4][8]) {
#pragma unroll
				for (int n = 0; n < 4; ++n) {
					const float b = __ldg(pb + oc0 + n);
#pragma unroll
					for (int j = 0; j < 8; ++j) t16[(oc0 + n) * 512 + (od * 8 + oh) * 8 + j] = acc[n][j] + b;
				}
			});
		}
		__syncthreads();
		gn_stats<16, 4, 512>(t16, s_mean, s_rstd);
		__syncthreads();
		for (int i = tid; i < 16 * 512; i += kEncThreads) {
			const int c = i >> 9, g = c >> 2;
			const float v = (t16[i] - s_mean[g]) * s_rstd[g] * __ldg(w.pre_gn_w + c) + __ldg(w.pre_gn_b + c);
			x16[i] = fmaxf(v, 0.f);
		}
		__syncthreads();

		// ---- pre.3: ResidualBlock(16) at 8^3 ----
		res_block<16, 8, 4>(x16, h16, t16, s_mean, s_rstd, w.res16);

		// ---- down: Conv3d(16,32,k4,s2,p1) : x16 (via halo) -> x32 ----
		copy_to_halo<16, 8>(x16, h16);
		__syncthreads();
		{
			const float* db = w.down_b;
			conv_rows<16, 32, 4, 4, 2, 4>(h16, w.down_w, [&](int oc0, int od, int oh, float (&acc)[4][4]) {
#pragma unroll
				for (int n = 0; n < 4; ++n) {
					const float b = __ldg(db + oc0 + n);
#pragma unroll
					for (int j = 0; j < 4; ++j) x32[(oc0 + n) * 64 + (od * 4 + oh) * 4 + j] = acc[n][j] + b;
				}
			});
		}
		__syncthreads();

		// ---- res_stack.0: ResidualBlock(32) at 4^3 ; attn: ChannelAttention(32) ----
		res_block<32, 4, 4>(x32, h32, t32, s_mean, s_rstd, w.res32);
		channel_attention<32, 8, 64>(x32, w.fc0, w.fc2, s_tmp);

		// ---- proj: Conv3d(32,128,k1) -> z [128][64] ----
		for (int i = tid; i < 128 * 64; i += kEncThreads) {
			const int oc = i >> 6, p = i & 63;
			float a = 0.f;
#pragma unroll 8
			for (int ic = 0; ic < 32; ++ic) a = fmaf(x32[ic * 64 + p], __ldg(w.proj_w + ic * 128 + oc), a);
			zbuf[i] = a + __ldg(w.proj_b + oc);
		}
		__syncthreads();

		// ---- VQ: argmin_k (|z|^2 + |e_k|^2) - 2 z.e_k, first minimum wins ----
		{
			const int p = tid & 63, q = tid >> 6;  // q is warp-uniform: 64 positions = 2 warps
			float dot[64];
#pragma unroll
			for (int k = 0; k < 64; ++k) dot[k] = 0.f;
			float zz = 0.f;
			const float* et = w.emb_t + q * 64;
#pragma unroll 1
			for (int d = 0; d < 128; ++d) {
				const float zv = zbuf[d * 64 + p];
				zz = fmaf(zv, zv, zz);
				const float4* e4 = reinterpret_cast<const float4*>(et + d * 256);
#pragma unroll
				for (int k4 = 0; k4 < 16; ++k4) {
					const float4 e = __ldg(e4 + k4);
					dot[4 * k4] = fmaf(zv, e.x, dot[4 * k4]);
					dot[4 * k4 + 1] = fmaf(zv, e.y, dot[4 * k4 + 1]);
					dot[4 * k4 + 2] = fmaf(zv, e.z, dot[4 * k4 + 2]);
					dot[4 * k4 + 3] = fmaf(zv, e.w, dot[4 * k4 + 3]);
				}
			}
			float best = INFINITY;
			int bi = 0;
#pragma unroll
			for (int k = 0; k < 64; ++k) {
				const float dist = (zz + __ldg(w.emb_sq + q * 64 + k)) - 2.f * dot[k];
				if (dist < best) {
					best = dist;
					bi = q * 64 + k;
				}
			}
			vq_dist[q * 64 + p] = best;
			vq_idx[q * 64 + p] = bi;
		}
		__syncthreads();
		if (tid < 64) {
			float best = vq_dist[tid];
			int bi = vq_idx[tid];
#pragma unroll
			for (int q = 1; q < 4; ++q) {
				const float dq = vq_dist[q * 64 + tid];
				if (dq < best) {
					best = dq;
					bi = vq_idx[q * 64 + tid];
				}
			}
			indices[leaf * 64 + tid] = (uint8_t)bi;  // latent position p = (d*4+h)*4+w, matching view(B,4,4,4)
		}
		__syncthreads();
	}
}

}  // namespace

cudaError_t configure_encode_fp32() {
	return cudaFuncSetAttribute(encode_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
	                            EncSmem::total * (int)sizeof(float));
}

cudaError_t launch_encode_fp32(const EncoderWeights& w, const float* dev_leaves, int64_t n_leaves,
                               uint8_t* dev_indices, int num_sms, cudaStream_t stream) {
	if (n_leaves <= 0) return cudaSuccess;
	const int grid = (int)(n_leaves < (int64_t)num_sms ? n_leaves : (int64_t)num_sms);
	encode_fp32_kernel<<<grid, kEncThreads, EncSmem::total * sizeof(float), stream>>>(w, dev_leaves, n_leaves,
	                                                                                  dev_indices);
	return cudaGetLastError();
}

}  // namespace vqvdb
