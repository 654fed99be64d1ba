// Architecture-generic fp32 encode / decode kernels: one CTA per leaf, runtime channel counts, activations in a
// per-CTA global-memory scratch (L2-resident).  This is the path for models the specialised kernels do not cover —
// today the reference's vec3 architecture (EncoderVec3 / DecoderVec3, python/VQVAE_v2.py:278-325; BASELINE config 4).
// It follows the same layer sequence and summation order as oracle/vqvae_oracle.c: every multiply-add is an fp32 FMA on
// the CUDA cores.  The 3x3x3 convolutions with >= 16 input channels (99 % of the MACs) run register-tiled from a
// shared-memory input chunk (conv_tiled_g); the small layers (Cin = 3, Cout = 3, the 1x1 projection) stay on the plain
// one-output-per-thread loop (conv_g).
#include "generic_model.cuh"
#include "leaf_ops.cuh"

namespace vqvdb {

namespace {

constexpr int kGThreads = 256;
constexpr int kConvSmemFloats = 8 * 1000;  // conv_tiled_g input chunk: 8 channels at (8+2)^3, or 32 at (4+2)^3 (6912)

// out[oc][pos] = b[oc] + sum_{ic,kd,kh,kw} in[ic][...] * wt[ic][tap][oc]   (zero padding 1, weights transposed)
// `in` and `out` are activation buffers written earlier in the same kernel: no __restrict__/read-only path for them.
__device__ void conv_g(const float* in, int cin, int S, const float* __restrict__ wt, const float* __restrict__ b,
                       int cout, int k, int stride, float* out) {
	const int pad = k == 1 ? 0 : 1;  // every k>1 conv of the model pads by 1; the 1x1 projection does not pad
	const int So = S / stride, nso = So * So * So, k3 = k * k * k, nsi = S * S * S;
	for (int idx = threadIdx.x; idx < cout * nso; idx += kGThreads) {
		const int oc = idx % cout, pos = idx / cout;  // oc fastest: neighbouring threads read neighbouring weights
		const int od = pos / (So * So), oh = (pos / So) % So, ow = pos % So;
		float acc = 0.f;
		for (int ic = 0; ic < cin; ++ic) {
			const float* ip = in + ic * nsi;
			const float* wp = wt + (size_t)ic * k3 * cout + oc;
			for (int kd = 0; kd < k; ++kd) {
				const int id = od * stride - pad + kd;
				if ((unsigned)id >= (unsigned)S) continue;
				for (int kh = 0; kh < k; ++kh) {
					const int ih = oh * stride - pad + kh;
					if ((unsigned)ih >= (unsigned)S) continue;
					for (int kw = 0; kw < k; ++kw) {
						const int iw = ow * stride - pad + kw;
						if ((unsigned)iw >= (unsigned)S) continue;
						acc = fmaf(ip[(id * S + ih) * S + iw], __ldg(wp + ((kd * k + kh) * k + kw) * cout), acc);
					}
				}
			}
		}
		out[oc * nso + pos] = acc + __ldg(b + oc);
	}
	__syncthreads();
}

// Register-tiled direct convolution: a work item is one output row (So voxels along w) x 4 output channels; the CTA
// walks the output channels in tiles of 256 items and, inside a tile, the input channels in chunks that are staged —
// zero halo included — in shared memory ([chunk][S+2]^3).  Weights are read as float4 through L1 (a warp's lanes share
// the address).  Accumulation order per output is ic, kd, kh, kw ascending exactly as in conv_g; the halo taps add an
// exact 0, so the result is bit-identical to conv_g's.
template <int S, int K, int STRIDE>
__device__ void conv_tiled_g(const float* in, int cin, const float* __restrict__ wt, const float* __restrict__ b, int cout,
                             float* out, float* s_in) {
	constexpr int So = S / STRIDE, HP = S + 2, ROWS = So * So, K3 = K * K * K;
	constexpr int CHUNK = S == 8 ? 8 : 32;            // 8 x 1000 or 32 x 216 floats
	constexpr int NIN = (So - 1) * STRIDE + K;        // input voxels one output row touches
	constexpr int OCG_PER_TILE = kGThreads / ROWS;    // 4 (8^3 outputs) or 16 (4^3 outputs) groups of 4 channels
	static_assert(NIN <= HP && kGThreads % ROWS == 0, "row fits the haloed input; a tile is exactly one item per thread");
	const int row = threadIdx.x % ROWS, ocg = threadIdx.x / ROWS, od = row / So, oh = row % So;
	const int nsi = S * S * S, nso = So * So * So;
	for (int oc0 = 0; oc0 < cout; oc0 += 4 * OCG_PER_TILE) {
		const int oc = oc0 + ocg * 4;
		const bool active = oc < cout;  // cout is a multiple of 4 (checked by the caller)
		float acc[4][So];
#pragma unroll
		for (int n = 0; n < 4; ++n)
#pragma unroll
			for (int j = 0; j < So; ++j) acc[n][j] = 0.f;
		for (int ic0 = 0; ic0 < cin; ic0 += CHUNK) {
			const int nch = min(CHUNK, cin - ic0);
			__syncthreads();  // previous chunk fully consumed
			for (int i = threadIdx.x; i < nch * HP * HP * HP; i += kGThreads) {
				const int c = i / (HP * HP * HP), r = i - c * (HP * HP * HP);
				const int dz = r / (HP * HP), hy = (r / HP) % HP, wx = r % HP;
				const bool inside = dz >= 1 && dz <= S && hy >= 1 && hy <= S && wx >= 1 && wx <= S;
				s_in[i] = inside ? in[(size_t)(ic0 + c) * nsi + ((dz - 1) * S + hy - 1) * S + wx - 1] : 0.f;
			}
			__syncthreads();
			if (active) {
#pragma unroll 1
				for (int c = 0; c < nch; ++c) {
					const float* wp = wt + ((size_t)(ic0 + c) * K3) * cout + oc;
#pragma unroll 1
					for (int kd = 0; kd < K; ++kd) {
#pragma unroll
						for (int kh = 0; kh < K; ++kh) {
							const float* ip = s_in + ((c * HP + od * STRIDE + kd) * HP + oh * STRIDE + kh) * HP;
							float x[HP];
#pragma unroll
							for (int j = 0; j < HP / 2; ++j) {
								const float2 v = *reinterpret_cast<const float2*>(ip + 2 * j);
								x[2 * j] = v.x;
								x[2 * j + 1] = v.y;
							}
#pragma unroll
							for (int kw = 0; kw < K; ++kw) {
								const float4 w4 = __ldg(reinterpret_cast<const float4*>(wp + (size_t)((kd * K + kh) * K + kw) * cout));
								const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
								for (int j = 0; j < So; ++j)
#pragma unroll
									for (int n = 0; n < 4; ++n) acc[n][j] = fmaf(x[j * STRIDE + kw], wv[n], acc[n][j]);
							}
						}
					}
				}
			}
		}
		if (active) {
#pragma unroll
			for (int n = 0; n < 4; ++n) {
				const float bias = __ldg(b + oc + n);
#pragma unroll
				for (int j = 0; j < So; ++j) out[(size_t)(oc + n) * nso + row * So + j] = acc[n][j] + bias;
			}
		}
	}
	__syncthreads();
}

// Picks the tiled kernel when the layer has the shape it is written for, the plain loop otherwise.
__device__ void conv_any_g(const float* in, int cin, int S, const float* __restrict__ wt, const float* __restrict__ b, int cout, int k,
                           int stride, float* out, float* s_in) {
	const bool tileable = cin >= 16 && cout % 4 == 0 && (reinterpret_cast<uintptr_t>(wt) & 15) == 0;
	if (tileable && k == 3 && stride == 1 && S == 8) conv_tiled_g<8, 3, 1>(in, cin, wt, b, cout, out, s_in);
	else if (tileable && k == 3 && stride == 1 && S == 4) conv_tiled_g<4, 3, 1>(in, cin, wt, b, cout, out, s_in);
	else if (tileable && k == 3 && stride == 2 && S == 8) conv_tiled_g<8, 3, 2>(in, cin, wt, b, cout, out, s_in);
	else if (tileable && k == 4 && stride == 2 && S == 8) conv_tiled_g<8, 4, 2>(in, cin, wt, b, cout, out, s_in);
	else conv_g(in, cin, S, wt, b, cout, k, stride, out);
}

// GroupNorm(groups, C) + optional ReLU, in place.  One warp per group, two-pass variance.
__device__ void gn_g(float* x, int C, int nsp, int groups, const float* __restrict__ gamma, const float* __restrict__ beta, bool relu) {
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, cg = C / groups, cnt = cg * nsp;
	for (int g = warp; g < groups; g += kGThreads / 32) {
		float* p = x + (size_t)g * cnt;
		float s = 0.f;
		for (int i = lane; i < cnt; i += 32) s += p[i];
		const float mean = warp_sum(s) / (float)cnt;
		float q = 0.f;
		for (int i = lane; i < cnt; i += 32) {
			const float d = p[i] - mean;
			q = fmaf(d, d, q);
		}
		const float rstd = 1.f / sqrtf(warp_sum(q) / (float)cnt + kGnEps);
		for (int i = lane; i < cnt; i += 32) {
			const int c = g * cg + i / nsp;
			float v = (p[i] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
			p[i] = relu ? relu_f(v) : v;
		}
	}
	__syncthreads();
}

// ResidualBlock (VQVAE_v2.py:204-210): x += 0.1 * conv2(relu(gn2(conv1(relu(gn1(x))))))
__device__ void res_g(float* x, int C, int S, const GenericRes& r, float* t0, float* t1, float* s_in) {
	const int n = C * S * S * S;
	for (int i = threadIdx.x; i < n; i += kGThreads) t0[i] = x[i];
	__syncthreads();
	gn_g(t0, C, S * S * S, 8, r.gn1_w, r.gn1_b, true);
	conv_any_g(t0, C, S, r.c1_w, r.c1_b, C, 3, 1, t1, s_in);
	gn_g(t1, C, S * S * S, 8, r.gn2_w, r.gn2_b, true);
	conv_any_g(t1, C, S, r.c2_w, r.c2_b, C, 3, 1, t0, s_in);
	for (int i = threadIdx.x; i < n; i += kGThreads) x[i] = x[i] + kResScale * t0[i];
	__syncthreads();
}

// ChannelAttention (VQVAE_v2.py:213-228), C <= 256, R <= 64
__device__ void attn_g(float* x, int C, int nsp, const float* __restrict__ fc0, const float* __restrict__ fc2, int R, float* s_tmp) {
	float* s_mean = s_tmp;
	float* s_hid = s_tmp + 256;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (int c = warp; c < C; c += kGThreads / 32) {
		float s = 0.f;
		for (int i = lane; i < nsp; i += 32) s += x[c * nsp + i];
		s = warp_sum(s);
		if (lane == 0) s_mean[c] = s / (float)nsp;
	}
	__syncthreads();
	for (int j = warp; j < R; j += kGThreads / 32) {
		float s = 0.f;
		for (int c = lane; c < C; c += 32) s = fmaf(__ldg(fc0 + j * C + c), s_mean[c], s);
		s = warp_sum(s);
		if (lane == 0) s_hid[j] = relu_f(s);
	}
	__syncthreads();
	for (int c = threadIdx.x; c < C; c += kGThreads) {
		float s = 0.f;
		for (int j = 0; j < R; ++j) s = fmaf(__ldg(fc2 + c * R + j), s_hid[j], s);
		s_mean[c] = sigmoid_f(s);
	}
	__syncthreads();
	for (int i = threadIdx.x; i < C * nsp; i += kGThreads) x[i] *= s_mean[i / nsp];
	__syncthreads();
}

__global__ void __launch_bounds__(kGThreads)
encode_generic_kernel(const GenericModel m, const float* __restrict__ leaves, int64_t n_leaves, uint8_t* __restrict__ indices,
                      float* __restrict__ scratch, float* __restrict__ down_out) {
	__shared__ float s_tmp[320];
	__shared__ float s_best[4 * 64];
	__shared__ int s_bi[4 * 64];
	__shared__ __align__(16) float s_in[kConvSmemFloats];
	const size_t buf = 64 * 512;  // largest activation: 64 channels at 8^3
	float* s0 = scratch + (size_t)blockIdx.x * (3 * buf + 128 * 64);
	float* s1 = s0 + buf;
	float* s2 = s1 + buf;
	float* z = s2 + buf;
	const size_t leaf_sz = (size_t)m.cin * 512;
	for (int64_t leaf = blockIdx.x; leaf < n_leaves; leaf += gridDim.x) {
		conv_g(leaves + leaf * leaf_sz, m.cin, 8, m.e_pre_w, m.e_pre_b, m.e_c0, 3, 1, s0);
		gn_g(s0, m.e_c0, 512, m.e_gn0, m.e_gn_w, m.e_gn_b, true);
		res_g(s0, m.e_c0, 8, m.e_res0, s1, s2, s_in);
		if (down_out) {  // front half only: the stride-2 conv's output [e_c1][64] is the result (encode_tc128.cu continues)
			conv_any_g(s0, m.e_c0, 8, m.e_down_w, m.e_down_b, m.e_c1, m.e_down_k, 2, down_out + leaf * (int64_t)(m.e_c1 * 64), s_in);
			continue;
		}
		conv_any_g(s0, m.e_c0, 8, m.e_down_w, m.e_down_b, m.e_c1, m.e_down_k, 2, s1, s_in);
		for (int r = 0; r < m.e_nres; ++r) res_g(s1, m.e_c1, 4, m.e_res[r], s0, s2, s_in);
		attn_g(s1, m.e_c1, 64, m.e_fc0, m.e_fc2, m.e_red, s_tmp);
		conv_g(s1, m.e_c1, 4, m.e_proj_w, m.e_proj_b, m.D, 1, 1, z);
		// VQ (save_for_inference.py:55-61): thread = (position, quarter of the codes); fp32, sequential in d
		{
			const int p = threadIdx.x & 63, q = threadIdx.x >> 6, kq = m.K / 4;
			float zz = 0.f;
			for (int d = 0; d < m.D; ++d) zz = fmaf(z[d * 64 + p], z[d * 64 + p], zz);
			float best = INFINITY;
			int bi = 0;
			for (int k = q * kq; k < (q + 1) * kq; ++k) {
				float dot = 0.f;
				for (int d = 0; d < m.D; ++d) dot = fmaf(z[d * 64 + p], __ldg(m.emb + (size_t)k * m.D + d), dot);
				const float dist = (zz + __ldg(m.emb_sq + k)) - 2.f * dot;
				if (dist < best) {
					best = dist;
					bi = k;
				}
			}
			s_best[q * 64 + p] = best;
			s_bi[q * 64 + p] = bi;
		}
		__syncthreads();
		if (threadIdx.x < 64) {
			float best = s_best[threadIdx.x];
			int bi = s_bi[threadIdx.x];
			for (int q = 1; q < 4; ++q)
				if (s_best[q * 64 + threadIdx.x] < best) {
					best = s_best[q * 64 + threadIdx.x];
					bi = s_bi[q * 64 + threadIdx.x];
				}
			indices[leaf * 64 + threadIdx.x] = (uint8_t)bi;
		}
		__syncthreads();
	}
}

__global__ void __launch_bounds__(kGThreads)
decode_generic_kernel(const GenericModel m, const uint8_t* __restrict__ indices, int64_t n_leaves, float* __restrict__ voxels,
                      float* __restrict__ scratch) {
	__shared__ float s_tmp[320];
	__shared__ __align__(16) float s_in[kConvSmemFloats];
	const size_t buf = 256 * 64;  // largest activation: up_conv output, 256 channels at 4^3 (= 32 channels at 8^3)
	float* s0 = scratch + (size_t)blockIdx.x * (3 * buf);
	float* s1 = s0 + buf;
	float* s2 = s1 + buf;
	const size_t leaf_sz = (size_t)m.cin * 512;
	const int C = m.d_c;
	for (int64_t leaf = blockIdx.x; leaf < n_leaves; leaf += gridDim.x) {
		for (int i = threadIdx.x; i < m.D * 64; i += kGThreads) {  // F.embedding + permute: q[d][p] = emb[idx[p]][d]
			const int d = i / 64, p = i % 64;
			s0[i] = __ldg(m.emb + (size_t)indices[leaf * 64 + p] * m.D + d);
		}
		__syncthreads();
		conv_any_g(s0, m.D, 4, m.d_stem_w, m.d_stem_b, C, 3, 1, s1, s_in);
		gn_g(s1, C, 64, 8, m.d_gn_w, m.d_gn_b, true);
		for (int r = 0; r < m.d_nres; ++r) res_g(s1, C, 4, m.d_res[r], s0, s2, s_in);
		attn_g(s1, C, 64, m.d_fc0, m.d_fc2, m.d_red, s_tmp);
		conv_any_g(s1, C, 4, m.d_up_w, m.d_up_b, 256, 3, 1, s0, s_in);
		for (int i = threadIdx.x; i < 256 * 64; i += kGThreads) {  // PixelShuffle3D(2), VQVAE_v2.py:177-187
			const int c = i / 64, p = i % 64, d = p / 16, h = (p / 4) % 4, w = p % 4;
			const int oc = c >> 3, rd = (c >> 2) & 1, rh = (c >> 1) & 1, rw = c & 1;
			s2[((oc * 8 + 2 * d + rd) * 8 + 2 * h + rh) * 8 + 2 * w + rw] = s0[i];
		}
		__syncthreads();
		conv_g(s2, 32, 8, m.d_fin_w, m.d_fin_b, m.cin, 3, 1, s0);
		for (int i = threadIdx.x; i < (int)leaf_sz; i += kGThreads)
			voxels[leaf * leaf_sz + i] = m.cin == 1 ? sigmoid_f(s0[i]) : tanhf(s0[i]);  // VQVAE_v2.py:275 / :325
		__syncthreads();
	}
}

}  // namespace

size_t generic_scratch_floats(int grid) { return (size_t)grid * (3 * 64 * 512 + 128 * 64); }

cudaError_t launch_encode_generic(const GenericModel& m, const float* leaves, int64_t n, uint8_t* indices, float* scratch, int grid,
                                  cudaStream_t stream) {
	if (n <= 0) return cudaSuccess;
	const int g = (int)(n < grid ? n : grid);
	encode_generic_kernel<<<g, kGThreads, 0, stream>>>(m, leaves, n, indices, scratch, nullptr);
	return cudaGetLastError();
}

cudaError_t launch_encode_generic_front(const GenericModel& m, const float* leaves, int64_t n, float* down_out, float* scratch, int grid,
                                        cudaStream_t stream) {
	if (n <= 0) return cudaSuccess;
	const int g = (int)(n < grid ? n : grid);
	encode_generic_kernel<<<g, kGThreads, 0, stream>>>(m, leaves, n, nullptr, scratch, down_out);
	return cudaGetLastError();
}

cudaError_t launch_decode_generic(const GenericModel& m, const uint8_t* indices, int64_t n, float* voxels, float* scratch, int grid,
                                  cudaStream_t stream) {
	if (n <= 0) return cudaSuccess;
	const int g = (int)(n < grid ? n : grid);
	decode_generic_kernel<<<g, kGThreads, 0, stream>>>(m, indices, n, voxels, scratch);
	return cudaGetLastError();
}

}  // namespace vqvdb
