/* TEST INFRASTRUCTURE — see vqvae_oracle.h.
 *
 * Plain-C fp32 restatement of the reference's inference arithmetic.  The
 * reference delegates it to LibTorch/ONNX Runtime (third-party, not under
 * /root/reference; call sites TorchBackend.cpp:149,180); its semantics are
 * fully specified by python/VQVAE_v2.py + the shipped weights, which is what
 * each function below cites.  Compiled with -ffp-contract=off: every multiply
 * and add is a separate IEEE fp32 operation, in the loop order written here.
 */
#define _GNU_SOURCE
#include "vqvae_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define MAX_TENSORS 96
#define MAX_RES 2

typedef struct {
	char name[96];
	int ndim;
	int dims[5];
	const float* data;
} vqo_tensor;

typedef struct {
	const float *gn1_w, *gn1_b, *c1_w, *c1_b, *gn2_w, *gn2_b, *c2_w, *c2_b;
} vqo_res;

struct vqo_model {
	unsigned char* blob;
	int n_tensors;
	vqo_tensor t[MAX_TENSORS];
	int cin, D, K;
	/* encoder */
	int e_c0;        /* channels at 8^3 (16 float / 64 vec3) */
	int e_c1;        /* channels at 4^3 (32 / 128) */
	int e_gn0;       /* groups of pre.1 (4 / 8) */
	int e_down_k;    /* 4 / 3 */
	int e_nres;      /* 1 / 2 */
	int e_red;       /* attention hidden width */
	const float *e_pre_w, *e_pre_b, *e_gn_w, *e_gn_b;
	vqo_res e_res0;
	const float *e_down_w, *e_down_b;
	vqo_res e_res[MAX_RES];
	const float *e_fc0, *e_fc2, *e_proj_w, *e_proj_b;
	const float* emb;
	float* emb_sq; /* sum(embedding**2, dim=1) */
	/* decoder */
	int d_c;    /* stem width (64 / 128) */
	int d_nres; /* 1 / 2 */
	int d_red;
	const float *d_stem_w, *d_stem_b, *d_gn_w, *d_gn_b;
	vqo_res d_res[MAX_RES];
	const float *d_fc0, *d_fc2, *d_up_w, *d_up_b, *d_fin_w, *d_fin_b;
};

static const vqo_tensor* find(const vqo_model* m, const char* name) {
	for (int i = 0; i < m->n_tensors; ++i)
		if (strcmp(m->t[i].name, name) == 0) return &m->t[i];
	return NULL;
}
static const float* need(const vqo_model* m, const char* name, int* ok) {
	const vqo_tensor* t = find(m, name);
	if (!t) {
		fprintf(stderr, "vqvae_oracle: missing tensor %s\n", name);
		*ok = 0;
		return NULL;
	}
	return t->data;
}
static void load_res(const vqo_model* m, const char* prefix, vqo_res* r, int* ok) {
	char nm[128];
#define G(field, suffix)                              \
	snprintf(nm, sizeof nm, "%s.%s", prefix, suffix); \
	r->field = need(m, nm, ok);
	G(gn1_w, "gn1.weight") G(gn1_b, "gn1.bias") G(c1_w, "conv1.weight") G(c1_b, "conv1.bias")
	G(gn2_w, "gn2.weight") G(gn2_b, "gn2.bias") G(c2_w, "conv2.weight") G(c2_b, "conv2.bias")
#undef G
}

vqo_model* vqo_load(const char* path) {
	FILE* f = fopen(path, "rb");
	if (!f) return NULL;
	fseek(f, 0, SEEK_END);
	long sz = ftell(f);
	fseek(f, 0, SEEK_SET);
	unsigned char* blob = (unsigned char*)malloc((size_t)sz);
	if (!blob || fread(blob, 1, (size_t)sz, f) != (size_t)sz) {
		fclose(f);
		free(blob);
		return NULL;
	}
	fclose(f);
	if (sz < 24 || memcmp(blob, "VQVDBW01", 8) != 0) {
		free(blob);
		return NULL;
	}
	vqo_model* m = (vqo_model*)calloc(1, sizeof *m);
	m->blob = blob;
	size_t pos = 8;
	unsigned n, cin, D, K;
	memcpy(&n, blob + pos, 4);
	memcpy(&cin, blob + pos + 4, 4);
	memcpy(&D, blob + pos + 8, 4);
	memcpy(&K, blob + pos + 12, 4);
	pos += 16;
	m->cin = (int)cin;
	m->D = (int)D;
	m->K = (int)K;
	if (n > MAX_TENSORS) goto fail;
	unsigned long long offs[MAX_TENSORS];
	for (unsigned i = 0; i < n; ++i) {
		unsigned ln, nd;
		memcpy(&ln, blob + pos, 4);
		pos += 4;
		if (ln >= sizeof m->t[i].name) goto fail;
		memcpy(m->t[i].name, blob + pos, ln);
		m->t[i].name[ln] = 0;
		pos += ln;
		memcpy(&nd, blob + pos, 4);
		pos += 4;
		if (nd > 5) goto fail;
		m->t[i].ndim = (int)nd;
		for (unsigned d = 0; d < nd; ++d) {
			unsigned v;
			memcpy(&v, blob + pos, 4);
			pos += 4;
			m->t[i].dims[d] = (int)v;
		}
		unsigned long long nb;
		memcpy(&offs[i], blob + pos, 8);
		memcpy(&nb, blob + pos + 8, 8);
		pos += 16;
	}
	pos += 8; /* payload_bytes */
	pos = (pos + 63) & ~(size_t)63;
	m->n_tensors = (int)n;
	for (unsigned i = 0; i < n; ++i) m->t[i].data = (const float*)(blob + pos + offs[i]);

	int ok = 1;
	const int vec3 = (m->cin != 1);
	m->e_pre_w = need(m, "encoder.pre.0.weight", &ok);
	m->e_pre_b = need(m, "encoder.pre.0.bias", &ok);
	m->e_gn_w = need(m, "encoder.pre.1.weight", &ok);
	m->e_gn_b = need(m, "encoder.pre.1.bias", &ok);
	load_res(m, "encoder.pre.3", &m->e_res0, &ok);
	const char* down = vec3 ? "encoder.down1" : "encoder.down"; /* VQVAE_v2.py:240 vs :288 */
	char nm[128];
	snprintf(nm, sizeof nm, "%s.weight", down);
	m->e_down_w = need(m, nm, &ok);
	const vqo_tensor* dw = find(m, nm);
	snprintf(nm, sizeof nm, "%s.bias", down);
	m->e_down_b = need(m, nm, &ok);
	if (!ok || !dw) goto fail;
	m->e_c0 = dw->dims[1];
	m->e_c1 = dw->dims[0];
	m->e_down_k = dw->dims[2];
	m->e_gn0 = vec3 ? 8 : 4; /* GroupNorm(4,16) VQVAE_v2.py:236 ; GroupNorm(8,64) :283 */
	m->e_nres = vec3 ? 2 : 1;
	for (int r = 0; r < m->e_nres; ++r) {
		snprintf(nm, sizeof nm, "encoder.res_stack.%d", r);
		load_res(m, nm, &m->e_res[r], &ok);
	}
	m->e_fc0 = need(m, "encoder.attn.fc.0.weight", &ok);
	m->e_fc2 = need(m, "encoder.attn.fc.2.weight", &ok);
	if (!ok) goto fail;
	m->e_red = find(m, "encoder.attn.fc.0.weight")->dims[0];
	m->e_proj_w = need(m, "encoder.proj.weight", &ok);
	m->e_proj_b = need(m, "encoder.proj.bias", &ok);
	m->emb = need(m, "quantizer.embedding", &ok);
	m->d_stem_w = need(m, "decoder.stem.0.weight", &ok);
	m->d_stem_b = need(m, "decoder.stem.0.bias", &ok);
	m->d_gn_w = need(m, "decoder.stem.1.weight", &ok);
	m->d_gn_b = need(m, "decoder.stem.1.bias", &ok);
	if (!ok) goto fail;
	m->d_c = find(m, "decoder.stem.0.weight")->dims[0];
	m->d_nres = vec3 ? 2 : 1;
	for (int r = 0; r < m->d_nres; ++r) {
		snprintf(nm, sizeof nm, "decoder.res_stack.%d", r);
		load_res(m, nm, &m->d_res[r], &ok);
	}
	m->d_fc0 = need(m, "decoder.attn.fc.0.weight", &ok);
	m->d_fc2 = need(m, "decoder.attn.fc.2.weight", &ok);
	if (!ok) goto fail;
	m->d_red = find(m, "decoder.attn.fc.0.weight")->dims[0];
	m->d_up_w = need(m, "decoder.up_conv.weight", &ok);
	m->d_up_b = need(m, "decoder.up_conv.bias", &ok);
	m->d_fin_w = need(m, "decoder.final.weight", &ok);
	m->d_fin_b = need(m, "decoder.final.bias", &ok);
	if (!ok) goto fail;

	/* torch.sum(self.embedding ** 2, dim=1)   save_for_inference.py:58 */
	m->emb_sq = (float*)malloc(sizeof(float) * (size_t)m->K);
	for (int k = 0; k < m->K; ++k) {
		float s = 0.f;
		for (int d = 0; d < m->D; ++d) s += m->emb[k * m->D + d] * m->emb[k * m->D + d];
		m->emb_sq[k] = s;
	}
	return m;
fail:
	vqo_free(m);
	return NULL;
}

void vqo_free(vqo_model* m) {
	if (!m) return;
	free(m->emb_sq);
	free(m->blob);
	free(m);
}

int vqo_in_channels(const vqo_model* m) { return m->cin; }

/* ---- primitive ops on one leaf, tensors are [C][S][S][S] row-major ---- */

/* nn.Conv3d(cin, cout, k, stride, pad, bias=True), zero padding. */
static void conv3d(const float* in, int cin, int S, const float* w, const float* b, int cout, int k, int stride,
                   int pad, float* out) {
	const int So = (S + 2 * pad - k) / stride + 1;
	for (int oc = 0; oc < cout; ++oc)
		for (int od = 0; od < So; ++od)
			for (int oh = 0; oh < So; ++oh)
				for (int ow = 0; ow < So; ++ow) {
					float acc = 0.f;
					for (int ic = 0; ic < cin; ++ic)
						for (int kd = 0; kd < k; ++kd) {
							const int id = od * stride - pad + kd;
							if (id < 0 || id >= S) continue;
							for (int kh = 0; kh < k; ++kh) {
								const int ih = oh * stride - pad + kh;
								if (ih < 0 || ih >= S) continue;
								for (int kw = 0; kw < k; ++kw) {
									const int iw = ow * stride - pad + kw;
									if (iw < 0 || iw >= S) continue;
									acc += in[((ic * S + id) * S + ih) * S + iw] *
									       w[(((oc * cin + ic) * k + kd) * k + kh) * k + kw];
								}
							}
						}
					out[((oc * So + od) * So + oh) * So + ow] = acc + b[oc];
				}
}

/* nn.GroupNorm(groups, C), eps=1e-5, affine, biased variance over (C/groups x S^3); optional ReLU. */
static void groupnorm(float* x, int C, int n_sp, int groups, const float* gamma, const float* beta, int relu) {
	const int cg = C / groups;
	for (int g = 0; g < groups; ++g) {
		float* p = x + (size_t)g * cg * n_sp;
		const int cnt = cg * n_sp;
		double s = 0.0, ss = 0.0;
		for (int i = 0; i < cnt; ++i) {
			s += p[i];
			ss += (double)p[i] * p[i];
		}
		const double mean = s / cnt;
		double var = ss / cnt - mean * mean;
		if (var < 0) var = 0;
		const float rstd = (float)(1.0 / sqrt(var + 1e-5));
		const float fmean = (float)mean;
		for (int c = 0; c < cg; ++c) {
			const float ga = gamma[g * cg + c], be = beta[g * cg + c];
			for (int i = 0; i < n_sp; ++i) {
				float v = (p[c * n_sp + i] - fmean) * rstd * ga + be;
				if (relu && v < 0.f) v = 0.f;
				p[c * n_sp + i] = v;
			}
		}
	}
}

/* ResidualBlock.forward, VQVAE_v2.py:204-210: x + 0.1*conv2(relu(gn2(conv1(relu(gn1(x)))))), groups=8. */
static void resblock_tap(float* x, int C, int S, const vqo_res* r, float* t0, float* t1, float* conv1_tap) {
	const int n_sp = S * S * S;
	memcpy(t0, x, sizeof(float) * (size_t)C * n_sp);
	groupnorm(t0, C, n_sp, 8, r->gn1_w, r->gn1_b, 1);
	conv3d(t0, C, S, r->c1_w, r->c1_b, C, 3, 1, 1, t1);
	if (conv1_tap) memcpy(conv1_tap, t1, sizeof(float) * (size_t)C * n_sp);
	groupnorm(t1, C, n_sp, 8, r->gn2_w, r->gn2_b, 1);
	conv3d(t1, C, S, r->c2_w, r->c2_b, C, 3, 1, 1, t0);
	for (int i = 0; i < C * n_sp; ++i) x[i] = x[i] + 0.1f * t0[i];
}
static void resblock(float* x, int C, int S, const vqo_res* r, float* t0, float* t1) { resblock_tap(x, C, S, r, t0, t1, NULL); }

static float sigmoidf(float v) { return 1.f / (1.f + expf(-v)); }

/* ChannelAttention.forward, VQVAE_v2.py:224-228; both Linear layers bias-free (:218,220). */
static void channel_attention(float* x, int C, int n_sp, const float* fc0, const float* fc2, int red) {
	float mean[256], hid[64];
	for (int c = 0; c < C; ++c) {
		float s = 0.f;
		for (int i = 0; i < n_sp; ++i) s += x[c * n_sp + i];
		mean[c] = s / (float)n_sp;
	}
	for (int j = 0; j < red; ++j) {
		float s = 0.f;
		for (int c = 0; c < C; ++c) s += fc0[j * C + c] * mean[c];
		hid[j] = s > 0.f ? s : 0.f;
	}
	for (int c = 0; c < C; ++c) {
		float s = 0.f;
		for (int j = 0; j < red; ++j) s += fc2[c * red + j] * hid[j];
		const float y = sigmoidf(s);
		for (int i = 0; i < n_sp; ++i) x[c * n_sp + i] *= y;
	}
}

/* EncoderFloat.forward VQVAE_v2.py:245-250 / EncoderVec3.forward :293-299.  z is [D][4][4][4]. */
/* tap_stage (kernel bring-up): 0 = pre (GN+ReLU) [c0][8^3], 1 = first res block out [c0][8^3], 2 = down out [c1][4^3],
 * 3 = res stack out, 4 = after attention, 6 = first res block's conv1 out (+bias) [c0][8^3], 7 = res_stack.0 conv1 out [c1][4^3]. */
static void encoder_forward_tap(const vqo_model* m, const float* leaf, float* z, float* s0, float* s1, float* s2, int tap_stage,
                                float* tap) {
	const int c0 = m->e_c0, c1 = m->e_c1;
	conv3d(leaf, m->cin, 8, m->e_pre_w, m->e_pre_b, c0, 3, 1, 1, s0);
	groupnorm(s0, c0, 512, m->e_gn0, m->e_gn_w, m->e_gn_b, 1);
	if (tap_stage == 0) memcpy(tap, s0, sizeof(float) * (size_t)c0 * 512);
	resblock_tap(s0, c0, 8, &m->e_res0, s1, s2, tap_stage == 6 ? tap : NULL);
	if (tap_stage == 1) memcpy(tap, s0, sizeof(float) * (size_t)c0 * 512);
	conv3d(s0, c0, 8, m->e_down_w, m->e_down_b, c1, m->e_down_k, 2, 1, s1); /* -> [c1][4][4][4] */
	if (tap_stage == 2) memcpy(tap, s1, sizeof(float) * (size_t)c1 * 64);
	for (int r = 0; r < m->e_nres; ++r) resblock_tap(s1, c1, 4, &m->e_res[r], s0, s2, (tap_stage == 7 && r == 0) ? tap : NULL);
	if (tap_stage == 3) memcpy(tap, s1, sizeof(float) * (size_t)c1 * 64);
	channel_attention(s1, c1, 64, m->e_fc0, m->e_fc2, m->e_red);
	if (tap_stage == 4) memcpy(tap, s1, sizeof(float) * (size_t)c1 * 64);
	conv3d(s1, c1, 4, m->e_proj_w, m->e_proj_b, m->D, 1, 1, 0, z);
}
static void encoder_forward(const vqo_model* m, const float* leaf, float* z, float* s0, float* s1, float* s2) {
	encoder_forward_tap(m, leaf, z, s0, s1, s2, -1, NULL);
}

/* InferenceVectorQuantizer.get_indices, save_for_inference.py:55-61:
 * argmin_k( sum(z^2) + sum(e_k^2) - 2 * z.e_k ), first minimum wins (torch.argmin). */
static void quantize(const vqo_model* m, const float* z, uint8_t* idx, float* margin) {
	const int D = m->D, K = m->K;
	for (int p = 0; p < 64; ++p) {
		float zz = 0.f;
		for (int d = 0; d < D; ++d) zz += z[d * 64 + p] * z[d * 64 + p];
		float best = INFINITY, second = INFINITY;
		int bi = 0;
		for (int k = 0; k < K; ++k) {
			float dot = 0.f;
			for (int d = 0; d < D; ++d) dot += z[d * 64 + p] * m->emb[k * D + d];
			const float dist = (zz + m->emb_sq[k]) - 2.f * dot;
			if (dist < best) {
				second = best;
				best = dist;
				bi = k;
			} else if (dist < second) {
				second = dist;
			}
		}
		idx[p] = (uint8_t)bi; /* int64 -> uint8 cast: TorchBackend.cpp:150 */
		if (margin) margin[p] = second - best;
	}
}

/* InferenceVQVAE.decode, save_for_inference.py:91-104 + DecoderFloat.forward VQVAE_v2.py:270-275. */
static void decoder_forward(const vqo_model* m, const uint8_t* idx, float* out, float* s0, float* s1, float* s2,
                            int tap_stage, float* tap) {
	const int D = m->D, C = m->d_c;
	/* F.embedding + permute(0,4,1,2,3): q[d][p] = embedding[idx[p]][d] */
	for (int p = 0; p < 64; ++p)
		for (int d = 0; d < D; ++d) s0[d * 64 + p] = m->emb[(int)idx[p] * D + d];
	conv3d(s0, D, 4, m->d_stem_w, m->d_stem_b, C, 3, 1, 1, s1);
	groupnorm(s1, C, 64, 8, m->d_gn_w, m->d_gn_b, 1);
	if (tap_stage == 0) {
		memcpy(tap, s1, sizeof(float) * (size_t)C * 64);
		return;
	}
	for (int r = 0; r < m->d_nres; ++r) resblock(s1, C, 4, &m->d_res[r], s0, s2);
	if (tap_stage == 1) {
		memcpy(tap, s1, sizeof(float) * (size_t)C * 64);
		return;
	}
	channel_attention(s1, C, 64, m->d_fc0, m->d_fc2, m->d_red);
	if (tap_stage == 2) {
		memcpy(tap, s1, sizeof(float) * (size_t)C * 64);
		return;
	}
	conv3d(s1, C, 4, m->d_up_w, m->d_up_b, 256, 3, 1, 1, s0); /* [256][4][4][4] */
	/* PixelShuffle3D(2), VQVAE_v2.py:177-187: out[oc,2d+rd,2h+rh,2w+rw] = in[oc*8+rd*4+rh*2+rw,d,h,w] */
	for (int oc = 0; oc < 32; ++oc)
		for (int d = 0; d < 4; ++d)
			for (int h = 0; h < 4; ++h)
				for (int w = 0; w < 4; ++w)
					for (int rd = 0; rd < 2; ++rd)
						for (int rh = 0; rh < 2; ++rh)
							for (int rw = 0; rw < 2; ++rw)
								s2[((oc * 8 + 2 * d + rd) * 8 + 2 * h + rh) * 8 + 2 * w + rw] =
								    s0[(oc * 8 + rd * 4 + rh * 2 + rw) * 64 + (d * 4 + h) * 4 + w];
	if (tap_stage == 3) {
		memcpy(tap, s2, sizeof(float) * 32 * 512);
		return;
	}
	conv3d(s2, 32, 8, m->d_fin_w, m->d_fin_b, m->cin, 3, 1, 1, out);
	const int n = m->cin * 512;
	if (m->cin == 1)
		for (int i = 0; i < n; ++i) out[i] = sigmoidf(out[i]); /* VQVAE_v2.py:275 */
	else
		for (int i = 0; i < n; ++i) out[i] = tanhf(out[i]); /* VQVAE_v2.py:325 */
}

#define SCRATCH_FLOATS (256 * 512)

/* ---- batch loops: leaves are independent, so a small pthread work-queue shares them out ---- */

static int g_threads = 0; /* 0 = number of online cores */

int vqo_set_threads(int n) {
	if (n > 0) g_threads = n;
	if (g_threads <= 0) {
		long c = sysconf(_SC_NPROCESSORS_ONLN);
		g_threads = c > 0 ? (int)c : 1;
	}
	return g_threads;
}

typedef struct {
	const vqo_model* m;
	const float* leaves;
	const uint8_t* idx_in;
	uint8_t* idx_out;
	float* margins;
	float* fout; /* latents / voxels */
	float* tap;
	int mode;    /* 0 encode, 1 latents, 2 decode, 3 encoder tap */
	int stage;
	int64_t n;
	int64_t next; /* atomic cursor */
} vqo_job;

static void* worker(void* arg) {
	vqo_job* j = (vqo_job*)arg;
	const vqo_model* m = j->m;
	const size_t leaf_sz = (size_t)m->cin * 512;
	const size_t tap_sz = j->stage == 3 ? 32 * 512 : (size_t)m->d_c * 64;
	float* s = (float*)malloc(sizeof(float) * (size_t)(3 * SCRATCH_FLOATS + m->D * 64 + leaf_sz));
	float* z = s + 3 * SCRATCH_FLOATS;
	float* vox = z + m->D * 64;
	for (;;) {
		const int64_t lo = __atomic_fetch_add(&j->next, 4, __ATOMIC_RELAXED);
		if (lo >= j->n) break;
		const int64_t hi = lo + 4 < j->n ? lo + 4 : j->n;
		for (int64_t i = lo; i < hi; ++i) {
			if (j->mode == 0) {
				encoder_forward(m, j->leaves + i * leaf_sz, z, s, s + SCRATCH_FLOATS, s + 2 * SCRATCH_FLOATS);
				quantize(m, z, j->idx_out + i * 64, j->margins ? j->margins + i * 64 : NULL);
			} else if (j->mode == 1) {
				encoder_forward(m, j->leaves + i * leaf_sz, j->fout + i * (size_t)m->D * 64, s, s + SCRATCH_FLOATS,
				                s + 2 * SCRATCH_FLOATS);
			} else if (j->mode == 3) {
				const int big = (j->stage == 0 || j->stage == 1 || j->stage == 6);
				const size_t esz = big ? (size_t)m->e_c0 * 512 : (size_t)m->e_c1 * 64;
				encoder_forward_tap(m, j->leaves + i * leaf_sz, z, s, s + SCRATCH_FLOATS, s + 2 * SCRATCH_FLOATS, j->stage,
				                    j->tap + i * esz);
			} else {
				decoder_forward(m, j->idx_in + i * 64, j->fout ? j->fout + i * leaf_sz : vox, s, s + SCRATCH_FLOATS,
				                s + 2 * SCRATCH_FLOATS, j->stage, j->tap ? j->tap + i * tap_sz : NULL);
			}
		}
	}
	free(s);
	return NULL;
}

static int run_job(vqo_job* j) {
	if (!j->m || j->n < 0) return -1;
	int nt = vqo_set_threads(0);
	if ((int64_t)nt > (j->n + 3) / 4) nt = (int)((j->n + 3) / 4);
	if (nt <= 1) {
		worker(j);
		return 0;
	}
	pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nt);
	int started = 0;
	for (int t = 0; t < nt - 1; ++t)
		if (pthread_create(&th[started], NULL, worker, j) == 0) ++started;
	worker(j);
	for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
	free(th);
	return 0;
}

int vqo_encode(const vqo_model* m, const float* leaves, int64_t n, uint8_t* indices, float* margins) {
	vqo_job j = {m, leaves, NULL, indices, margins, NULL, NULL, 0, -1, n, 0};
	return run_job(&j);
}

int vqo_encode_latents(const vqo_model* m, const float* leaves, int64_t n, float* zout) {
	vqo_job j = {m, leaves, NULL, NULL, NULL, zout, NULL, 1, -1, n, 0};
	return run_job(&j);
}

int vqo_encode_tap(const vqo_model* m, const float* leaves, int64_t n, int stage, float* out) {
	if (!(stage >= 0 && stage <= 4) && stage != 6 && stage != 7) return -1;
	vqo_job j = {m, leaves, NULL, NULL, NULL, NULL, out, 3, stage, n, 0};
	return run_job(&j);
}

int vqo_decode(const vqo_model* m, const uint8_t* indices, int64_t n, float* voxels) {
	vqo_job j = {m, NULL, indices, NULL, NULL, voxels, NULL, 2, -1, n, 0};
	return run_job(&j);
}

int vqo_decode_tap(const vqo_model* m, const uint8_t* indices, int64_t n, int stage, float* out) {
	if (stage < 0 || stage > 3) return -1;
	vqo_job j = {m, NULL, indices, NULL, NULL, NULL, out, 2, stage, n, 0};
	return run_job(&j);
}
