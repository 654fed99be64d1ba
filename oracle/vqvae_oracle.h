/* TEST INFRASTRUCTURE — CPU restatement of the reference VQ-VAE leaf codec.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
 * this library, and only as the checker.  The product (libvqvdb_b200.so) never
 * links or calls it.
 *
 * Parity pin: the reference has no tests or golden vectors for this path
 * (SURVEY §4).  This restatement is pinned against outputs of the reference
 * itself — the shipped TorchScript blob run on CPU in the dev container —
 * committed under tests/golden/ by tools/make_goldens.py, and (where
 * oracle/_ref is present) against the reference's own compiled LibTorch
 * backend.  See tests/test_oracle.py.
 */
#ifndef VQVAE_ORACLE_H
#define VQVAE_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vqo_model vqo_model;

/* Loads a VQVDBW01 weight pack (tools/weights_pack.py). NULL on failure. */
vqo_model* vqo_load(const char* pack_path);
void vqo_free(vqo_model* m);
int vqo_in_channels(const vqo_model* m);
int vqo_set_threads(int n); /* OpenMP threads used by the batch loops; returns the value in effect */

/* leaves [n, C, 8, 8, 8] fp32 -> indices [n, 4, 4, 4] uint8.
 * margins (nullable) [n, 64]: second-best minus best distance per latent, used
 * by the parity tests to classify an index mismatch as a near-tie. */
int vqo_encode(const vqo_model* m, const float* leaves, int64_t n, uint8_t* indices, float* margins);

/* indices [n, 4, 4, 4] uint8 -> voxels [n, C, 8, 8, 8] fp32. */
int vqo_decode(const vqo_model* m, const uint8_t* indices, int64_t n, float* voxels);

/* Encoder output before quantisation, [n, D, 4, 4, 4] fp32 (debug tap). */
int vqo_encode_latents(const vqo_model* m, const float* leaves, int64_t n, float* z);

/* Encoder activation taps for kernel bring-up: stage 0 = pre (GN+ReLU) [c0,8,8,8], 1 = first residual block out
 * [c0,8,8,8], 2 = down out [c1,4,4,4], 3 = residual stack out, 4 = after attention, 6 = first residual block's conv1
 * out [c0,8,8,8], 7 = res_stack.0 conv1 out [c1,4,4,4]. */
int vqo_encode_tap(const vqo_model* m, const float* leaves, int64_t n, int stage, float* out);

/* Decoder activation taps for kernel bring-up: stage 0 = stem (post GN+ReLU) [64,4,4,4],
 * 1 = after res block, 2 = after attention, 3 = up_conv+pixel-shuffle [32,8,8,8]. */
int vqo_decode_tap(const vqo_model* m, const uint8_t* indices, int64_t n, int stage, float* out);

#ifdef __cplusplus
}
#endif
#endif
