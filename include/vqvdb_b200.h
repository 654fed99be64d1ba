/* vqvdb_b200.h — C ABI of the B200-native VQ-VAE leaf codec.
 *
 * This is the drop-in boundary for the reference's backend interface
 *     class IVQVAECodec        /root/reference/src/core/IVQVAECodec.hpp:99-137
 * whose only callers are VQVAECodec::encodeBatch/decodeBatch
 *     /root/reference/src/orchestrator/VQVAECodec.cpp:210,212
 * A reference-side backend (`B200Backend final : IVQVAECodec`, see
 * vqvdb_b200/cpp/B200Backend.{hpp,cpp} and INTEGRATION.md) forwards each virtual
 * to exactly one entry point below.  Plain pointers and sizes only: no C++ types,
 * no exceptions across the boundary, no torch/ORT types.
 *
 * Conventions
 *   - every function returns 0 on success or a negative vqvdb_b200_status;
 *     vqvdb_b200_last_error() gives the message (per codec; pass NULL for the
 *     calling thread's last create() failure).
 *   - a codec is bound to ONE CUDA device and owns its streams and staging buffers;
 *     calls on one codec must be serialised by the caller (the reference's batch
 *     loop is single-threaded: VQVAECodec.cpp:108-127); different codecs are independent.
 *   - there is NO CPU fallback: if the device or the sm_100a kernels are unavailable,
 *     create() fails (the reference's TorchBackend.cpp:64-72 silently falls back; this does not).
 *   - layouts are the reference's: leaves float32 [n, C, 8, 8, 8] with OpenVDB leaf-buffer
 *     order (IVQVAECodec.hpp:116-121), indices uint8 [n, 4, 4, 4] (d-major), n may be 0.
 */
#ifndef VQVDB_B200_H
#define VQVDB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define VQVDB_B200_API __declspec(dllexport)
#else
#define VQVDB_B200_API __attribute__((visibility("default")))
#endif

typedef struct vqvdb_b200_codec vqvdb_b200_codec;

typedef enum vqvdb_b200_status {
	VQVDB_B200_OK = 0,
	VQVDB_B200_ERR_INVALID_ARGUMENT = -1,
	VQVDB_B200_ERR_NO_DEVICE = -2,      /* no CUDA device / not an sm_100 part / kernels missing */
	VQVDB_B200_ERR_BAD_WEIGHTS = -3,    /* weight pack missing, truncated or of an unknown architecture */
	VQVDB_B200_ERR_CUDA = -4,           /* a CUDA runtime call failed; message has the cudaError string */
	VQVDB_B200_ERR_OUT_OF_MEMORY = -5,
	VQVDB_B200_ERR_UNSUPPORTED = -6
} vqvdb_b200_status;

/* Decoder arithmetic.  The encoder and the codebook argmin always run fp32-faithful
 * (index parity with the reference is a bit-exactness requirement); the decoder may
 * use BF16 tensor-core operands with fp32 accumulation. */
typedef enum vqvdb_b200_decode_precision {
	VQVDB_B200_DECODE_DEFAULT = 0, /* the tensor-core path: meets the 0.1 dB PSNR budget */
	VQVDB_B200_DECODE_FP32 = 1,    /* CUDA-core fp32 path (checking path: agrees with the reference reconstruction to 2e-5) */
	VQVDB_B200_DECODE_BF16_TC = 2  /* tcgen05.mma + TMEM accumulators, bf16 operands, fp32 accumulation; kw taps concatenated
	                                * along N (N = 192); up_conv -> PixelShuffle3D -> final folded on the host into one 64 -> 64
	                                * convolution + an 8-term gather per voxel (exact algebra, zero padding at both resolutions
	                                * included; weights folded in double, rounded to bf16 once) */
} vqvdb_b200_decode_precision;
#define VQVDB_B200_DECODE_DEFAULT_KIND VQVDB_B200_DECODE_BF16_TC

/* Encoder arithmetic.  Both paths are fp32-faithful (index parity with the reference is a bit-exactness requirement):
 * the tensor-core path splits every operand into two fp16 planes (22 significant bits, three products, fp32
 * accumulation); the float model re-scores the codebook shortlist in exact fp32, the vec3 model (two kernels,
 * csrc/encode_tc128_front.cu + csrc/encode_tc128.cu) computes proj and all distances as exact fp32 FMA chains. */
typedef enum vqvdb_b200_encode_precision {
	VQVDB_B200_ENCODE_DEFAULT = 0,
	VQVDB_B200_ENCODE_FP32 = 1,      /* CUDA-core fp32 FFMA path */
	VQVDB_B200_ENCODE_FP16X2_TC = 2  /* tcgen05.mma + TMEM accumulators on split-fp16 operands */
} vqvdb_b200_encode_precision;
#define VQVDB_B200_ENCODE_DEFAULT_KIND VQVDB_B200_ENCODE_FP16X2_TC

/* Replaces CodecConfig{device, source} (IVQVAECodec.hpp:85-89).  Zero-initialise, set
 * struct_size = sizeof(vqvdb_b200_config), then fill what you need. */
typedef struct vqvdb_b200_config {
	uint32_t struct_size;
	int32_t device;               /* CUDA ordinal (the reference hard-codes 0: OnnxBackend_Cuda.cpp:21) */
	const void* weights_data;     /* VQVDBW01 pack in memory, or NULL */
	uint64_t weights_size;
	const char* weights_path;     /* VQVDBW01 pack on disk, or a directory holding encoder.onnx + decoder.onnx (the reference's
	                                 model-directory source, OnnxBackendFactory.cpp:97-119), or NULL.  All sources NULL = the
	                                 embedded float model (the reference's EmbeddedModel source, IVQVAECodec.hpp:27) */
	uint32_t chunk_leaves;        /* leaves per internal pipeline chunk for the host-pointer calls; 0 = default */
	uint32_t decode_precision;    /* vqvdb_b200_decode_precision */
	uint32_t encode_precision;    /* vqvdb_b200_encode_precision */
	uint32_t reserved0;
	const char* onnx_encoder_path; /* the reference's OnnxModelPaths source (IVQVAECodec.hpp:29-33): the two graphs written by */
	const char* onnx_decoder_path; /* python/to_onnx.py; only their initializers (weights) are read — no ONNX Runtime involved */
	uint32_t reserved[2];
} vqvdb_b200_config;

/* IVQVAECodec::create (IVQVAECodec.cpp:76-110).  On failure *out is NULL. */
VQVDB_B200_API int vqvdb_b200_create(const vqvdb_b200_config* cfg, vqvdb_b200_codec** out);
VQVDB_B200_API void vqvdb_b200_destroy(vqvdb_b200_codec* codec);

/* IVQVAECodec::getLatentShape (IVQVAECodec.hpp:136): writes {4,4,4}. */
VQVDB_B200_API int vqvdb_b200_latent_shape(const vqvdb_b200_codec* codec, int64_t out_dhw[3]);
/* Channel count of the loaded model: 1 (FloatGrid) or 3 (Vec3fGrid). */
VQVDB_B200_API int vqvdb_b200_in_channels(const vqvdb_b200_codec* codec);
VQVDB_B200_API int vqvdb_b200_num_embeddings(const vqvdb_b200_codec* codec);

/* IVQVAECodec::encode (IVQVAECodec.hpp:121; TorchBackend.cpp:133-164).
 * host_leaves: n*C*512 floats, caller-owned HOST memory (pageable or pinned);
 * host_indices: n*64 bytes, caller-owned HOST memory.  Synchronous. */
VQVDB_B200_API int vqvdb_b200_encode(vqvdb_b200_codec* codec, const float* host_leaves, int64_t n_leaves,
                                     uint8_t* host_indices);

/* IVQVAECodec::decode (IVQVAECodec.hpp:130; TorchBackend.cpp:166-194).  Synchronous. */
VQVDB_B200_API int vqvdb_b200_decode(vqvdb_b200_codec* codec, const uint8_t* host_indices, int64_t n_leaves,
                                     float* host_voxels);

/* Device-pointer variants for the pipelined batch loop and the multi-GPU path: buffers live
 * on the codec's device, work is enqueued on `cuda_stream` (a cudaStream_t, used exactly as given:
 * NULL is the CUDA legacy default stream) and the call returns without synchronising.
 * Alignment: dev_leaves and dev_voxels 16 bytes (the kernels move leaves as 128-bit vectors), dev_indices 4 bytes;
 * anything cudaMalloc returns, offset by whole leaves, qualifies.  Calls on ONE codec must be serialised by the
 * caller even when they target different streams (the vec3 model's kernels share one per-codec scratch area). */
VQVDB_B200_API int vqvdb_b200_encode_device(vqvdb_b200_codec* codec, const float* dev_leaves, int64_t n_leaves,
                                            uint8_t* dev_indices, void* cuda_stream);
VQVDB_B200_API int vqvdb_b200_decode_device(vqvdb_b200_codec* codec, const uint8_t* dev_indices, int64_t n_leaves,
                                            float* dev_voxels, void* cuda_stream);
/* Blocks until the codec's own pipeline streams are idle (work enqueued on caller streams is the caller's). */
VQVDB_B200_API int vqvdb_b200_synchronize(vqvdb_b200_codec* codec);

/* Multi-GPU reassembly without a staging copy (leaves shard across the GPUs of one box, decoded blocks
 * travel to the rank that rebuilds the grid — the step the reference's single-GPU loop, VQVAECodec.cpp:166-200, does
 * not have).  The reassembly rank creates the gathered buffer on its device and exports a CUDA IPC handle; every other
 * rank (one process per GPU) opens it and passes `base + leaf_offset * 2048` as dev_voxels to vqvdb_b200_decode_device,
 * so the decode kernel's own epilogue stores land in the owner's HBM over NVLink: compute and gather are one kernel.
 *   create: cudaMalloc on the codec's device; handle_out receives the 64-byte cudaIpcMemHandle_t.
 *   open  : maps a peer's buffer into this process (lazy peer access); not valid in the creating process.
 *   close : unmaps (opened != 0) or frees (opened == 0). */
VQVDB_B200_API int vqvdb_b200_peer_buffer_create(vqvdb_b200_codec* codec, uint64_t bytes, void** dev_ptr_out,
                                                 unsigned char handle_out[64]);
VQVDB_B200_API int vqvdb_b200_peer_buffer_open(vqvdb_b200_codec* codec, const unsigned char handle[64], void** dev_ptr_out);
VQVDB_B200_API int vqvdb_b200_peer_buffer_close(vqvdb_b200_codec* codec, void* dev_ptr, int opened);

/* Kernel launches issued by this codec since creation (bench.py's gpu_launches). */
VQVDB_B200_API uint64_t vqvdb_b200_kernel_launches(const vqvdb_b200_codec* codec);
/* Name of the decode path actually in use: "bf16_tcgen05_n192_fold" (float model), "bf16_tcgen05_c128_fold" (the vec3
 * model's 128-channel decoder), "fp32", or "fp32_generic" (vec3 model with VQVDB_B200_DECODE_FP32). */
VQVDB_B200_API const char* vqvdb_b200_decode_path(const vqvdb_b200_codec* codec);

/* Bring-up aid for the tensor-core decoders: runs one and also writes the fp32 activation after stage
 * {0: stem+GroupNorm+ReLU, 1: residual block(s), 2: channel attention} as [n][64 ch][64 pos] (float model) or
 * [n][128 ch][64 pos] (vec3 model) to dev_tap. */
VQVDB_B200_API int vqvdb_b200_debug_decode_tap(vqvdb_b200_codec* codec, const uint8_t* dev_indices, int64_t n_leaves,
                                               int stage, float* dev_tap, float* dev_voxels, void* cuda_stream);

/* Host-only checker hook (no device needed): the folded decoder tail of the weight pack at `weights_path` (NULL or ""
 * = the embedded pack) — conv weights [64 (r*8+eps)][64 cin][27 taps] and bias [64] in fp32, exactly what the
 * tensor-core decode path rounds to bf16 (see VQVDB_B200_DECODE_BF16_TC; tests/test_decoder_fold.py compares it with
 * up_conv -> PixelShuffle3D -> final of python/VQVAE_v2.py:266-275 evaluated layer by layer). */
VQVDB_B200_API int vqvdb_b200_debug_fold_decoder_tail(const char* weights_path, float* weights_out, float* bias_out);

/* Host-only checker hook (no device needed): the tensor-core encoder's `proj x codebook` fold of the weight pack at
 * `weights_path` (NULL or "" = the embedded pack) — m_out [256][32] = E . proj.weight, esq_out [256] = |e_k|^2 - 2 proj.bias.e_k,
 * norm_out [257] = |M_k| rounded up and their maximum (csrc/encode_tc_stream.hpp; tests/test_encoder_fold.py checks
 * |W x + b - e_k|^2 - |W x + b|^2 == esq_k - 2 x.M_k against python/save_for_inference.py:55-61 evaluated directly).
 * For a vec3 pack (csrc/encode_tc128_stream.hpp): m_out [256][128], esq_out [256], norm_out [257] with [0] = max_k |M_k|,
 * [1] = an upper bound of |proj.weight|_2, [2] = |proj.bias| + max_k |e_k|, the rest 0. */
VQVDB_B200_API int vqvdb_b200_debug_fold_encoder_vq(const char* weights_path, float* m_out, float* esq_out, float* norm_out);

/* Name of the encode path in use: "fp16x2_tcgen05" or "fp32" (float model); "fp16x2_tcgen05_c128" or "fp32_generic"
 * (vec3 model). */
VQVDB_B200_API const char* vqvdb_b200_encode_path(const vqvdb_b200_codec* codec);
/* Bring-up aid for the tensor-core encoder: runs it and also writes an fp32 activation to dev_tap — stage 0: pre
 * (GroupNorm+ReLU) [n][16][512], 6: first residual block's conv1 [n][16][512], 1: first residual block [n][16][512],
 * 2: down [n][32][64], 7: res_stack.0 conv1 [n][32][64], 3: residual stack [n][32][64], 4: channel attention
 * [n][32][64], 5: z [n][128][64].
 * vec3 model (at most one batch = SM count x 28 leaves): stage 4: pre (GroupNorm+ReLU) [n][64][512], 5: the 8^3 residual
 * block [n][64][512], 6: down1 [n][128][64], 0: res_stack.0 [n][128][64], 1: res_stack.1, 2: channel attention, 3: z;
 * 100: phase timestamps of both kernels (tools/check_vec3_encode.py). */
VQVDB_B200_API int vqvdb_b200_debug_encode_tap(vqvdb_b200_codec* codec, const float* dev_leaves, int64_t n_leaves, int stage,
                                               float* dev_tap, uint8_t* dev_indices, void* cuda_stream);

/* Host-only tool (no device needed): reads the weights out of encoder.onnx + decoder.onnx and writes them as a VQVDBW01
 * pack.  On failure the message is vqvdb_b200_last_error(NULL). */
VQVDB_B200_API int vqvdb_b200_convert_onnx(const char* encoder_onnx_path, const char* decoder_onnx_path, const char* out_pack_path);

VQVDB_B200_API const char* vqvdb_b200_last_error(const vqvdb_b200_codec* codec);
VQVDB_B200_API const char* vqvdb_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* VQVDB_B200_H */
