// TEST INFRASTRUCTURE — not part of the product path.
//
// C-ABI shim around the UNMODIFIED reference backend so that tests/ and
// bench.py (--impl reference, cpu_baseline) can drive it through ctypes.
// The reference sources are compiled where they lie, never copied:
//   /root/reference/src/core/IVQVAECodec.cpp           (factory, :76-110)
//   /root/reference/src/backends/torch/TorchBackend.cpp (encode :133-164, decode :166-194)
// built by oracle/Makefile into oracle/_ref/libvqvdb_ref.so (git-ignored).
//
// Every call below goes through the reference's own public interface
// (IVQVAECodec::create / encode / decode / getLatentShape) with host pointers,
// exactly as VQVAECodec::encodeBatch/decodeBatch do (orchestrator/VQVAECodec.cpp:210-212).
#include <torch/torch.h>

#include <cstdint>
#include <cstring>
#include <memory>
#include <string>

#include "core/IVQVAECodec.hpp"

namespace {
thread_local std::string g_err;
struct RefHandle {
	std::unique_ptr<IVQVAECodec> codec;
};
}  // namespace

extern "C" {

// device: 0 = CodecConfig::Device::CPU, 1 = CodecConfig::Device::CUDA.
// threads > 0 overrides the reference's own hw/2 choice (TorchBackend.cpp:74-78)
// AFTER construction; threads <= 0 keeps the reference default.
void* vqvdb_ref_create(int device, int threads) {
	CodecConfig cfg;
	cfg.device = device ? CodecConfig::Device::CUDA : CodecConfig::Device::CPU;
	cfg.source = EmbeddedModel{};
	auto codec = IVQVAECodec::create(cfg, BackendType::LibTorch);
	if (!codec) {
		g_err = "IVQVAECodec::create returned nullptr";
		return nullptr;
	}
	if (threads > 0) torch::set_num_threads(threads);
	auto* h = new RefHandle{std::move(codec)};
	return h;
}

void vqvdb_ref_destroy(void* h) { delete static_cast<RefHandle*>(h); }

int vqvdb_ref_threads() { return torch::get_num_threads(); }

int vqvdb_ref_cuda_available() { return torch::cuda::is_available() ? 1 : 0; }

int vqvdb_ref_latent_shape(void* h, int64_t* out, int cap) {
	const auto& s = static_cast<RefHandle*>(h)->codec->getLatentShape();
	for (int i = 0; i < (int)s.size() && i < cap; ++i) out[i] = s[i];
	return (int)s.size();
}

int vqvdb_ref_encode(void* h, const float* leaves, int64_t n, uint8_t* indices) {
	try {
		TensorView v;
		v.data = leaves;
		v.shape = {n, 1, 8, 8, 8};
		v.dtype = DataType::FLOAT32;
		Tensor t = static_cast<RefHandle*>(h)->codec->encode(v);
		std::memcpy(indices, t.buffer.data(), t.buffer.size());
		return 0;
	} catch (const std::exception& e) {
		g_err = e.what();
		return -1;
	}
}

int vqvdb_ref_decode(void* h, const uint8_t* indices, int64_t n, float* voxels) {
	try {
		TensorView v;
		v.data = indices;
		v.shape = {n, 4, 4, 4};
		v.dtype = DataType::UINT8;
		Tensor t = static_cast<RefHandle*>(h)->codec->decode(v);
		std::memcpy(voxels, t.buffer.data(), t.buffer.size());
		return 0;
	} catch (const std::exception& e) {
		g_err = e.what();
		return -1;
	}
}

const char* vqvdb_ref_last_error() { return g_err.c_str(); }

}  // extern "C"
