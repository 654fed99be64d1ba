#include "B200Backend.hpp"

#include <cstdlib>
#include <iostream>
#include <mutex>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "vqvdb_b200.h"

#if defined(__linux__)
#include <sys/mman.h>
#endif

namespace {
// Tensor::buffer is a std::vector<std::byte> (IVQVAECodec.hpp:61-80), and vector::resize value-initialises: for a decoded
// 1 M-leaf grid that is 2 GB of zeros written by one thread, every page faulted in on the way, before the first
// decoded voxel can land — 0.7 s against 0.06 s of decoding.  With libstdc++ the buffer is instead reserve()d (address
// space only), filled by the backend's staging threads (first touch, hence the page faults, spread over them and
// overlapped with the GPU), and then given its size without a second pass.  Elsewhere: plain resize().
#if defined(__GLIBCXX__)
struct ByteVectorAccess : std::vector<std::byte> {
	static void setSizeWithinCapacity(std::vector<std::byte>& v, size_t n) {
		auto& impl = static_cast<ByteVectorAccess&>(v)._M_impl;
		impl._M_finish = impl._M_start + n;
	}
};
std::byte* uninitializedBytes(std::vector<std::byte>& v, size_t n) {
	v.reserve(n);
#if defined(__linux__)
	if (n >= (size_t(64) << 20)) {  // a fresh mapping: ask for huge pages before the first touch (512x fewer faults where THP allows)
		const uintptr_t lo = (reinterpret_cast<uintptr_t>(v.data()) + 0x1fffff) & ~uintptr_t(0x1fffff);
		const uintptr_t hi = (reinterpret_cast<uintptr_t>(v.data()) + n) & ~uintptr_t(0x1fffff);
		if (hi > lo) madvise(reinterpret_cast<void*>(lo), hi - lo, MADV_HUGEPAGE);
	}
#endif
	return v.data();
}
void commitBytes(std::vector<std::byte>& v, size_t n) { ByteVectorAccess::setSizeWithinCapacity(v, n); }
#else
std::byte* uninitializedBytes(std::vector<std::byte>& v, size_t n) {
	v.resize(n);
	return v.data();
}
void commitBytes(std::vector<std::byte>&, size_t) {}
#endif

int64_t leadingDim(const TensorView& v, const char* what) {
	if (v.shape.empty() || v.shape[0] < 0) throw std::runtime_error(std::string(what) + ": tensor has no batch dimension");
	if (v.shape[0] > 0 && v.data == nullptr) throw std::runtime_error(std::string(what) + ": null data pointer");
	return v.shape[0];
}

// The reference builds a torch tensor from the view's shape and the model rejects anything but [B,C,8,8,8] /
// [B,4,4,4] (TorchBackend.cpp:141-149,173-180); here the raw pointer is all the C ABI sees, so the shape is
// checked before it is trusted.
void expectShape(const TensorView& v, const std::vector<int64_t>& tail, const char* what) {
	bool ok = v.shape.size() == tail.size() + 1;
	for (size_t i = 0; ok && i < tail.size(); ++i) ok = v.shape[i + 1] == tail[i];
	if (ok) return;
	std::string want = "[B";
	for (int64_t d : tail) want += "," + std::to_string(d);
	std::string got = "[";
	for (size_t i = 0; i < v.shape.size(); ++i) got += (i ? "," : "") + std::to_string(v.shape[i]);
	throw std::runtime_error(std::string(what) + ": expected shape " + want + "], got " + got + "]");
}

std::mutex g_optMutex;
std::optional<B200Options> g_defaultOptions;

bool envIs(const char* name, const char* value) {
	const char* v = std::getenv(name);
	return v && std::string(v) == value;
}
}  // namespace

B200Options B200Options::fromEnvironment() {
	B200Options o;
	if (const char* v = std::getenv("VQVDB_B200_DEVICE")) o.cudaDevice = std::atoi(v);
	if (const char* v = std::getenv("VQVDB_B200_CHUNK_LEAVES")) o.chunkLeaves = (uint32_t)std::strtoul(v, nullptr, 10);
	o.fp32Decode = envIs("VQVDB_B200_DECODE", "fp32");
	o.fp32Encode = envIs("VQVDB_B200_ENCODE", "fp32");
	return o;
}

void B200Backend::setDefaultOptions(const B200Options& options) {
	std::lock_guard<std::mutex> lock(g_optMutex);
	g_defaultOptions = options;
}

B200Options B200Backend::defaultOptions() {
	std::lock_guard<std::mutex> lock(g_optMutex);
	return g_defaultOptions ? *g_defaultOptions : B200Options::fromEnvironment();
}

B200Backend::B200Backend(const CodecConfig& config) { init(config, defaultOptions()); }
B200Backend::B200Backend(const CodecConfig& config, const B200Options& options) { init(config, options); }

void B200Backend::init(const CodecConfig& config, const B200Options& options) {
	if (config.device != CodecConfig::Device::CUDA)
		throw std::runtime_error("B200 backend needs CodecConfig::Device::CUDA (no CPU path exists)");
	vqvdb_b200_config c{};
	c.struct_size = sizeof(c);
	c.device = options.cudaDevice;
	c.chunk_leaves = options.chunkLeaves;
	c.decode_precision = options.fp32Decode ? VQVDB_B200_DECODE_FP32 : VQVDB_B200_DECODE_DEFAULT;
	c.encode_precision = options.fp32Encode ? VQVDB_B200_ENCODE_FP32 : VQVDB_B200_ENCODE_DEFAULT;
	std::string path, path2;
	if (std::holds_alternative<std::filesystem::path>(config.source)) {
		path = std::get<std::filesystem::path>(config.source).string();   // a VQVDBW01 pack, or a directory with encoder.onnx + decoder.onnx
		c.weights_path = path.c_str();
	} else if (std::holds_alternative<OnnxModelPaths>(config.source)) {  // weights are read from the graphs' initializers
		path = std::get<OnnxModelPaths>(config.source).encoder_path.string();
		path2 = std::get<OnnxModelPaths>(config.source).decoder_path.string();
		c.onnx_encoder_path = path.c_str();
		c.onnx_decoder_path = path2.c_str();
	}
	if (vqvdb_b200_create(&c, &handle_) != VQVDB_B200_OK || !handle_)
		throw std::runtime_error(std::string("vqvdb_b200_create: ") + vqvdb_b200_last_error(nullptr));
	int64_t dhw[3];
	vqvdb_b200_latent_shape(handle_, dhw);
	latentShape_.assign(dhw, dhw + 3);
	channels_ = vqvdb_b200_in_channels(handle_);
	std::cout << "B200Backend: " << vqvdb_b200_version() << ", device " << c.device << ", encode path "
	          << vqvdb_b200_encode_path(handle_) << ", decode path " << vqvdb_b200_decode_path(handle_) << std::endl;
}

B200Backend::~B200Backend() { vqvdb_b200_destroy(handle_); }

void B200Backend::encodeInto(const float* hostLeaves, int64_t n, uint8_t* hostIndices) const {
	if (vqvdb_b200_encode(handle_, hostLeaves, n, hostIndices) != VQVDB_B200_OK)
		throw std::runtime_error(std::string("B200 encode failed: ") + vqvdb_b200_last_error(handle_));
}

void B200Backend::decodeInto(const uint8_t* hostIndices, int64_t n, float* hostVoxels) const {
	if (vqvdb_b200_decode(handle_, hostIndices, n, hostVoxels) != VQVDB_B200_OK)
		throw std::runtime_error(std::string("B200 decode failed: ") + vqvdb_b200_last_error(handle_));
}

Tensor B200Backend::encode(const TensorView& leafBatch) const {
	if (leafBatch.dtype != DataType::FLOAT32) throw std::runtime_error("encode expects FLOAT32 data.");
	const int64_t n = leadingDim(leafBatch, "encode");
	expectShape(leafBatch, {channels_, 8, 8, 8}, "encode");
	Tensor out;
	out.dtype = DataType::UINT8;
	out.shape = {n, latentShape_[0], latentShape_[1], latentShape_[2]};
	const size_t bytes = (size_t)n * 64;
	encodeInto(static_cast<const float*>(leafBatch.data), n, reinterpret_cast<uint8_t*>(uninitializedBytes(out.buffer, bytes)));
	commitBytes(out.buffer, bytes);
	return out;
}

Tensor B200Backend::decode(const TensorView& indices) const {
	if (indices.dtype != DataType::UINT8) throw std::runtime_error("decode expects UINT8 data.");
	const int64_t n = leadingDim(indices, "decode");
	expectShape(indices, latentShape_, "decode");
	Tensor out;
	out.dtype = DataType::FLOAT32;
	out.shape = {n, channels_, 8, 8, 8};
	const size_t bytes = (size_t)n * channels_ * 512 * sizeof(float);
	decodeInto(static_cast<const uint8_t*>(indices.data), n, reinterpret_cast<float*>(uninitializedBytes(out.buffer, bytes)));
	commitBytes(out.buffer, bytes);
	return out;
}
