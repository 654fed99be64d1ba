#include "B200Backend.hpp"

#include <iostream>
#include <stdexcept>
#include <string>

#include "../../include/vqvdb_b200.h"

namespace {
int64_t leadingDim(const TensorView& v, const char* what) {
	if (v.shape.empty() || v.shape[0] < 0) throw std::runtime_error(std::string(what) + ": tensor has no batch dimension");
	if (v.shape[0] > 0 && v.data == nullptr) throw std::runtime_error(std::string(what) + ": null data pointer");
	return v.shape[0];
}
}  // namespace

B200Backend::B200Backend(const CodecConfig& config) {
	if (config.device != CodecConfig::Device::CUDA)
		throw std::runtime_error("B200 backend needs CodecConfig::Device::CUDA (no CPU path exists)");
	vqvdb_b200_config c{};
	c.struct_size = sizeof(c);
	c.device = config.cudaDevice;
	c.chunk_leaves = config.chunkLeaves;
	c.decode_precision = config.fp32Decode ? VQVDB_B200_DECODE_FP32 : VQVDB_B200_DECODE_DEFAULT;
	c.encode_precision = config.fp32Encode ? VQVDB_B200_ENCODE_FP32 : VQVDB_B200_ENCODE_DEFAULT;
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
	Tensor out;
	out.dtype = DataType::UINT8;
	out.shape = {n, latentShape_[0], latentShape_[1], latentShape_[2]};
	out.buffer.resize((size_t)n * 64);
	encodeInto(static_cast<const float*>(leafBatch.data), n, out.getData<uint8_t>());
	return out;
}

Tensor B200Backend::decode(const TensorView& indices) const {
	if (indices.dtype != DataType::UINT8) throw std::runtime_error("decode expects UINT8 data.");
	const int64_t n = leadingDim(indices, "decode");
	Tensor out;
	out.dtype = DataType::FLOAT32;
	out.shape = {n, channels_, 8, 8, 8};
	out.buffer.resize((size_t)n * channels_ * 512 * sizeof(float));
	decodeInto(static_cast<const uint8_t*>(indices.data), n, out.getData<float>());
	return out;
}

// Factory.  Same contract as the reference (src/core/IVQVAECodec.cpp:76-110): swallow, log, return null.
std::unique_ptr<IVQVAECodec> IVQVAECodec::create(const CodecConfig& config, BackendType type) {
	try {
		if (type == BackendType::B200) return std::make_unique<B200Backend>(config);
		throw std::runtime_error("Requested backend type is not available or disabled in the build configuration.");
	} catch (const std::exception& e) {
		std::cerr << "Failed to create VQ-VAE backend: " << e.what() << std::endl;
		return nullptr;
	}
}
