// Host-side reader for the VQVDBW01 flat weight pack (tools/weights_pack.py) and the
// device-side weight tables the kernels consume.
//
// The reference keeps its model as an embedded TorchScript/ONNX blob
// (src/Bin/bin_model.h:14, src/Bin/bin_onnx.h) and lets LibTorch/ORT interpret it
// (TorchBackend.cpp:38-60, OnnxBackendFactory.cpp:97-145).  Here the same 45 fp32 tensors
// are read once, re-laid-out for the kernels (convolution weights transposed to
// [cin][kd][kh][kw][cout] so a thread's output-channel slice is contiguous) and uploaded.
#pragma once

#include <cstddef>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace vqvdb {

struct PackTensor {
	std::vector<int> dims;
	const float* data = nullptr;  // into WeightPack::blob
	size_t numel() const {
		size_t n = 1;
		for (int d : dims) n *= (size_t)d;
		return n;
	}
};

struct WeightPack {
	std::vector<unsigned char> blob;
	std::map<std::string, PackTensor> tensors;
	int in_channels = 0, embedding_dim = 0, num_embeddings = 0;

	// Throws std::runtime_error on any malformed input.
	void parse(const void* data, size_t size);
	void load_file(const std::string& path);
	const PackTensor& get(const std::string& name) const;
};

// The float model's pack, linked into the library (weights_embed.S); the counterpart of the
// reference's EmbeddedModel source.
extern "C" const unsigned char vqvdb_b200_embedded_pack[];
extern "C" const unsigned char vqvdb_b200_embedded_pack_end[];

// ONNX-initializer reader (weights_onnx.cpp): the two graphs python/to_onnx.py writes -> VQVDBW01 pack bytes.
std::vector<unsigned char> onnx_to_pack(const std::vector<unsigned char>& encoder_onnx, const std::vector<unsigned char>& decoder_onnx);
std::vector<unsigned char> read_file_bytes(const std::string& path);

// Conv weight [cout][cin][k][k][k] -> [cin][k][k][k][cout]
std::vector<float> transpose_conv_weight(const PackTensor& w);

}  // namespace vqvdb
