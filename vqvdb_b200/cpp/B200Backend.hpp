// IVQVAECodec implementation that forwards to the C-ABI of libvqvdb_b200.so (include/vqvdb_b200.h).
// It replaces the reference's TorchBackend / OnnxCudaBackend (src/backends/torch/TorchBackend.cpp:84-194,
// src/backends/onnx/OnnxBackend_Cuda.cpp:19-165) behind the same interface.
#pragma once

#include "IVQVAECodec.hpp"

struct vqvdb_b200_codec;

class B200Backend final : public IVQVAECodec {
   public:
	explicit B200Backend(const CodecConfig& config);  // throws std::runtime_error; there is no CPU fallback
	~B200Backend() override;
	B200Backend(const B200Backend&) = delete;
	B200Backend& operator=(const B200Backend&) = delete;

	Tensor encode(const TensorView& leafBatch) const override;
	Tensor decode(const TensorView& indices) const override;
	const std::vector<int64_t>& getLatentShape() const override { return latentShape_; }

	// Zero-copy forms used by the batch loop: results are written straight into caller memory, which
	// avoids the allocate + zero-fill + memcpy of the Tensor return path (IVQVAECodec.hpp:61-80 in the reference).
	void encodeInto(const float* hostLeaves, int64_t nLeaves, uint8_t* hostIndices) const;
	void decodeInto(const uint8_t* hostIndices, int64_t nLeaves, float* hostVoxels) const;
	int channels() const { return channels_; }

   private:
	vqvdb_b200_codec* handle_ = nullptr;
	std::vector<int64_t> latentShape_;
	int channels_ = 1;
};
