// IVQVAECodec implementation that forwards to the C-ABI of libvqvdb_b200.so (include/vqvdb_b200.h).
// It replaces the reference's TorchBackend / OnnxCudaBackend (src/backends/torch/TorchBackend.cpp:84-194,
// src/backends/onnx/OnnxBackend_Cuda.cpp:19-165) behind the same interface.
//
// This file and B200Backend.cpp compile against the reference's own src/core/IVQVAECodec.hpp once the
// `B200` enumerator is added to BackendType (tests/test_boundary.py does exactly that
// compile): they use nothing of CodecConfig beyond `device` and `source` (IVQVAECodec.hpp:85-89).
#pragma once

#include <cstdint>

#include "IVQVAECodec.hpp"

struct vqvdb_b200_codec;

// Everything the B200 backend can be told that the reference's CodecConfig has no field for.  A side channel on
// purpose: the reference's struct stays untouched.  Defaults come from the environment (read once per backend):
//   VQVDB_B200_DEVICE=<ordinal>  VQVDB_B200_CHUNK_LEAVES=<n>  VQVDB_B200_DECODE=fp32  VQVDB_B200_ENCODE=fp32
// and can be overridden process-wide with setDefaultOptions() or per backend with the two-argument constructor.
struct B200Options {
	int cudaDevice = 0;        // the reference hard-codes device 0 (OnnxBackend_Cuda.cpp:21)
	uint32_t chunkLeaves = 0;  // leaves per internal pipeline chunk of the host-pointer calls, 0 = default
	bool fp32Decode = false;   // CUDA-core fp32 decoder (checking path) instead of the bf16 tensor-core one
	bool fp32Encode = false;   // CUDA-core fp32 FFMA encoder (checking path) instead of the split-fp16 tensor-core one
	static B200Options fromEnvironment();
};

class B200Backend final : public IVQVAECodec {
   public:
	explicit B200Backend(const CodecConfig& config);  // options = defaultOptions(); throws std::runtime_error; no CPU fallback
	B200Backend(const CodecConfig& config, const B200Options& options);
	~B200Backend() override;
	B200Backend(const B200Backend&) = delete;
	B200Backend& operator=(const B200Backend&) = delete;

	// Process-wide defaults used by the one-argument constructor, i.e. by IVQVAECodec::create(config, BackendType::B200).
	static void setDefaultOptions(const B200Options& options);
	static B200Options defaultOptions();  // what was set, else fromEnvironment()

	Tensor encode(const TensorView& leafBatch) const override;
	Tensor decode(const TensorView& indices) const override;
	const std::vector<int64_t>& getLatentShape() const override { return latentShape_; }

	// Zero-copy forms used by the batch loop: results are written straight into caller memory, which
	// avoids the allocate + zero-fill + memcpy of the Tensor return path (IVQVAECodec.hpp:61-80 in the reference).
	void encodeInto(const float* hostLeaves, int64_t nLeaves, uint8_t* hostIndices) const;
	void decodeInto(const uint8_t* hostIndices, int64_t nLeaves, float* hostVoxels) const;
	int channels() const { return channels_; }

   private:
	void init(const CodecConfig& config, const B200Options& options);
	vqvdb_b200_codec* handle_ = nullptr;
	std::vector<int64_t> latentShape_;
	int channels_ = 1;
};
