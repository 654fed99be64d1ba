// Backend interface of the VQ-VAE leaf codec, as the reference's orchestrator and SOPs see it.
//
// This header keeps the reference's boundary (names, argument meaning, error behaviour) so that a file
// written against /root/reference/src/core/IVQVAECodec.hpp:21-137 compiles against this one unchanged:
//   BackendType, EmbeddedModel, OnnxModelPaths, ModelSource, DataType, TensorView, Tensor, CodecConfig,
//   IVQVAECodec::{create, encode, decode, getLatentShape}.
// The ONLY addition is the BackendType::B200 enumerator: CodecConfig is the reference's, field for field (B200-only
// knobs — device ordinal, chunk size, checking paths — travel through B200Options in B200Backend.hpp, not through it),
// so B200Backend.cpp compiles against the reference's own header with that one token added (tests/test_boundary.py).
// The reference's own LibTorch / ONNX backends are not part of this repository; asking create() for them
// yields nullptr exactly as a reference build without ENABLE_*_BACKEND does (IVQVAECodec.cpp:99-103).
#pragma once

#include <cstddef>
#include <cstdint>
#include <filesystem>
#include <memory>
#include <variant>
#include <vector>

enum class BackendType { LibTorch, ONNX, B200 };

struct EmbeddedModel {};
struct OnnxModelPaths {
	std::filesystem::path encoder_path;
	std::filesystem::path decoder_path;
};
// For BackendType::B200 a std::filesystem::path names a VQVDBW01 weight pack (tools/weights_pack.py) or a directory holding
// encoder.onnx + decoder.onnx; OnnxModelPaths names the two graphs; only their initializers (weights) are read.
using ModelSource = std::variant<EmbeddedModel, std::filesystem::path, OnnxModelPaths>;

enum class DataType { FLOAT32, UINT8 };

// Borrowed, read-only tensor: `data` is caller-owned HOST memory valid for the duration of the call.
struct TensorView {
	const void* data = nullptr;
	std::vector<int64_t> shape;
	DataType dtype;
};

// Owning tensor returned by the codec.
struct Tensor {
	std::vector<std::byte> buffer;
	std::vector<int64_t> shape;
	DataType dtype;

	template <typename T>
	const T* getData() const { return reinterpret_cast<const T*>(buffer.data()); }
	template <typename T>
	T* getData() { return reinterpret_cast<T*>(buffer.data()); }
};

struct CodecConfig {
	enum class Device { CPU, CUDA };
	Device device = Device::CPU;
	ModelSource source = EmbeddedModel{};
};

class IVQVAECodec {
   public:
	virtual ~IVQVAECodec() = default;

	// Never throws: any failure is logged to stderr and reported as nullptr (IVQVAECodec.cpp:106-109).
	static std::unique_ptr<IVQVAECodec> create(const CodecConfig& config, BackendType type);

	// FLOAT32 [B, C, 8, 8, 8] host leaves -> UINT8 [B, 4, 4, 4] codebook indices.  Throws std::runtime_error.
	virtual Tensor encode(const TensorView& leafBatch) const = 0;
	// UINT8 [B, 4, 4, 4] host indices -> FLOAT32 [B, C, 8, 8, 8] reconstructed leaves.  Throws std::runtime_error.
	virtual Tensor decode(const TensorView& indices) const = 0;
	virtual const std::vector<int64_t>& getLatentShape() const = 0;
};
