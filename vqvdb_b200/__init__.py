"""vqvdb_b200 — B200-native VQ-VAE encode/decode engine for OpenVDB 8^3 leaf blocks.

Product code: csrc/ (sm_100a CUDA kernels + the C-ABI of include/vqvdb_b200.h), cpp/ (C++ backend
for the reference's IVQVAECodec interface and the OpenVDB-free batch loop), codec.py (ctypes mirror).
Nothing in this package imports oracle/.
"""
from .codec import (B200Codec, BackendType, CodecConfig, DataType, EmbeddedModel, IVQVAECodec, OnnxModelPaths, Tensor,
                    TensorView, convert_onnx, load_library, EXPORTS, LIB_PATH)

__all__ = ["B200Codec", "BackendType", "CodecConfig", "DataType", "EmbeddedModel", "IVQVAECodec", "OnnxModelPaths", "Tensor",
           "TensorView", "convert_onnx", "load_library", "EXPORTS", "LIB_PATH"]
