"""ctypes binding of the C-ABI (include/vqvdb_b200.h) plus a Python mirror of the reference's
backend interface, so tests and bench.py read like calls on the reference's IVQVAECodec
(/root/reference/src/core/IVQVAECodec.hpp:99-137).

There is no fallback of any kind here: if libvqvdb_b200.so is missing, cannot be loaded, or finds
no sm_100 device, construction raises.
"""
from __future__ import annotations

import ctypes as C
import enum
import os
from dataclasses import dataclass, field
from typing import Optional, Sequence, Union

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VQVDB_B200_LIB") or os.path.join(HERE, "libvqvdb_b200.so")   # the override is for tuning builds (vqvdb_b200/build.py)

EXPORTS = [  # every symbol include/vqvdb_b200.h declares
    "vqvdb_b200_create", "vqvdb_b200_destroy", "vqvdb_b200_latent_shape", "vqvdb_b200_in_channels",
    "vqvdb_b200_num_embeddings", "vqvdb_b200_encode", "vqvdb_b200_decode", "vqvdb_b200_encode_device",
    "vqvdb_b200_decode_device", "vqvdb_b200_synchronize", "vqvdb_b200_kernel_launches",
    "vqvdb_b200_decode_path", "vqvdb_b200_last_error", "vqvdb_b200_version", "vqvdb_b200_debug_decode_tap",
    "vqvdb_b200_encode_path", "vqvdb_b200_debug_encode_tap",
    "vqvdb_b200_peer_buffer_create", "vqvdb_b200_peer_buffer_open", "vqvdb_b200_peer_buffer_close",
    "vqvdb_b200_convert_onnx", "vqvdb_b200_debug_fold_decoder_tail", "vqvdb_b200_debug_fold_encoder_vq",
]


class _Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("device", C.c_int32),
        ("weights_data", C.c_void_p),
        ("weights_size", C.c_uint64),
        ("weights_path", C.c_char_p),
        ("chunk_leaves", C.c_uint32),
        ("decode_precision", C.c_uint32),
        ("encode_precision", C.c_uint32),
        ("reserved0", C.c_uint32),
        ("onnx_encoder_path", C.c_char_p),
        ("onnx_decoder_path", C.c_char_p),
        ("reserved", C.c_uint32 * 2),
    ]


_lib = None


def load_library() -> C.CDLL:
    """Loads the in-tree shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libvqvdb_b200.so is not built (python -m vqvdb_b200.build); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    L.vqvdb_b200_create.argtypes = [C.POINTER(_Config), C.POINTER(C.c_void_p)]
    L.vqvdb_b200_create.restype = C.c_int
    L.vqvdb_b200_destroy.argtypes = [C.c_void_p]
    L.vqvdb_b200_destroy.restype = None
    L.vqvdb_b200_latent_shape.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    L.vqvdb_b200_in_channels.argtypes = [C.c_void_p]
    L.vqvdb_b200_num_embeddings.argtypes = [C.c_void_p]
    for fn in ("vqvdb_b200_encode", "vqvdb_b200_decode"):
        getattr(L, fn).argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        getattr(L, fn).restype = C.c_int
    for fn in ("vqvdb_b200_encode_device", "vqvdb_b200_decode_device"):
        getattr(L, fn).argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        getattr(L, fn).restype = C.c_int
    L.vqvdb_b200_debug_decode_tap.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.vqvdb_b200_debug_decode_tap.restype = C.c_int
    L.vqvdb_b200_debug_encode_tap.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.vqvdb_b200_debug_encode_tap.restype = C.c_int
    L.vqvdb_b200_encode_path.argtypes = [C.c_void_p]
    L.vqvdb_b200_encode_path.restype = C.c_char_p
    L.vqvdb_b200_peer_buffer_create.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p), C.c_char_p]
    L.vqvdb_b200_peer_buffer_create.restype = C.c_int
    L.vqvdb_b200_peer_buffer_open.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]
    L.vqvdb_b200_peer_buffer_open.restype = C.c_int
    L.vqvdb_b200_peer_buffer_close.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.vqvdb_b200_peer_buffer_close.restype = C.c_int
    L.vqvdb_b200_convert_onnx.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
    L.vqvdb_b200_convert_onnx.restype = C.c_int
    L.vqvdb_b200_debug_fold_decoder_tail.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p]
    L.vqvdb_b200_debug_fold_decoder_tail.restype = C.c_int
    L.vqvdb_b200_debug_fold_encoder_vq.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.vqvdb_b200_debug_fold_encoder_vq.restype = C.c_int
    L.vqvdb_b200_synchronize.argtypes = [C.c_void_p]
    L.vqvdb_b200_kernel_launches.argtypes = [C.c_void_p]
    L.vqvdb_b200_kernel_launches.restype = C.c_uint64
    L.vqvdb_b200_decode_path.argtypes = [C.c_void_p]
    L.vqvdb_b200_decode_path.restype = C.c_char_p
    L.vqvdb_b200_last_error.argtypes = [C.c_void_p]
    L.vqvdb_b200_last_error.restype = C.c_char_p
    L.vqvdb_b200_version.restype = C.c_char_p
    _lib = L
    return L


# ---- mirror of the reference's boundary types (IVQVAECodec.hpp:21-89) ----

class BackendType(enum.Enum):
    LibTorch = 0   # reference backends: not provided by this package
    ONNX = 1
    B200 = 2       # the backend this package adds


class DataType(enum.Enum):
    FLOAT32 = 0
    UINT8 = 1


class EmbeddedModel:
    """Tag: use the model linked into the library (reference: IVQVAECodec.hpp:27)."""


@dataclass
class OnnxModelPaths:
    """The reference's two-graph source (IVQVAECodec.hpp:29-33); only the initializers (weights) are read."""
    encoder: str
    decoder: str


def fold_decoder_tail(weights_path: str = ""):
    """The folded decoder tail (up_conv -> PixelShuffle3D -> final as one conv) the *_fold decode path uses:
    (weights [64][64][3][3][3], bias [64]) as fp32 numpy arrays.  Host-only."""
    import numpy as np
    L = load_library()
    w = np.empty((64, 64, 3, 3, 3), dtype=np.float32)
    b = np.empty((64,), dtype=np.float32)
    rc = L.vqvdb_b200_debug_fold_decoder_tail(os.fspath(weights_path).encode(), w.ctypes.data, b.ctypes.data)
    if rc != 0:
        raise RuntimeError("vqvdb_b200_debug_fold_decoder_tail failed (%d): %s" % (rc, L.vqvdb_b200_last_error(None).decode()))
    return w, b


def fold_encoder_vq(weights_path: str = "", channels: int = 32):
    """The tensor-core encoders' `proj x codebook` fold: (M [256][channels], esq [256], norm [257]) as fp32 numpy arrays;
    channels = 32 for the float model, 128 for the vec3 model (norm[:3] = the bound's constants there).  Host-only."""
    import numpy as np
    L = load_library()
    m = np.empty((256, channels), dtype=np.float32)
    esq = np.empty((256,), dtype=np.float32)
    norm = np.empty((257,), dtype=np.float32)
    rc = L.vqvdb_b200_debug_fold_encoder_vq(os.fspath(weights_path).encode(), m.ctypes.data, esq.ctypes.data, norm.ctypes.data)
    if rc != 0:
        raise RuntimeError("vqvdb_b200_debug_fold_encoder_vq failed (%d): %s" % (rc, L.vqvdb_b200_last_error(None).decode()))
    return m, esq, norm


def convert_onnx(encoder_onnx: str, decoder_onnx: str, out_pack: str) -> None:
    """encoder.onnx + decoder.onnx -> VQVDBW01 weight pack (host-only, no device needed)."""
    L = load_library()
    rc = L.vqvdb_b200_convert_onnx(os.fspath(encoder_onnx).encode(), os.fspath(decoder_onnx).encode(), os.fspath(out_pack).encode())
    if rc != 0:
        raise RuntimeError("vqvdb_b200_convert_onnx failed (%d): %s" % (rc, L.vqvdb_b200_last_error(None).decode()))


@dataclass
class CodecConfig:
    class Device(enum.Enum):
        CPU = 0
        CUDA = 1

    device: "CodecConfig.Device" = None  # type: ignore[assignment]
    source: Union[EmbeddedModel, "OnnxModelPaths", str, os.PathLike] = field(default_factory=EmbeddedModel)
    device_index: int = 0            # extension: which GPU (the reference hard-codes 0)
    chunk_leaves: int = 0            # extension: pipeline chunk of the host-pointer calls
    decode_precision: str = "default"  # "default" | "fp32" (checking path) | "bf16_tc" (tcgen05, the default)
    encode_precision: str = "default"  # "default" | "fp32" (FFMA) | "fp16x2_tc" (tcgen05, split-fp16 operands)

    def __post_init__(self):
        if self.device is None:
            self.device = CodecConfig.Device.CPU


@dataclass
class TensorView:
    """Non-owning view (IVQVAECodec.hpp:49-53): `data` is a C-contiguous numpy array."""
    data: np.ndarray
    shape: Sequence[int]
    dtype: DataType


@dataclass
class Tensor:
    """Owning result (IVQVAECodec.hpp:61-80)."""
    buffer: np.ndarray
    shape: Sequence[int]
    dtype: DataType

    def getData(self) -> np.ndarray:
        return self.buffer


_PRECISION = {"default": 0, "fp32": 1, "bf16_tc": 2}
_ENC_PRECISION = {"default": 0, "fp32": 1, "fp16x2_tc": 2}


class B200Codec:
    """IVQVAECodec implementation over the C-ABI.  Construct through IVQVAECodec.create()."""

    def __init__(self, config: CodecConfig):
        L = load_library()
        if config.device != CodecConfig.Device.CUDA:
            # The reference's TorchBackend falls back to CPU (TorchBackend.cpp:64-72); this backend must not.
            raise RuntimeError("B200 backend requires CodecConfig.Device.CUDA; there is no CPU path")
        cfg = _Config()
        cfg.struct_size = C.sizeof(_Config)
        cfg.device = int(config.device_index)
        cfg.chunk_leaves = int(config.chunk_leaves)
        cfg.decode_precision = _PRECISION[config.decode_precision]
        cfg.encode_precision = _ENC_PRECISION[config.encode_precision]
        self._keep = None
        if isinstance(config.source, EmbeddedModel):
            pass
        elif isinstance(config.source, OnnxModelPaths):
            self._keep = (os.fspath(config.source.encoder).encode(), os.fspath(config.source.decoder).encode())
            cfg.onnx_encoder_path, cfg.onnx_decoder_path = self._keep
        elif isinstance(config.source, (bytes, bytearray)):
            self._keep = C.create_string_buffer(bytes(config.source), len(config.source))
            cfg.weights_data = C.cast(self._keep, C.c_void_p)
            cfg.weights_size = len(config.source)
        else:
            self._keep = os.fspath(config.source).encode()
            cfg.weights_path = self._keep
        h = C.c_void_p()
        rc = L.vqvdb_b200_create(C.byref(cfg), C.byref(h))
        if rc != 0 or not h:
            raise RuntimeError("vqvdb_b200_create failed (%d): %s" % (rc, L.vqvdb_b200_last_error(None).decode()))
        self._L = L
        self._h = h
        shp = (C.c_int64 * 3)()
        L.vqvdb_b200_latent_shape(h, shp)
        self._latent = [int(v) for v in shp]
        self.channels = int(L.vqvdb_b200_in_channels(h))

    # -- lifecycle --
    def close(self):
        if getattr(self, "_h", None):
            self._L.vqvdb_b200_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise RuntimeError("%s failed (%d): %s" % (what, rc, self._L.vqvdb_b200_last_error(self._h).decode()))

    # -- IVQVAECodec --
    def getLatentShape(self):
        return list(self._latent)

    def encode(self, leafBatch: TensorView) -> Tensor:
        if leafBatch.dtype != DataType.FLOAT32:
            raise RuntimeError("encode expects FLOAT32 data.")     # TorchBackend.cpp:134-136
        x = leafBatch.data
        n = int(leafBatch.shape[0])
        # the raw pointer is all the C ABI sees: shape, dtype and size are checked here (the reference builds a torch
        # tensor from the view's shape and fails cleanly on a mismatch, TorchBackend.cpp:141-149)
        if list(leafBatch.shape[1:]) != [self.channels, 8, 8, 8]:
            raise RuntimeError("encode: expected shape [B,%d,8,8,8], got %s" % (self.channels, list(leafBatch.shape)))
        if not isinstance(x, np.ndarray) or x.dtype != np.float32 or x.size != n * self.channels * 512:
            raise RuntimeError("encode: data must be a float32 array of %d elements" % (n * self.channels * 512))
        out = np.empty((n, *self._latent), dtype=np.uint8)
        self.encode_into(x, n, out)
        return Tensor(out, list(out.shape), DataType.UINT8)

    def decode(self, indices: TensorView) -> Tensor:
        if indices.dtype != DataType.UINT8:
            raise RuntimeError("decode expects UINT8 data.")       # TorchBackend.cpp:167-169
        n = int(indices.shape[0])
        if list(indices.shape[1:]) != self._latent:
            raise RuntimeError("decode: expected shape [B,%d,%d,%d], got %s" % (*self._latent, list(indices.shape)))
        if not isinstance(indices.data, np.ndarray) or indices.data.dtype != np.uint8 or indices.data.size != n * 64:
            raise RuntimeError("decode: data must be a uint8 array of %d elements" % (n * 64))
        out = np.empty((n, self.channels, 8, 8, 8), dtype=np.float32)
        self.decode_into(indices.data, n, out)
        return Tensor(out, list(out.shape), DataType.FLOAT32)

    # -- raw-pointer forms (host numpy arrays or integer addresses) --
    @staticmethod
    def _addr(a) -> int:
        if isinstance(a, np.ndarray):
            if not a.flags["C_CONTIGUOUS"]:
                raise ValueError("array must be C-contiguous")
            return a.ctypes.data
        if hasattr(a, "data_ptr"):
            return int(a.data_ptr())
        return int(a)

    def encode_into(self, host_leaves, n: int, host_indices):
        self._check(self._L.vqvdb_b200_encode(self._h, self._addr(host_leaves), n, self._addr(host_indices)), "encode")

    def decode_into(self, host_indices, n: int, host_voxels):
        self._check(self._L.vqvdb_b200_decode(self._h, self._addr(host_indices), n, self._addr(host_voxels)), "decode")

    def encode_device(self, dev_leaves, n: int, dev_indices, stream: int = 0):
        self._check(self._L.vqvdb_b200_encode_device(self._h, self._addr(dev_leaves), n, self._addr(dev_indices),
                                                     C.c_void_p(stream)), "encode_device")

    def decode_device(self, dev_indices, n: int, dev_voxels, stream: int = 0):
        self._check(self._L.vqvdb_b200_decode_device(self._h, self._addr(dev_indices), n, self._addr(dev_voxels),
                                                     C.c_void_p(stream)), "decode_device")

    def debug_decode_tap(self, dev_indices, n: int, stage: int, dev_tap, dev_voxels, stream: int = 0):
        self._check(self._L.vqvdb_b200_debug_decode_tap(self._h, self._addr(dev_indices), n, stage, self._addr(dev_tap),
                                                        self._addr(dev_voxels), C.c_void_p(stream)), "debug_decode_tap")

    def debug_encode_tap(self, dev_leaves, n: int, stage: int, dev_tap, dev_indices, stream: int = 0):
        self._check(self._L.vqvdb_b200_debug_encode_tap(self._h, self._addr(dev_leaves), n, stage, self._addr(dev_tap),
                                                        self._addr(dev_indices), C.c_void_p(stream)), "debug_encode_tap")

    # -- multi-GPU reassembly buffer (CUDA IPC): see include/vqvdb_b200.h --
    def peer_buffer_create(self, nbytes: int):
        """cudaMalloc on this codec's device; returns (device pointer, 64-byte IPC handle)."""
        ptr = C.c_void_p()
        handle = C.create_string_buffer(64)
        self._check(self._L.vqvdb_b200_peer_buffer_create(self._h, int(nbytes), C.byref(ptr), handle), "peer_buffer_create")
        return int(ptr.value), handle.raw

    def peer_buffer_open(self, handle: bytes) -> int:
        ptr = C.c_void_p()
        self._check(self._L.vqvdb_b200_peer_buffer_open(self._h, C.create_string_buffer(handle, 64), C.byref(ptr)), "peer_buffer_open")
        return int(ptr.value)

    def peer_buffer_close(self, ptr: int, opened: bool):
        self._check(self._L.vqvdb_b200_peer_buffer_close(self._h, C.c_void_p(ptr), 1 if opened else 0), "peer_buffer_close")

    def synchronize(self):
        self._check(self._L.vqvdb_b200_synchronize(self._h), "synchronize")

    @property
    def kernel_launches(self) -> int:
        return int(self._L.vqvdb_b200_kernel_launches(self._h))

    @property
    def decode_path(self) -> str:
        return self._L.vqvdb_b200_decode_path(self._h).decode()

    @property
    def encode_path(self) -> str:
        return self._L.vqvdb_b200_encode_path(self._h).decode()


class IVQVAECodec:
    """Factory with the reference's contract (IVQVAECodec.cpp:76-110): never raises, logs and
    returns None on failure.  Only BackendType.B200 is served by this package."""

    @staticmethod
    def create(config: CodecConfig, type: BackendType = BackendType.B200) -> Optional[B200Codec]:
        import sys
        try:
            if type != BackendType.B200:
                raise RuntimeError("Requested backend type is not available or disabled in the build configuration.")
            return B200Codec(config)
        except Exception as e:  # noqa: BLE001 — mirrors the reference's catch-all
            print("Failed to create VQ-VAE backend: %s" % e, file=sys.stderr)
            return None
