"""ctypes binding of the host-layer C ABI (include/vqvdb_b200_host.h): .vqvdb v3 files and the
compress / decompress batch loops of the reference's orchestrator (VQVAECodec.cpp:78-208) over flat leaf arrays."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(HERE, "libvqvdb_b200_host.so")

HOST_EXPORTS = [
    "vqvdb_host_write_file", "vqvdb_host_reader_open", "vqvdb_host_reader_close", "vqvdb_host_reader_num_grids",
    "vqvdb_host_reader_num_embeddings", "vqvdb_host_reader_next_grid", "vqvdb_host_reader_next_batch",
    "vqvdb_host_compress", "vqvdb_host_decompress", "vqvdb_host_last_error",
    "vqvdb_host_backend_create", "vqvdb_host_backend_destroy", "vqvdb_host_backend_encode", "vqvdb_host_backend_decode",
    "vqvdb_host_backend_result", "vqvdb_host_backend_encode_into", "vqvdb_host_backend_decode_into",
    "vqvdb_host_backend_roundtrip_batched",
    "vqvdb_host_orchestrator_accepts",
]

_lib = None


def load_host_library() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise RuntimeError("libvqvdb_b200_host.so is not built (make -C vqvdb_b200/cpp)")
        L = C.CDLL(HOST_LIB_PATH)
        L.vqvdb_host_last_error.restype = C.c_char_p
        L.vqvdb_host_reader_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.vqvdb_host_reader_close.argtypes = [C.c_void_p]
        L.vqvdb_host_reader_close.restype = None
        L.vqvdb_host_reader_num_grids.argtypes = [C.c_void_p]
        L.vqvdb_host_reader_num_embeddings.argtypes = [C.c_void_p]
        L.vqvdb_host_reader_num_embeddings.restype = C.c_uint32
        L.vqvdb_host_reader_next_grid.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
        L.vqvdb_host_reader_next_batch.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.vqvdb_host_reader_next_batch.restype = C.c_int64
        L.vqvdb_host_write_file.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                            C.c_void_p, C.c_void_p, C.c_void_p]
        L.vqvdb_host_compress.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_int64]
        L.vqvdb_host_decompress.argtypes = [C.c_int, C.c_char_p, C.POINTER(C.c_int), C.c_void_p, C.c_int, C.c_void_p,
                                            C.c_void_p, C.c_int64, C.c_int]
        L.vqvdb_host_backend_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.vqvdb_host_backend_destroy.argtypes = [C.c_void_p]
        L.vqvdb_host_backend_destroy.restype = None
        for fn in ("vqvdb_host_backend_encode", "vqvdb_host_backend_decode"):
            getattr(L, fn).argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_double)]
        for fn in ("vqvdb_host_backend_encode_into", "vqvdb_host_backend_decode_into"):
            getattr(L, fn).argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_double)]
        L.vqvdb_host_backend_roundtrip_batched.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                                           C.POINTER(C.c_double)]
        L.vqvdb_host_orchestrator_accepts.argtypes = [C.c_int, C.c_char_p]
        L.vqvdb_host_backend_result.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.vqvdb_host_backend_result.restype = C.c_void_p
        _lib = L
    return _lib


@dataclass
class LeafGrid:
    """Flat result of the reference's leaf walk: one entry per active 8^3 leaf."""
    name: str
    origins: np.ndarray                      # int32 [n, 3]
    voxels: np.ndarray = None                # float32 [n, 512] (or [n,1,8,8,8]); None for index-only grids
    indices: np.ndarray = None               # uint8 [n, 64] (or [n,4,4,4])
    transform: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))


def _err(L, what):
    raise RuntimeError("%s: %s" % (what, L.vqvdb_host_last_error().decode()))


def _ptr_array(arrs: Sequence[np.ndarray]):
    return (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


def write_file(path: str, grids: List[LeafGrid], num_embeddings: int = 256, latent_shape=(4, 4, 4)):
    L = load_host_library()
    names = (C.c_char_p * len(grids))(*[g.name.encode() for g in grids])
    tr = (np.ascontiguousarray(np.stack([np.asarray(g.transform, np.float32).reshape(16) for g in grids]))
          if grids else np.zeros((1, 16), np.float32))
    counts = np.array([len(g.origins) for g in grids], dtype=np.int64)
    og = [np.ascontiguousarray(g.origins, dtype=np.int32) for g in grids]
    ix = [np.ascontiguousarray(g.indices, dtype=np.uint8).reshape(len(g.origins), int(np.prod(latent_shape))) for g in grids]
    lat = np.array(latent_shape, dtype=np.int64)
    rc = L.vqvdb_host_write_file(path.encode(), len(grids), names, tr.ctypes.data, lat.ctypes.data, num_embeddings,
                                 counts.ctypes.data, _ptr_array(og), _ptr_array(ix))
    if rc != 0:
        _err(L, "write_file")


def read_file(path: str, batch: int = 1 << 20) -> List[LeafGrid]:
    L = load_host_library()
    h = C.c_void_p()
    if L.vqvdb_host_reader_open(path.encode(), C.byref(h)) != 0:
        _err(L, "open")
    try:
        out = []
        for _ in range(L.vqvdb_host_reader_num_grids(h)):
            name = C.create_string_buffer(4096)
            tr = np.zeros(16, np.float32)
            lat = np.zeros(3, np.int64)
            n = C.c_int64()
            if L.vqvdb_host_reader_next_grid(h, name, 4096, tr.ctypes.data, lat.ctypes.data, C.byref(n)) != 0:
                _err(L, "next_grid")
            bb = int(np.prod(lat))
            og = np.empty((n.value, 3), np.int32)
            ix = np.empty((n.value, bb), np.uint8)
            done = 0
            while done < n.value:
                got = L.vqvdb_host_reader_next_batch(h, min(batch, n.value - done), og[done:].ctypes.data, ix[done:].ctypes.data)
                if got < 0:
                    _err(L, "next_batch")
                if got == 0:
                    break
                done += got
            out.append(LeafGrid(name.value.decode(), og, None, ix.reshape(n.value, *[int(v) for v in lat]), tr.reshape(4, 4)))
        return out
    finally:
        L.vqvdb_host_reader_close(h)


def compress(grids: List[LeafGrid], out_path: str, batch_size: int = 0, device: int = 0):
    """VQVAECodec::compress over flat grids through the B200 backend (GPU required)."""
    L = load_host_library()
    names = (C.c_char_p * len(grids))(*[g.name.encode() for g in grids])
    tr = np.ascontiguousarray(np.stack([np.asarray(g.transform, np.float32).reshape(16) for g in grids]))
    counts = np.array([len(g.origins) for g in grids], dtype=np.int64)
    og = [np.ascontiguousarray(g.origins, dtype=np.int32) for g in grids]
    vx = [np.ascontiguousarray(g.voxels, dtype=np.float32).reshape(len(g.origins), 512) for g in grids]
    rc = L.vqvdb_host_compress(device, out_path.encode(), len(grids), names, tr.ctypes.data, counts.ctypes.data,
                               _ptr_array(og), _ptr_array(vx), batch_size)
    if rc != 0:
        _err(L, "compress")


def decompress(in_path: str, batch_size: int = 0, device: int = 0, fp32_decode: bool = False) -> List[LeafGrid]:
    """VQVAECodec::decompress: returns grids with origins + decoded voxels (names/transforms via read_file)."""
    L = load_host_library()
    n = C.c_int()
    counts = np.zeros(256, np.int64)
    if L.vqvdb_host_decompress(device, in_path.encode(), C.byref(n), counts.ctypes.data, 256, None, None, batch_size, 0) != 0:
        _err(L, "decompress(size query)")
    og = [np.empty((int(counts[g]), 3), np.int32) for g in range(n.value)]
    vx = [np.empty((int(counts[g]), 512), np.float32) for g in range(n.value)]
    if L.vqvdb_host_decompress(device, in_path.encode(), C.byref(n), counts.ctypes.data, 256, _ptr_array(og), _ptr_array(vx),
                               batch_size, int(fp32_decode)) != 0:
        _err(L, "decompress")
    meta = read_file(in_path)
    return [LeafGrid(meta[g].name, og[g], vx[g].reshape(-1, 1, 8, 8, 8), meta[g].indices, meta[g].transform)
            for g in range(n.value)]


class HostBackend:
    """The C++ IVQVAECodec (B200Backend) driven through its virtual encode / decode, as the reference's orchestrator
    drives its backend: TensorView over caller memory in, owning Tensor out (GPU required)."""

    def __init__(self, device: int = 0):
        self.L = load_host_library()
        self.h = C.c_void_p()
        if self.L.vqvdb_host_backend_create(device, C.byref(self.h)) != 0:
            _err(self.L, "backend_create")

    def close(self):
        if getattr(self, "h", None):
            self.L.vqvdb_host_backend_destroy(self.h)
            self.h = None

    __del__ = close

    def _result(self, dtype, shape):
        nbytes = C.c_uint64()
        p = self.L.vqvdb_host_backend_result(self.h, C.byref(nbytes))
        buf = (C.c_char * nbytes.value).from_address(p)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)   # a view of the Tensor held by the handle

    def encode(self, leaves: np.ndarray):
        """-> (indices view [n,4,4,4] uint8, seconds of the virtual call)"""
        n = leaves.shape[0]
        sec = C.c_double()
        if self.L.vqvdb_host_backend_encode(self.h, leaves.ctypes.data, n, C.byref(sec)) != 0:
            _err(self.L, "backend_encode")
        return self._result(np.uint8, (n, 4, 4, 4)), sec.value

    def decode(self, indices: np.ndarray):
        n = indices.shape[0]
        sec = C.c_double()
        if self.L.vqvdb_host_backend_decode(self.h, indices.ctypes.data, n, C.byref(sec)) != 0:
            _err(self.L, "backend_decode")
        return self._result(np.float32, (n, 1, 8, 8, 8)), sec.value

    def encode_into(self, leaves: np.ndarray, indices_out: np.ndarray) -> float:
        sec = C.c_double()
        if self.L.vqvdb_host_backend_encode_into(self.h, leaves.ctypes.data, leaves.shape[0], indices_out.ctypes.data, C.byref(sec)) != 0:
            _err(self.L, "backend_encode_into")
        return sec.value

    def roundtrip_batched(self, leaves_addr: int, n: int, batch: int, indices_addr: int, voxels_addr: int) -> float:
        """encodeInto + decodeInto per `batch` leaves in a native loop over raw host addresses; -> seconds of the loop."""
        sec = C.c_double()
        if self.L.vqvdb_host_backend_roundtrip_batched(self.h, C.c_void_p(leaves_addr), n, batch, C.c_void_p(indices_addr),
                                                       C.c_void_p(voxels_addr), C.byref(sec)) != 0:
            _err(self.L, "backend_roundtrip_batched")
        return sec.value

    def decode_into(self, indices: np.ndarray, voxels_out: np.ndarray) -> float:
        sec = C.c_double()
        if self.L.vqvdb_host_backend_decode_into(self.h, indices.ctypes.data, indices.shape[0], voxels_out.ctypes.data, C.byref(sec)) != 0:
            _err(self.L, "backend_decode_into")
        return sec.value
