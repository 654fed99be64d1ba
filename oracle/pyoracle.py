"""TEST INFRASTRUCTURE — ctypes access to the two checkers.

  COracle   oracle/libvqvae_oracle.so  plain-C restatement (vqvae_oracle.c)
  RefCodec  oracle/_ref/libvqvdb_ref.so  the reference's own LibTorch backend
            (src/backends/torch/TorchBackend.cpp compiled unmodified + ref_shim.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's baseline legs import this
module.  Nothing in vqvdb_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
ORACLE_SO = os.path.join(HERE, "libvqvae_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libvqvdb_ref.so")
FLOAT_PACK = os.path.join(REPO, "vqvdb_b200", "weights", "vqvae_float.vqw")
VEC3_PACK = os.path.join(REPO, "vqvdb_b200", "weights", "vqvae_vec3_seed0.vqw")


def build_oracle(force: bool = False) -> str:
    if force or not os.path.exists(ORACLE_SO) or (
            os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(HERE, "vqvae_oracle.c"))):
        subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=subprocess.DEVNULL)
    return ORACLE_SO


class COracle:
    def __init__(self, pack: str = FLOAT_PACK, threads: int = 0):
        self.lib = C.CDLL(build_oracle())
        L = self.lib
        L.vqo_load.restype = C.c_void_p
        L.vqo_load.argtypes = [C.c_char_p]
        L.vqo_free.argtypes = [C.c_void_p]
        L.vqo_in_channels.argtypes = [C.c_void_p]
        L.vqo_set_threads.argtypes = [C.c_int]
        for fn in ("vqo_encode",):
            getattr(L, fn).argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.vqo_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.vqo_encode_latents.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.vqo_decode_tap.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
        L.vqo_encode_tap.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
        self.h = L.vqo_load(pack.encode())
        if not self.h:
            raise RuntimeError("vqo_load failed for %s" % pack)
        self.channels = L.vqo_in_channels(self.h)
        self.threads = L.vqo_set_threads(threads)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.vqo_free(self.h)
            self.h = None

    def encode(self, leaves: np.ndarray, with_margins: bool = False):
        x = np.ascontiguousarray(leaves, dtype=np.float32)
        n = x.shape[0]
        idx = np.empty((n, 4, 4, 4), dtype=np.uint8)
        mar = np.empty((n, 4, 4, 4), dtype=np.float32) if with_margins else None
        rc = self.lib.vqo_encode(self.h, x.ctypes.data, n, idx.ctypes.data,
                                 mar.ctypes.data if with_margins else None)
        assert rc == 0
        return (idx, mar) if with_margins else idx

    def latents(self, leaves: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(leaves, dtype=np.float32)
        n = x.shape[0]
        z = np.empty((n, 128, 4, 4, 4), dtype=np.float32)
        assert self.lib.vqo_encode_latents(self.h, x.ctypes.data, n, z.ctypes.data) == 0
        return z

    def encode_tap(self, leaves: np.ndarray, stage: int, c0: int = 16, c1: int = 32) -> np.ndarray:
        x = np.ascontiguousarray(leaves, dtype=np.float32)
        n = x.shape[0]
        shape = (n, c0, 8, 8, 8) if stage in (0, 1, 6) else (n, c1, 4, 4, 4)
        out = np.empty(shape, dtype=np.float32)
        assert self.lib.vqo_encode_tap(self.h, x.ctypes.data, n, stage, out.ctypes.data) == 0
        return out

    def decode(self, indices: np.ndarray) -> np.ndarray:
        idx = np.ascontiguousarray(indices, dtype=np.uint8)
        n = idx.shape[0]
        out = np.empty((n, self.channels, 8, 8, 8), dtype=np.float32)
        assert self.lib.vqo_decode(self.h, idx.ctypes.data, n, out.ctypes.data) == 0
        return out

    def decode_tap(self, indices: np.ndarray, stage: int, width: int = 64) -> np.ndarray:
        idx = np.ascontiguousarray(indices, dtype=np.uint8)
        n = idx.shape[0]
        shape = (n, 32, 8, 8, 8) if stage == 3 else (n, width, 4, 4, 4)
        out = np.empty(shape, dtype=np.float32)
        assert self.lib.vqo_decode_tap(self.h, idx.ctypes.data, n, stage, out.ctypes.data) == 0
        return out


def ref_available() -> bool:
    return os.path.exists(REF_SO)


class RefCodec:
    """The reference's IVQVAECodec (LibTorch backend) behind oracle/ref_shim.cpp.

    Runs in a worker process (oracle/ref_worker.py) that never imports the Python
    torch package: the reference prints through std::cout and that segfaults when
    libtorch_python is already resident in the same process.  Arrays travel through
    /dev/shm files; the worker times each call itself (seconds returned in
    `last_seconds`), so pipe/file overhead is outside the measured interval.
    """

    def __init__(self, device: str = "cpu", threads: int = 0):
        import subprocess
        import sys
        import tempfile
        if not ref_available():
            raise RuntimeError("oracle/_ref/libvqvdb_ref.so not built (make -C oracle ref)")
        self.device = device
        self.tmp = tempfile.mkdtemp(prefix="vqvdb_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        self.proc = subprocess.Popen(
            [sys.executable, os.path.join(HERE, "ref_worker.py"), device, str(threads)],
            stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True, bufsize=1)
        hello = self._readline()
        if not hello.startswith("ready"):
            raise RuntimeError("reference worker failed to start: %s" % hello)
        parts = hello.split()
        self.threads = int(parts[1])
        self._latent = [int(v) for v in parts[2:]]
        self.last_seconds = 0.0

    def _readline(self) -> str:
        while True:
            line = self.proc.stdout.readline()
            if not line:
                raise RuntimeError("reference worker died (exit %s)" % self.proc.poll())
            line = line.strip()
            if line.startswith("@@"):  # protocol lines; everything else is the reference's own chatter
                return line[2:].strip()

    def _call(self, op: str, arr: np.ndarray, out_shape, out_dtype) -> np.ndarray:
        src = os.path.join(self.tmp, "in.bin")
        dst = os.path.join(self.tmp, "out.bin")
        arr.tofile(src)
        self.proc.stdin.write("%s %s %d %s\n" % (op, src, arr.shape[0], dst))
        self.proc.stdin.flush()
        resp = self._readline()
        if not resp.startswith("ok"):
            raise RuntimeError("reference %s failed: %s" % (op, resp))
        self.last_seconds = float(resp.split()[1])
        return np.fromfile(dst, dtype=out_dtype).reshape(out_shape)

    def close(self):
        if getattr(self, "proc", None) and self.proc.poll() is None:
            try:
                self.proc.stdin.write("quit\n")
                self.proc.stdin.flush()
                self.proc.wait(timeout=10)
            except Exception:
                self.proc.kill()
        self.proc = None
        if getattr(self, "tmp", None):
            import shutil
            shutil.rmtree(self.tmp, ignore_errors=True)
            self.tmp = None

    __del__ = close

    def latent_shape(self):
        return list(self._latent)

    def encode(self, leaves: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(leaves, dtype=np.float32)
        return self._call("encode", x, (x.shape[0], 4, 4, 4), np.uint8)

    def encode_tap(self, leaves: np.ndarray, stage: int, c0: int = 16, c1: int = 32) -> np.ndarray:
        x = np.ascontiguousarray(leaves, dtype=np.float32)
        n = x.shape[0]
        shape = (n, c0, 8, 8, 8) if stage in (0, 1, 6) else (n, c1, 4, 4, 4)
        out = np.empty(shape, dtype=np.float32)
        assert self.lib.vqo_encode_tap(self.h, x.ctypes.data, n, stage, out.ctypes.data) == 0
        return out

    def decode(self, indices: np.ndarray) -> np.ndarray:
        idx = np.ascontiguousarray(indices, dtype=np.uint8)
        return self._call("decode", idx, (idx.shape[0], 1, 8, 8, 8), np.float32)


# ---------------------------------------------------------------------------------------------
# Format oracle: the reference's own VDBStreamWriter / VDBStreamReader (src/Utils/VQVDB_Reader.cpp), compiled
# unmodified against oracle/stub/openvdb/Types.h into oracle/_ref/libvqvdb_fmt.so (oracle/Makefile: fmt).
# ---------------------------------------------------------------------------------------------
FMT_LIB = os.path.join(HERE, "_ref", "libvqvdb_fmt.so")


def fmt_available() -> bool:
    return os.path.exists(FMT_LIB)


class RefFormat:
    """Write / read .vqvdb v3 files with the reference's own code.  Grids are (name, origins int32 [n,3],
    indices uint8 [n,64], transform float32 [16]) tuples."""

    def __init__(self):
        import ctypes as C
        self.C = C
        L = C.CDLL(FMT_LIB)
        L.reffmt_last_error.restype = C.c_char_p
        L.reffmt_write.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_int64]
        L.reffmt_reader_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.reffmt_reader_close.argtypes = [C.c_void_p]
        L.reffmt_reader_close.restype = None
        L.reffmt_reader_has_next_grid.argtypes = [C.c_void_p]
        L.reffmt_reader_has_next.argtypes = [C.c_void_p]
        L.reffmt_reader_next_grid.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int),
                                              C.POINTER(C.c_int64), C.POINTER(C.c_uint32)]
        L.reffmt_reader_next_batch.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.reffmt_reader_next_batch.restype = C.c_int64
        self.L = L

    def write(self, path, grids, num_embeddings=256, latent_shape=(4, 4, 4), batch=64):
        C, L = self.C, self.L
        names = (C.c_char_p * max(1, len(grids)))(*[g[0].encode() for g in grids])
        tr = (np.ascontiguousarray(np.stack([np.asarray(g[3], np.float32).reshape(16) for g in grids]))
              if grids else np.zeros((1, 16), np.float32))
        counts = np.array([len(g[1]) for g in grids], dtype=np.int64)
        og = [np.ascontiguousarray(g[1], dtype=np.int32) for g in grids]
        ix = [np.ascontiguousarray(g[2], dtype=np.uint8).reshape(len(g[1]), -1) for g in grids]
        lat = np.array(latent_shape, dtype=np.int64)
        po = (C.c_void_p * max(1, len(grids)))(*[a.ctypes.data for a in og])
        pi = (C.c_void_p * max(1, len(grids)))(*[a.ctypes.data for a in ix])
        rc = L.reffmt_write(os.fspath(path).encode(), len(grids), names, tr.ctypes.data, lat.ctypes.data, len(latent_shape),
                            num_embeddings, counts.ctypes.data, po, pi, batch)
        if rc != 0:
            raise RuntimeError("reference writer: " + L.reffmt_last_error().decode())

    def read(self, path, batch=1 << 20, stop_on_empty_batch=True):
        """Every grid through VDBStreamReader exactly as the reference's decompress loop drives it
        (VQVAECodec.cpp:147-196): `while hasNextGrid: meta; while hasNext: nextBatch(batch)`."""
        C, L = self.C, self.L
        h = C.c_void_p()
        if L.reffmt_reader_open(os.fspath(path).encode(), C.byref(h)) != 0:
            raise RuntimeError("reference reader: " + L.reffmt_last_error().decode())
        out = []
        try:
            while L.reffmt_reader_has_next_grid(h):
                name = C.create_string_buffer(4096)
                tr = np.zeros(16, np.float32)
                lat = np.zeros(8, np.int64)
                rank, n, k = C.c_int(), C.c_int64(), C.c_uint32()
                if L.reffmt_reader_next_grid(h, name, 4096, tr.ctypes.data, lat.ctypes.data, C.byref(rank), C.byref(n), C.byref(k)) != 0:
                    raise RuntimeError("reference reader: " + L.reffmt_last_error().decode())
                bb = int(np.prod(lat[:rank.value])) if rank.value else 1
                og_parts, ix_parts = [], []
                while L.reffmt_reader_has_next(h):
                    og = np.empty((batch, 3), np.int32)
                    ix = np.empty((batch, bb), np.uint8)
                    got = L.reffmt_reader_next_batch(h, batch, og.ctypes.data, ix.ctypes.data)
                    if got < 0:
                        raise RuntimeError("reference reader: " + L.reffmt_last_error().decode())
                    if got == 0 and stop_on_empty_batch:
                        break
                    og_parts.append(og[:got].copy())
                    ix_parts.append(ix[:got].copy())
                og = np.concatenate(og_parts) if og_parts else np.zeros((0, 3), np.int32)
                ix = np.concatenate(ix_parts) if ix_parts else np.zeros((0, bb), np.uint8)
                out.append(dict(name=name.value.decode(), origins=og, indices=ix, transform=tr, latent_shape=[int(v) for v in lat[:rank.value]],
                                total_blocks=int(n.value), num_embeddings=int(k.value)))
        finally:
            L.reffmt_reader_close(h)
        return out
