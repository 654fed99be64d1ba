"""TEST INFRASTRUCTURE — worker process hosting the reference's own backend.

Loads oracle/_ref/libvqvdb_ref.so (reference TorchBackend.cpp + IVQVAECodec.cpp compiled
unmodified, plus ref_shim.cpp) WITHOUT importing the Python torch package, and serves
encode/decode requests from oracle/pyoracle.py:RefCodec over stdin/stdout.

Protocol (lines starting with "@@" are ours; anything else on stdout is the reference's
own std::cout output and is ignored by the parent):
    -> "encode <in.bin> <n_leaves> <out.bin>"     float32 [n,1,8,8,8] -> uint8 [n,4,4,4]
    -> "decode <in.bin> <n_leaves> <out.bin>"     uint8 [n,4,4,4]     -> float32 [n,1,8,8,8]
    -> "bench <encode|decode|roundtrip> <in.bin> <n_leaves> <batch> <steps> <warmup>"
    <- "@@ok <seconds> [...]" | "@@err <message>"
The timed interval is only the reference's encode()/decode() calls, batch by batch, host
pointers in and out, exactly like the loop in orchestrator/VQVAECodec.cpp:108-127,166-196.
"""
import ctypes as C
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    device = sys.argv[1] if len(sys.argv) > 1 else "cpu"
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    tdir = None
    for p in sys.path:
        cand = os.path.join(p, "torch", "lib")
        if os.path.isdir(cand):
            tdir = cand
            break
    if tdir:  # the shim has an rpath to the build-time torch/lib; this covers a relocated venv
        for name in ("libc10.so", "libtorch_cpu.so"):
            try:
                C.CDLL(os.path.join(tdir, name), mode=C.RTLD_GLOBAL)
            except OSError:
                pass
    L = C.CDLL(os.path.join(HERE, "_ref", "libvqvdb_ref.so"))
    L.vqvdb_ref_create.restype = C.c_void_p
    L.vqvdb_ref_create.argtypes = [C.c_int, C.c_int]
    L.vqvdb_ref_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    L.vqvdb_ref_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    L.vqvdb_ref_latent_shape.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.vqvdb_ref_last_error.restype = C.c_char_p
    if device == "cuda" and not L.vqvdb_ref_cuda_available():
        print("@@err cuda not available to the reference backend", flush=True)
        return 1
    h = L.vqvdb_ref_create(1 if device == "cuda" else 0, threads)
    if not h:
        print("@@err %s" % L.vqvdb_ref_last_error().decode(), flush=True)
        return 1
    buf = (C.c_int64 * 8)()
    nd = L.vqvdb_ref_latent_shape(h, buf, 8)
    print("@@ready %d %s" % (L.vqvdb_ref_threads(), " ".join(str(buf[i]) for i in range(nd))), flush=True)

    def enc(x, out):
        if L.vqvdb_ref_encode(h, x.ctypes.data, x.shape[0], out.ctypes.data) != 0:
            raise RuntimeError(L.vqvdb_ref_last_error().decode())

    def dec(x, out):
        if L.vqvdb_ref_decode(h, x.ctypes.data, x.shape[0], out.ctypes.data) != 0:
            raise RuntimeError(L.vqvdb_ref_last_error().decode())

    for line in sys.stdin:
        parts = line.split()
        if not parts:
            continue
        try:
            if parts[0] == "quit":
                break
            if parts[0] == "encode":
                n = int(parts[2])
                x = np.fromfile(parts[1], dtype=np.float32).reshape(n, 1, 8, 8, 8)
                out = np.empty((n, 4, 4, 4), dtype=np.uint8)
                t = time.perf_counter()
                if n:
                    enc(x, out)
                dt = time.perf_counter() - t
                out.tofile(parts[3])
                print("@@ok %.9f" % dt, flush=True)
            elif parts[0] == "decode":
                n = int(parts[2])
                x = np.fromfile(parts[1], dtype=np.uint8).reshape(n, 4, 4, 4)
                out = np.empty((n, 1, 8, 8, 8), dtype=np.float32)
                t = time.perf_counter()
                if n:
                    dec(x, out)
                dt = time.perf_counter() - t
                out.tofile(parts[3])
                print("@@ok %.9f" % dt, flush=True)
            elif parts[0] == "bench":
                what, path, n, batch, steps, warmup = parts[1], parts[2], int(parts[3]), int(parts[4]), int(parts[5]), int(parts[6])
                if what == "decode":
                    idx = np.fromfile(path, dtype=np.uint8).reshape(n, 4, 4, 4)
                    x = None
                else:
                    x = np.fromfile(path, dtype=np.float32).reshape(n, 1, 8, 8, 8)
                    idx = np.empty((n, 4, 4, 4), dtype=np.uint8)
                vox = np.empty((n, 1, 8, 8, 8), dtype=np.float32)

                def step():
                    for lo in range(0, n, batch):
                        hi = min(n, lo + batch)
                        if what in ("encode", "roundtrip"):
                            enc(x[lo:hi], idx[lo:hi])
                        if what in ("decode", "roundtrip"):
                            dec(idx[lo:hi], vox[lo:hi])

                for _ in range(warmup):
                    step()
                times = []
                for _ in range(steps):
                    t = time.perf_counter()
                    step()
                    times.append(time.perf_counter() - t)
                print("@@ok %.9f %s" % (sum(times), " ".join("%.9f" % v for v in times)), flush=True)
            else:
                print("@@err unknown op %s" % parts[0], flush=True)
        except Exception as e:  # noqa: BLE001
            print("@@err %s" % str(e).replace("\n", " "), flush=True)
    L.vqvdb_ref_destroy.argtypes = [C.c_void_p]
    L.vqvdb_ref_destroy(h)
    return 0


if __name__ == "__main__":
    sys.exit(main())
