"""In-tree build of libvqvdb_b200.so (CUDA kernels + C-ABI) for sm_100a.

    python -m vqvdb_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  Objects go to build/ (git-ignored); the shared library is
written next to this file so it travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
# Tuning builds: VQVDB_B200_DEFINES="A=1,B=2" adds -DA=1 -DB=2 to every compile, VQVDB_B200_VARIANT=name writes
# build/variants/libvqvdb_b200_name.so instead of the product library (load it with VQVDB_B200_LIB=<path>; tools/ab_encode.py).
VARIANT = os.environ.get("VQVDB_B200_VARIANT", "")
DEFINES = ["-D" + d for d in os.environ.get("VQVDB_B200_DEFINES", "").split(",") if d]
OBJ = os.path.join(REPO, "build", "obj" + ("_" + VARIANT if VARIANT else ""))
LIB = os.path.join(REPO, "build", "variants", "libvqvdb_b200_%s.so" % VARIANT) if VARIANT else os.path.join(HERE, "libvqvdb_b200.so")
PACK = os.path.join(HERE, "weights", "vqvae_float.vqw")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-lineinfo", "--expt-relaxed-constexpr", "-Xptxas", "-v",
              "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unknown-pragmas"] + ARCH
CXX_FLAGS = ["-std=c++17", "-O2", "-fPIC", "-fvisibility=hidden", "-Wall"]
NVCC_FLAGS += DEFINES
CXX_FLAGS += DEFINES


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _sources():
    cu = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    cpp = sorted(f for f in os.listdir(CSRC) if f.endswith(".cpp"))
    asm = sorted(f for f in os.listdir(CSRC) if f.endswith(".S"))
    return cu, cpp, asm


def _deps_mtime() -> float:
    latest = os.path.getmtime(PACK)
    for root in (CSRC, os.path.join(REPO, "include")):
        for f in os.listdir(root):
            latest = max(latest, os.path.getmtime(os.path.join(root, f)))
    return max(latest, os.path.getmtime(os.path.abspath(__file__)))


def _run(cmd, verbose, log):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log.append("$ " + " ".join(cmd) + "\n" + p.stdout)
    if p.returncode != 0:
        raise RuntimeError("build step failed:\n$ %s\n%s" % (" ".join(cmd), p.stdout))
    if verbose:
        print(p.stdout, end="")


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    nvcc = _nvcc()
    cu, cpp, asm = _sources()
    log: list[str] = []
    jobs = []
    for f in cu:
        jobs.append([nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, f), "-o", os.path.join(OBJ, f + ".o")])
    for f in cpp:
        jobs.append(["g++"] + CXX_FLAGS + ["-c", os.path.join(CSRC, f), "-o", os.path.join(OBJ, f + ".o")])
    for f in asm:
        jobs.append(["gcc", "-c", "-DVQVDB_PACK_PATH=\"%s\"" % PACK, os.path.join(CSRC, f), "-o",
                     os.path.join(OBJ, f + ".o")])
    with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
        for fut in [ex.submit(_run, j, verbose, log) for j in jobs]:
            fut.result()
    objs = [os.path.join(OBJ, f + ".o") for f in cu + cpp + asm]
    _run([nvcc, "-shared", "-o", LIB] + ARCH + objs + ["-Xlinker", "--no-undefined"], verbose, log)
    with open(os.path.join(REPO, "build", "build.log"), "w") as fh:
        fh.write("\n".join(log))
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
    sys.exit(0)
