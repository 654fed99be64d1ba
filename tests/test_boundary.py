"""The drop-in boundary, checked against the reference's OWN header.

A maintainer copies vqvdb_b200/cpp/B200Backend.{hpp,cpp} into the reference tree and adds ONE enumerator to
`enum class BackendType` (src/core/IVQVAECodec.hpp:21).  This test performs exactly that edit on a scratch copy of the
reference header and compiles the backend against it — nothing else of the reference is touched, in particular
CodecConfig stays {device, source} (IVQVAECodec.hpp:85-89).  Skipped where /root/reference does not exist (GPU box).
"""
import os
import re
import shutil
import subprocess

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("VQVDB_REFERENCE_ROOT", "/root/reference")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _patched_reference_header(dst_dir):
    src = open(os.path.join(REF, "src", "core", "IVQVAECodec.hpp")).read()
    patched, n = re.subn(r"enum class BackendType \{ LibTorch, ONNX \};", "enum class BackendType { LibTorch, ONNX, B200 };", src)
    assert n == 1, "the reference's BackendType enum is not where INTEGRATION.md says it is"
    os.makedirs(os.path.join(dst_dir, "core"), exist_ok=True)
    open(os.path.join(dst_dir, "core", "IVQVAECodec.hpp"), "w").write(patched)
    open(os.path.join(dst_dir, "IVQVAECodec.hpp"), "w").write('#include "core/IVQVAECodec.hpp"\n')
    return src


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference tree not present")
def test_backend_compiles_against_the_reference_header(tmp_path):
    src = _patched_reference_header(str(tmp_path))
    # the reference's CodecConfig has exactly two fields; the backend must not need more
    cfg = re.search(r"struct CodecConfig \{(.*?)\n\};", src, re.S).group(1)
    assert "cudaDevice" not in cfg and re.findall(r"^\s*(?:Device|ModelSource)\s+(\w+)", cfg, re.M) == ["device", "source"]
    for f in ("B200Backend.hpp", "B200Backend.cpp"):
        shutil.copy(os.path.join(REPO, "vqvdb_b200", "cpp", f), str(tmp_path / f))
    r = subprocess.run([CXX, "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-I", str(tmp_path),
                        "-I", os.path.join(REPO, "include"), str(tmp_path / "B200Backend.cpp")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference tree not present")
def test_factory_case_compiles_inside_the_reference_factory(tmp_path):
    """the `case BackendType::B200` edit, applied to a scratch copy of the reference's factory
    (src/core/IVQVAECodec.cpp:76-110) with its own backends disabled, compiles and links against libvqvdb_b200.so."""
    from vqvdb_b200 import build
    lib = build.build()
    _patched_reference_header(str(tmp_path))
    for f in ("B200Backend.hpp", "B200Backend.cpp"):
        shutil.copy(os.path.join(REPO, "vqvdb_b200", "cpp", f), str(tmp_path / f))
    fac = open(os.path.join(REF, "src", "core", "IVQVAECodec.cpp")).read()
    fac, n1 = re.subn(r'(#include "IVQVAECodec.hpp"\n)', r'\1#ifdef ENABLE_B200_BACKEND\n#include "B200Backend.hpp"\n#endif\n', fac, count=1)
    fac, n2 = re.subn(r"(\n\t\t\tdefault:)", "\n#ifdef ENABLE_B200_BACKEND\n\t\t\tcase BackendType::B200:\n\t\t\t\treturn std::make_unique<B200Backend>(config);\n#endif" + r"\1", fac, count=1)
    assert n1 == 1 and n2 == 1, "the reference factory no longer has the include / switch INTEGRATION.md §2 edits"
    open(str(tmp_path / "core" / "IVQVAECodec.cpp"), "w").write(fac)
    main = str(tmp_path / "main.cpp")
    open(main, "w").write('#include "core/IVQVAECodec.hpp"\n#include <cstdio>\n'
                          "int main() { CodecConfig c; c.device = CodecConfig::Device::CUDA;\n"
                          " auto p = IVQVAECodec::create(c, BackendType::B200); std::puts(p ? \"created\" : \"null\"); return 0; }\n")
    exe = str(tmp_path / "factory_probe")
    r = subprocess.run([CXX, "-std=c++17", "-DENABLE_B200_BACKEND", "-I", str(tmp_path), "-I", str(tmp_path / "core"),
                        "-I", os.path.join(REPO, "include"), main, str(tmp_path / "core" / "IVQVAECodec.cpp"),
                        str(tmp_path / "B200Backend.cpp"), "-L", os.path.dirname(lib), "-lvqvdb_b200",
                        "-Wl,-rpath," + os.path.dirname(lib), "-pthread", "-o", exe],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    out = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    import torch
    # the reference's contract: create() never throws; without a device it logs and returns null (no CPU fallback)
    assert out.returncode == 0
    assert ("created" in out.stdout) if torch.cuda.is_available() else ("null" in out.stdout and "no CPU fallback" in out.stdout)


def test_this_repositorys_header_adds_only_the_enumerator():
    ours = open(os.path.join(REPO, "vqvdb_b200", "cpp", "IVQVAECodec.hpp")).read()
    cfg = re.search(r"struct CodecConfig \{(.*?)\n\};", ours, re.S).group(1)
    fields = re.findall(r"^\s*(?:Device|ModelSource|int|uint32_t|bool)\s+(\w+)", cfg, re.M)
    assert fields == ["device", "source"], fields
    assert "enum class BackendType { LibTorch, ONNX, B200 };" in ours
