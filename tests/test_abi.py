"""CPU-side checks of the boundary: the library loads, exports every declared symbol, and fails
loudly (no fallback) when there is no device.  No compute calls here."""
import ctypes as C
import os
import re

import pytest

import vqvdb_b200
from vqvdb_b200 import BackendType, CodecConfig, IVQVAECodec

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from vqvdb_b200 import build
    build.build()
    return vqvdb_b200.load_library()


def test_header_and_binding_agree(lib):
    hdr = open(os.path.join(REPO, "include", "vqvdb_b200.h")).read()
    declared = set(re.findall(r"VQVDB_B200_API\s+[\w\s\*]+?\b(vqvdb_b200_\w+)\s*\(", hdr))
    assert declared == set(vqvdb_b200.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), "libvqvdb_b200.so does not export %s" % name


def test_version_string(lib):
    assert b"sm_100a" in lib.vqvdb_b200_version()


def test_config_struct_matches_header():
    from vqvdb_b200.codec import _Config
    # uint32 + int32 + ptr + u64 + ptr + 3*u32 + 7*u32 on LP64
    assert C.sizeof(_Config) == 4 + 4 + 8 + 8 + 8 + 4 + 4 + 32


def test_null_arguments_are_rejected(lib):
    assert lib.vqvdb_b200_create(None, None) == -1
    assert lib.vqvdb_b200_kernel_launches(None) == 0
    assert lib.vqvdb_b200_in_channels(None) == -1


def test_no_silent_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device path cannot be exercised")
    h = C.c_void_p()
    rc = lib.vqvdb_b200_create(None, C.byref(h))
    assert rc == -2 and not h                      # VQVDB_B200_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.vqvdb_b200_last_error(None)
    # factory contract of the reference: create() swallows the failure and returns null (IVQVAECodec.cpp:106-109)
    assert IVQVAECodec.create(CodecConfig(device=CodecConfig.Device.CUDA), BackendType.B200) is None
    # a CPU device request is refused outright rather than served by some host path
    assert IVQVAECodec.create(CodecConfig(device=CodecConfig.Device.CPU), BackendType.B200) is None
    assert IVQVAECodec.create(CodecConfig(device=CodecConfig.Device.CUDA), BackendType.LibTorch) is None


def test_product_never_imports_oracle():
    pkg = os.path.join(REPO, "vqvdb_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(root, f), errors="ignore").read()
                for line in src.splitlines():
                    if re.match(r"\s*(from|import)\s+oracle|\s*#include\s+[\"<].*oracle", line):
                        raise AssertionError("%s references oracle/: %s" % (f, line))
                assert "libvqvae_oracle" not in src and "libvqvdb_ref" not in src, f
