"""vqvdb_b200/cpp/sop/SOP_VQVDB_B200.cpp — the Houdini SOP shim (reference: src/SOP/*.cpp) — syntax-checked against
declaration-only stand-ins for the HDK classes it uses (tests/stubs/hdk/) and the OpenVDB stand-in
(tests/stubs/openvdb/).  No HDK exists in this image: this pins that the shim parses and type-checks against the
call shapes it was written for, nothing more."""
import os
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
SHIM = os.path.join(REPO, "vqvdb_b200", "cpp", "sop", "SOP_VQVDB_B200.cpp")


def test_sop_shim_type_checks_against_hdk_stubs():
    r = subprocess.run([CXX, "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-DVQVDB_B200_WITH_HDK=1",
                        "-I", os.path.join(REPO, "tests", "stubs", "hdk"), "-I", os.path.join(REPO, "tests", "stubs"),
                        "-I", os.path.join(REPO, "include"), SHIM], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout


def test_sop_shim_is_inert_without_the_hdk():
    # no HDK headers on the include path: the translation unit must compile to nothing rather than fail
    r = subprocess.run([CXX, "-std=c++17", "-fsyntax-only", "-I", os.path.join(REPO, "include"), SHIM],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout


def test_sop_shim_keeps_the_reference_surface():
    src = open(SHIM).read()
    for token in ('"vqvdb_encoder"', '"vqvdb_decoder"', '"vdbname"', '"outputpath"', '"inputfile"', '"batchsize"', '"execute"',
                  "BackendType::B200", "newSopOperator", "wasInterrupted"):
        assert token in src, token
