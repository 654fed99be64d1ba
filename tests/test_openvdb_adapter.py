"""vqvdb_b200/cpp/openvdb_adapter.hpp — the FloatGrid <-> LeafGrid conversions and the grid-level compress / decompress
(reference: src/orchestrator/VQVAECodec.cpp:26-65, 84-101, 150-200) — compiled and RUN against the functional OpenVDB
stand-in of tests/stubs/openvdb/ (OpenVDB is not installed in this image), through the repository's orchestrator and
.vqvdb container with a deterministic CPU stand-in for the backend.  See tests/cpp/adapter_stub_test.cpp for what is pinned."""
import os
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def test_adapter_compiles_and_roundtrips_against_the_openvdb_stub(tmp_path):
    from vqvdb_b200 import build
    lib = build.build()
    cpp = os.path.join(REPO, "vqvdb_b200", "cpp")
    exe = str(tmp_path / "adapter_stub_test")
    r = subprocess.run([CXX, "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(REPO, "tests", "stubs"), "-I", cpp,
                        "-I", os.path.join(REPO, "include"), os.path.join(REPO, "tests", "cpp", "adapter_stub_test.cpp"),
                        os.path.join(cpp, "VQVAECodec.cpp"), os.path.join(cpp, "vqvdb_file.cpp"), os.path.join(cpp, "B200Backend.cpp"),
                        "-L", os.path.dirname(lib), "-lvqvdb_b200", "-Wl,-rpath," + os.path.dirname(lib), "-pthread", "-o", exe],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    out = subprocess.run([exe, str(tmp_path / "t.vqvdb")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert out.returncode == 0 and "adapter_stub_test: ok" in out.stdout, out.stdout


def test_adapter_parallel_variant_is_syntax_clean(tmp_path):
    # VQVDB_B200_ADAPTER_PARALLEL needs TBB; with a two-function stand-in for tbb::parallel_for it must at least parse
    os.makedirs(tmp_path / "tbb")
    (tmp_path / "tbb" / "blocked_range.h").write_text(
        "#pragma once\n#include <cstddef>\nnamespace tbb { template <class T> struct blocked_range { T b, e; std::size_t g;"
        " blocked_range(T b_, T e_, std::size_t g_ = 1) : b(b_), e(e_), g(g_) {} T begin() const { return b; } T end() const { return e; } }; }\n")
    (tmp_path / "tbb" / "parallel_for.h").write_text(
        "#pragma once\nnamespace tbb { template <class R, class F> void parallel_for(const R& r, const F& f) { f(r); } }\n")
    src = tmp_path / "p.cpp"
    src.write_text('#define VQVDB_B200_WITH_OPENVDB 1\n#define VQVDB_B200_ADAPTER_PARALLEL 1\n#include "openvdb_adapter.hpp"\nint main() { return 0; }\n')
    r = subprocess.run([CXX, "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-I", str(tmp_path), "-I", os.path.join(REPO, "tests", "stubs"),
                        "-I", os.path.join(REPO, "vqvdb_b200", "cpp"), "-I", os.path.join(REPO, "include"), str(src)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
