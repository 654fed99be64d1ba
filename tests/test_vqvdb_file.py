""".vqvdb v3 container: the C++ writer/reader against (1) the reference's own VDBStreamWriter / VDBStreamReader
(src/Utils/VQVDB_Reader.cpp:20-162,168-335, compiled unmodified into oracle/_ref/libvqvdb_fmt.so: the format oracle)
and (2) an independent Python statement of the byte layout (VQVDB_Reader.hpp:30-43)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from vqvdb_b200 import hostlib
from vqvdb_b200.hostlib import LeafGrid

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _built():
    from vqvdb_b200 import build
    build.build()
    subprocess.check_call(["make", "-s", "-C", os.path.join(REPO, "vqvdb_b200", "cpp")])


def py_write(path, grids, num_embeddings=256):
    """Appendix B, written out by hand."""
    with open(path, "wb") as f:
        f.write(b"VQVDB" + struct.pack("<BBIB", 3, len(grids), num_embeddings, 3))
        for g in grids:
            nm = g.name.encode()
            f.write(struct.pack("<I", len(nm)) + nm)
            f.write(np.asarray(g.transform, "<f4").reshape(16).tobytes())
            f.write(struct.pack("<3H", 4, 4, 4))
            f.write(struct.pack("<I", len(g.origins)))
            rec = np.zeros((len(g.origins), 76), np.uint8)
            rec[:, :12] = np.ascontiguousarray(g.origins, "<i4").view(np.uint8).reshape(-1, 12)
            rec[:, 12:] = g.indices.reshape(len(g.origins), 64)
            f.write(rec.tobytes())


def py_read(path):
    b = open(path, "rb").read()
    assert b[:5] == b"VQVDB"
    ver, ngrids, k, nd = struct.unpack_from("<BBIB", b, 5)
    assert ver == 3 and nd == 3
    pos, out = 12, []
    for _ in range(ngrids):
        (ln,) = struct.unpack_from("<I", b, pos); pos += 4
        name = b[pos:pos + ln].decode(); pos += ln
        tr = np.frombuffer(b, "<f4", 16, pos).reshape(4, 4); pos += 64
        lat = struct.unpack_from("<3H", b, pos); pos += 6
        (n,) = struct.unpack_from("<I", b, pos); pos += 4
        rec = np.frombuffer(b, np.uint8, n * 76, pos).reshape(n, 76); pos += n * 76
        out.append(LeafGrid(name, rec[:, :12].copy().view("<i4").reshape(n, 3), None, rec[:, 12:].reshape(n, 4, 4, 4), tr))
    assert pos == len(b)
    return k, out


def make_grids(rng, sizes):
    gs = []
    for i, n in enumerate(sizes):
        tr = np.eye(4, dtype=np.float32) * (0.1 * (i + 1)); tr[3, 3] = 1; tr[3, :3] = rng.normal(size=3)
        gs.append(LeafGrid("density_%d" % i, rng.integers(-4096, 4096, size=(n, 3), dtype=np.int32) * 8, None,
                           rng.integers(0, 256, size=(n, 4, 4, 4), dtype=np.uint8), tr))
    return gs


def same(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert x.name == y.name
        assert np.array_equal(x.origins, y.origins)
        assert np.array_equal(x.indices.reshape(len(x.origins), -1), y.indices.reshape(len(y.origins), -1))
        assert np.array_equal(np.asarray(x.transform, np.float32).reshape(16), np.asarray(y.transform, np.float32).reshape(16))


def test_host_library_exports_every_declared_symbol():
    import re
    L = hostlib.load_host_library()
    hdr = open(os.path.join(REPO, "include", "vqvdb_b200_host.h")).read()
    declared = set(re.findall(r"VQVDB_HOST_API\s+[\w\s\*]+?\b(vqvdb_host_\w+)\s*\(", hdr))
    assert declared == set(hostlib.HOST_EXPORTS)
    for name in declared:
        assert hasattr(L, name)


def test_cpp_writer_matches_byte_layout(tmp_path):
    rng = np.random.default_rng(0)
    grids = make_grids(rng, [300, 1, 77])
    p_cpp, p_py = str(tmp_path / "a.vqvdb"), str(tmp_path / "b.vqvdb")
    hostlib.write_file(p_cpp, grids)
    py_write(p_py, grids)
    assert open(p_cpp, "rb").read() == open(p_py, "rb").read()
    k, back = py_read(p_cpp)
    assert k == 256
    same(back, grids)
    # on-disk size: 12 + per grid (4 + len(name) + 64 + 6 + 4 + 76 n)   
    assert os.path.getsize(p_cpp) == 12 + sum(4 + len(g.name) + 64 + 6 + 4 + 76 * len(g.origins) for g in grids)


def test_cpp_reader_reads_python_written_file(tmp_path):
    rng = np.random.default_rng(1)
    grids = make_grids(rng, [1000, 5])
    p = str(tmp_path / "c.vqvdb")
    py_write(p, grids)
    same(hostlib.read_file(p), grids)
    same(hostlib.read_file(p, batch=64), grids)       # ragged batches: 1000 = 15*64 + 40


def test_multi_grid_file_larger_than_any_read_buffer(tmp_path):
    # The reference reader double-decrements its byte budget (VQVDB_Reader.cpp:297,322) and mis-reads multi-grid
    # files whose first grid exceeds its 64 MiB buffer; 900k leaves = 68 MB of records.  This reader must not.
    rng = np.random.default_rng(2)
    n0 = 900_000
    g0 = LeafGrid("big", np.arange(n0 * 3, dtype=np.int32).reshape(n0, 3), None,
                  rng.integers(0, 256, size=(n0, 64), dtype=np.uint8))
    g1 = make_grids(rng, [10])[0]
    p = str(tmp_path / "big.vqvdb")
    hostlib.write_file(p, [g0, g1])
    back = hostlib.read_file(p, batch=1 << 18)
    same(back, [g0, g1])


# ---------------------------------------------------------------------------------------------
# Format oracle: the reference's own reader and writer.  Built by `make -C oracle fmt` where /root/reference exists
# (the dev container); the prebuilt library travels with the repo snapshot, so these also run on the GPU box.
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ref_fmt():
    from oracle import pyoracle
    if not pyoracle.fmt_available() and os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(REPO, "oracle"), "fmt"])
    if not pyoracle.fmt_available():
        pytest.skip("oracle/_ref/libvqvdb_fmt.so not built (needs the reference tree)")
    return pyoracle.RefFormat()


def _tuples(grids):
    return [(g.name, g.origins, g.indices.reshape(len(g.origins), -1), np.asarray(g.transform, np.float32).reshape(16)) for g in grids]


def _same_as_ref(ref_grids, grids):
    assert len(ref_grids) == len(grids)
    for r, g in zip(ref_grids, grids):
        assert r["name"] == g.name and r["total_blocks"] == len(g.origins)
        assert r["latent_shape"] == [4, 4, 4] and r["num_embeddings"] == 256
        assert np.array_equal(r["origins"], g.origins)
        assert np.array_equal(r["indices"], g.indices.reshape(len(g.origins), -1))
        assert np.array_equal(r["transform"], np.asarray(g.transform, np.float32).reshape(16))


@pytest.mark.parametrize("sizes", [[302], [1], [1000, 5, 77], [64, 64, 1]])
def test_reference_reader_reads_files_written_here(tmp_path, ref_fmt, sizes):
    rng = np.random.default_rng(10)
    grids = make_grids(rng, sizes)
    p = str(tmp_path / "ours.vqvdb")
    hostlib.write_file(p, grids)
    for batch in (64, 8192):                            # the decoder SOP's default and maximum batch
        _same_as_ref(ref_fmt.read(p, batch=batch), grids)


@pytest.mark.parametrize("sizes", [[302], [1000, 5, 77]])
def test_reader_here_reads_reference_written_files_and_bytes_agree(tmp_path, ref_fmt, sizes):
    rng = np.random.default_rng(11)
    grids = make_grids(rng, sizes)
    p_ref, p_cpp = str(tmp_path / "ref.vqvdb"), str(tmp_path / "ours.vqvdb")
    ref_fmt.write(p_ref, _tuples(grids), batch=64)       # VDBStreamWriter, fed 64 leaves at a time like compress()
    hostlib.write_file(p_cpp, grids)
    assert open(p_ref, "rb").read() == open(p_cpp, "rb").read()
    same(hostlib.read_file(p_ref), grids)
    same(hostlib.read_file(p_ref, batch=100), grids)


def test_reference_writer_empty_file_header(tmp_path, ref_fmt):
    # no grids: the reference leaves its placeholder header (numGrids = 0, K = 0, latent rank 0: VQVDB_Reader.cpp:58-61)
    p_ref, p_cpp = str(tmp_path / "ref.vqvdb"), str(tmp_path / "ours.vqvdb")
    ref_fmt.write(p_ref, [])
    hostlib.write_file(p_cpp, [])
    assert open(p_ref, "rb").read() == open(p_cpp, "rb").read()
    assert hostlib.read_file(p_ref) == [] and ref_fmt.read(p_cpp) == []


def test_large_first_grid_divergence_is_the_documented_one(tmp_path, ref_fmt):
    # VDBStreamReader decrements remainingDataBytes_ twice (VQVDB_Reader.cpp:297 and :322).  With a
    # first grid larger than its 64 MiB buffer the budget wraps around, the second refill swallows the rest of the
    # file — the next grid's metadata included — into grid 0's record buffer, and the reader then fails on grid 1.
    # The file itself is fine: the reference WRITER produces the same bytes as ours, and this reader reads both grids.
    rng = np.random.default_rng(12)
    n0 = 900_000                                         # 68.4 MB of records > 64 MiB
    g0 = LeafGrid("big", np.arange(n0 * 3, dtype=np.int32).reshape(n0, 3), None, rng.integers(0, 256, size=(n0, 64), dtype=np.uint8))
    g1 = make_grids(rng, [10])[0]
    p_ref, p_cpp = str(tmp_path / "ref.vqvdb"), str(tmp_path / "ours.vqvdb")
    ref_fmt.write(p_ref, _tuples([g0, g1]), batch=8192)
    hostlib.write_file(p_cpp, [g0, g1])
    import hashlib
    assert hashlib.sha256(open(p_ref, "rb").read()).digest() == hashlib.sha256(open(p_cpp, "rb").read()).digest()
    same(hostlib.read_file(p_ref, batch=1 << 18), [g0, g1])
    try:
        got = ref_fmt.read(p_ref, batch=8192)
    except RuntimeError as e:
        assert "grid name length" in str(e) or "truncated" in str(e).lower(), str(e)
    else:                                                # did not throw: then it must have mis-read the second grid
        assert len(got) != 2 or not np.array_equal(got[1]["indices"], g1.indices.reshape(10, -1))
    # the single-grid form of the same size reads fine through the reference reader (wrap-around is harmless there)
    hostlib.write_file(p_cpp, [g0])
    _same_as_ref(ref_fmt.read(p_cpp, batch=8192), [g0])


def test_empty_file_and_empty_grid(tmp_path):
    p = str(tmp_path / "e.vqvdb")
    hostlib.write_file(p, [])
    assert os.path.getsize(p) == 12 and hostlib.read_file(p) == []
    g = LeafGrid("none", np.zeros((0, 3), np.int32), None, np.zeros((0, 64), np.uint8))
    hostlib.write_file(p, [g])
    back = hostlib.read_file(p)
    assert len(back) == 1 and len(back[0].origins) == 0 and back[0].name == "none"


def test_reader_rejects_bad_files(tmp_path):
    p = str(tmp_path / "x.vqvdb")
    open(p, "wb").write(b"NOTVQ" + b"\0" * 7)
    with pytest.raises(RuntimeError, match="bad magic"):
        hostlib.read_file(p)
    open(p, "wb").write(b"VQVDB" + struct.pack("<BBIB", 2, 0, 256, 3))
    with pytest.raises(RuntimeError, match="version"):
        hostlib.read_file(p)
    rng = np.random.default_rng(3)
    py_write(p, make_grids(rng, [50]))
    blob = open(p, "rb").read()
    open(p, "wb").write(blob[:-100])                    # truncated records
    with pytest.raises(RuntimeError, match="Unexpected end of file"):
        hostlib.read_file(p)
    with pytest.raises(RuntimeError, match="Cannot open"):
        hostlib.read_file(str(tmp_path / "missing.vqvdb"))


@pytest.mark.gpu
def test_compress_decompress_fog_sphere_through_host_layer(tmp_path):
    """Config 1: the 64^3 fog sphere, 302 leaves, through compress() -> .vqvdb -> decompress()."""
    from conftest import assert_indices_match, golden
    from vqvdb_b200 import synth
    og, leaves = synth.fog_sphere_grid()
    g = golden("fogsphere64")
    p = str(tmp_path / "fog.vqvdb")
    tr = np.eye(4, dtype=np.float32) * 0.05; tr[3, 3] = 1
    hostlib.compress([LeafGrid("density", og, leaves, None, tr)], p, batch_size=64)   # the SOP default batch
    assert os.path.getsize(p) == 12 + 4 + 7 + 64 + 6 + 4 + 76 * 302
    stored = hostlib.read_file(p)[0]
    assert stored.name == "density" and np.array_equal(stored.origins, og)
    assert np.array_equal(np.asarray(stored.transform), tr)
    assert_indices_match(stored.indices, g["indices"], g["margins"])
    out = hostlib.decompress(p, batch_size=0, fp32_decode=True)[0]
    assert np.array_equal(out.origins, og)
    same_idx = (stored.indices == g["indices"]).reshape(302, -1).all(axis=1)[:len(g["recon"])]
    assert np.abs(out.voxels[:len(g["recon"])][same_idx] - g["recon"][same_idx]).max() <= 2e-5
    out_tc = hostlib.decompress(p, batch_size=100)[0]
    assert synth.psnr(out_tc.voxels, out.voxels) >= 55.0


@pytest.mark.gpu
def test_orchestrator_refuses_a_three_channel_model():
    """compress / decompress are FloatGrid-only (512 floats per leaf, like the reference: VQVAECodec.hpp:40,49); a vec3
    weight pack behind the backend must be refused when the orchestrator is built, not overflow a buffer later."""
    import sys
    L = hostlib.load_host_library()
    pack = os.path.join(REPO, "vqvdb_b200", "weights", "vqvae_vec3_seed0.vqw")
    if not os.path.exists(pack):
        subprocess.check_call([sys.executable, os.path.join(REPO, "tools", "weights_pack.py"), "vec3"], stdout=subprocess.DEVNULL)
    assert L.vqvdb_host_orchestrator_accepts(0, None) == 0
    assert L.vqvdb_host_orchestrator_accepts(0, os.path.join(REPO, "vqvdb_b200", "weights", "vqvae_float.vqw").encode()) == 0
    assert L.vqvdb_host_orchestrator_accepts(0, pack.encode()) == -1
    assert b"1-channel" in L.vqvdb_host_last_error()


@pytest.mark.gpu
def test_backend_virtuals_on_pageable_memory_match_the_pointer_api():
    """IVQVAECodec::encode/decode through the C++ backend (TensorView over pageable numpy memory -> owning Tensor whose
    buffer is filled without a zero-fill pass) against the C-ABI called directly; also wrong shapes are refused."""
    from vqvdb_b200 import BackendType, CodecConfig, IVQVAECodec, synth
    x = synth.smoke_leaves(5000, seed=8)                 # > 2048: the chunked pipeline with threaded staging
    hb = hostlib.HostBackend(0)
    try:
        idx, _ = hb.encode(x)
        idx = np.array(idx)
        rec, _ = hb.decode(idx)
        rec = np.array(rec)
        small_idx, _ = hb.encode(x[:64])                  # <= 2048: the zero-copy path on staged pageable memory
        assert np.array_equal(np.array(small_idx), idx[:64])
    finally:
        hb.close()
    c = IVQVAECodec.create(CodecConfig(device=CodecConfig.Device.CUDA), BackendType.B200)
    try:
        idx2 = np.empty((5000, 4, 4, 4), np.uint8)
        rec2 = np.empty((5000, 1, 8, 8, 8), np.float32)
        c.encode_into(x, 5000, idx2)
        c.decode_into(idx2, 5000, rec2)
    finally:
        c.close()
    assert np.array_equal(idx, idx2) and np.array_equal(rec, rec2)


@pytest.mark.gpu
def test_native_batched_loop_equals_whole_grid_calls():
    """The SOPs' calling pattern — one synchronous backend call per 64 leaves and direction — as the native loop of
    vqvdb_host_backend_roundtrip_batched (what bench.py's e2e_small_batches.native_loop times): same indices and voxels as
    one call over the whole grid, ragged tail included."""
    from vqvdb_b200 import synth
    x = synth.smoke_leaves(300, seed=9)                   # 4 full batches of 64 + 44
    hb = hostlib.HostBackend(0)
    try:
        idx1 = np.empty((300, 4, 4, 4), np.uint8)
        vox1 = np.empty((300, 1, 8, 8, 8), np.float32)
        hb.encode_into(x, idx1)
        hb.decode_into(idx1, vox1)
        idx2 = np.zeros_like(idx1)
        vox2 = np.zeros_like(vox1)
        assert hb.roundtrip_batched(x.ctypes.data, 300, 64, idx2.ctypes.data, vox2.ctypes.data) > 0
        assert np.array_equal(idx1, idx2) and np.array_equal(vox1, vox2)
    finally:
        hb.close()
