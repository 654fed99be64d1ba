"""The oracle against the reference's own outputs (tests/golden, made by tools/make_goldens.py)."""
import hashlib

import numpy as np
import pytest

from conftest import RECON_ATOL_FP32, assert_indices_match, golden
from vqvdb_b200 import synth

CASES = {
    "kat256": lambda: synth.kat_leaves(256),
    "smoke1024_seed0": lambda: synth.smoke_leaves(1024, seed=0),
    "sparse1024_seed1": lambda: synth.smoke_leaves(1024, seed=1, sparse=True),
    "noise256_seed2": lambda: synth.noise_leaves(256, seed=2),
    "fogsphere64": lambda: synth.fog_sphere_grid()[1],
    "zeros4": lambda: np.zeros((4, 1, 8, 8, 8), np.float32),
}


@pytest.mark.parametrize("name", list(CASES))
def test_generators_reproduce_golden_inputs(name):
    x = CASES[name]()
    assert hashlib.sha256(x.tobytes()).hexdigest() == str(golden(name)["input_sha256"])


def test_known_answer_vector_of_survey():
    # values agreed by the TorchScript oracle and the C++ TorchBackend.
    g = golden("kat256")
    idx = g["indices"]
    assert idx.ravel()[:4].tolist() == [211, 164, 220, 99]
    assert int(idx.sum()) == 1894112
    assert hashlib.sha256(idx.tobytes()).hexdigest() == \
        "2d8b7f4f9c0866de2f4a0313768811ca4ddc31162a7de88ee461d2ce46b21bc3"
    assert abs(float(g["recon"].ravel()[0]) - 0.316964) < 1e-6
    assert abs(float(g["recon_sum"]) - 56535.984326) < 1e-3


@pytest.mark.parametrize("name", list(CASES))
def test_c_oracle_encode_matches_reference(c_oracle, name):
    g = golden(name)
    idx, margins = c_oracle.encode(CASES[name](), with_margins=True)
    assert_indices_match(idx, g["indices"], g["margins"])
    # the oracle's own margins agree with the reference's away from ties
    assert np.allclose(margins, g["margins"], atol=1e-3, rtol=1e-3)


@pytest.mark.parametrize("name", ["kat256", "sparse1024_seed1", "fogsphere64", "zeros4"])
def test_c_oracle_decode_matches_reference(c_oracle, name):
    g = golden(name)
    m = min(64, len(g["recon"]))
    rec = c_oracle.decode(g["indices"][:m])
    assert np.abs(rec - g["recon"][:m]).max() <= RECON_ATOL_FP32


def test_c_oracle_decode_random_indices(c_oracle):
    g = golden("decode_random128_seed1234")
    idx = synth.random_indices(128, seed=1234)
    assert hashlib.sha256(idx.tobytes()).hexdigest() == str(g["input_sha256"])
    rec = c_oracle.decode(idx[:48])
    assert np.abs(rec - g["recon"][:48]).max() <= RECON_ATOL_FP32


def test_zero_leaf_decodes_near_zero():
    # a zero leaf decodes to <= 1.3e-4
    assert float(golden("zeros4")["recon"].max()) <= 1.3e-4


def test_fog_sphere_leaf_count():
    og, leaves = synth.fog_sphere_grid()
    assert leaves.shape == (302, 1, 8, 8, 8) and og.shape == (302, 3)


def test_reference_backend_matches_golden_if_built():
    from oracle.pyoracle import RefCodec, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref not built")
    ref = RefCodec("cpu")
    try:
        assert ref.latent_shape() == [4, 4, 4]          # TorchBackend.cpp:97-119 probe
        g = golden("kat256")
        idx = ref.encode(synth.kat_leaves(256))
        assert np.array_equal(idx, g["indices"])
        rec = ref.decode(idx[:32])
        assert np.abs(rec - g["recon"][:32]).max() <= 1e-6
    finally:
        ref.close()


def test_c_oracle_vec3_matches_the_reference_classes():
    # config 4: the C restatement with the seeded vec3 pack against goldens made by the reference's own EncoderVec3 /
    # DecoderVec3 classes (tools/make_goldens.py); this is the checker of the vec3 GPU path and of bench.py --workload vec3
    import os
    import subprocess
    import sys
    from conftest import REPO
    from oracle.pyoracle import COracle
    pack = os.path.join(REPO, "vqvdb_b200", "weights", "vqvae_vec3_seed0.vqw")
    if not os.path.exists(pack):
        subprocess.check_call([sys.executable, os.path.join(REPO, "tools", "weights_pack.py"), "vec3"], stdout=subprocess.DEVNULL)
    o = COracle(pack, threads=os.cpu_count() or 1)
    g = golden("vec3_noise64_seed6")
    x = synth.noise_leaves(64, seed=6, channels=3)
    assert hashlib.sha256(x.tobytes()).hexdigest() == str(g["input_sha256"])
    idx, margins = o.encode(x, with_margins=True)
    assert_indices_match(idx, g["indices"], g["margins"])
    m = len(g["recon"])
    assert np.abs(o.decode(g["indices"][:m]) - g["recon"]).max() <= 5e-5
    big = golden("vec3_sparse1024_seed7")
    xb = synth.smoke_leaves(1024, seed=7, channels=3, sparse=True)
    assert hashlib.sha256(xb.tobytes()).hexdigest() == str(big["input_sha256"])
    idx_b, _ = o.encode(xb[:96], with_margins=True)
    assert_indices_match(idx_b, big["indices"][:96], big["margins"][:96])


@pytest.mark.parametrize("name,seed,ch", [("nonfinite8_seed11", 11, 1), ("vec3_nonfinite8_seed12", 12, 3)])
def test_c_oracle_non_finite_leaves_match_reference(name, seed, ch):
    # A NaN / +inf / -inf voxel poisons its leaf: the reference (TorchScript blob, the EncoderVec3 class) answers code 0 for
    # all of its 64 latents; the other leaves are untouched.
    import os
    from conftest import REPO
    from oracle.pyoracle import COracle, VEC3_PACK
    g = golden(name)
    x = synth.nonfinite_leaves(8, seed=seed, channels=ch)
    assert hashlib.sha256(x.tobytes()).hexdigest() == str(g["input_sha256"])
    poisoned = np.isnan(g["margins"]).reshape(8, -1).all(axis=1)
    assert poisoned.tolist() == [False, True, True, False, False, True, False, False]
    assert not g["indices"][poisoned].any()
    o = COracle(VEC3_PACK) if ch == 3 else COracle()
    idx = o.encode(x)
    assert not idx[poisoned].any()
    assert_indices_match(idx[~poisoned], g["indices"][~poisoned], g["margins"][~poisoned])
