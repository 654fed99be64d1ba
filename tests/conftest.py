import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

# Index parity rule: uint8 indices must equal the reference's, except where the
# reference's own fp32 top-2 distance margin is below this bound (a near-tie that any change of
# summation order can flip; SURVEY §7.4: the reference's fp32 formula itself disagrees with fp64 there).
TIE_MARGIN = 1e-4
# Reconstruction tolerance for fp32 paths, absolute, on sigmoid outputs in (0,1).
RECON_ATOL_FP32 = 2e-5


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden(name):
    import numpy as np
    return np.load(os.path.join(REPO, "tests", "golden", name + ".npz"))


def assert_indices_match(got, gold_idx, gold_margin, max_frac=2e-3):
    """Bit-exact except at reference near-ties; returns the mismatch count."""
    import numpy as np
    mm = got != gold_idx
    n_mm = int(mm.sum())
    if n_mm:
        worst = float(gold_margin[mm].max())
        assert worst <= TIE_MARGIN, "index mismatch at a latent whose reference margin is %.3e" % worst
        if got.size >= 4096:  # tiny sets of identical leaves can flip the same near-tie several times
            assert n_mm <= int(max_frac * got.size), "%d of %d indices differ" % (n_mm, got.size)
    return n_mm


@pytest.fixture(scope="session")
def c_oracle():
    from oracle.pyoracle import COracle
    return COracle()
