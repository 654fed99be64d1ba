"""The tensor-core encoder scores the codebook straight from the attention output: proj (1x1 conv) is folded into the
codebook on the host.  Host-only check of that fold (csrc/encode_tc_host.cpp, through vqvdb_b200_debug_fold_encoder_vq)
against the reference's two steps evaluated directly in float64: z = proj(x) (python/VQVAE_v2.py:250) and
dist_k = |z|^2 + |e_k|^2 - 2 z.e_k (python/save_for_inference.py:55-61)."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tools"))


def test_folded_scores_equal_projection_then_distance():
    from vqvdb_b200 import build
    from vqvdb_b200.codec import fold_encoder_vq
    from weights_pack import read_pack
    build.build()
    _, T = read_pack(os.path.join(REPO, "vqvdb_b200", "weights", "vqvae_float.vqw"))
    W = np.asarray(T["encoder.proj.weight"], dtype=np.float64).reshape(128, 32)
    b = np.asarray(T["encoder.proj.bias"], dtype=np.float64)
    E = np.asarray(T["quantizer.embedding"], dtype=np.float64)          # [256][128]
    m, esq, norm = fold_encoder_vq()                                       # the embedded pack == vqvae_float.vqw
    rng = np.random.default_rng(0)
    x = rng.standard_normal((512, 32)) * rng.uniform(0.01, 10.0, size=(512, 1))
    z = x @ W.T + b
    dist = (z * z).sum(1, keepdims=True) + (E * E).sum(1)[None, :] - 2.0 * z @ E.T     # the reference's formula
    folded = esq.astype(np.float64)[None, :] - 2.0 * x @ m.astype(np.float64).T         # = dist - |z|^2
    want = dist - (z * z).sum(1, keepdims=True)
    scale = np.abs(want).max()
    assert np.abs(folded - want).max() <= 1e-6 * scale                  # fp32 storage of M and esq
    assert (folded.argmin(1) == dist.argmin(1)).mean() >= 0.999          # same arg-min up to fp32-storage near-ties
    # the bound's ingredients never under-estimate
    true_norm = np.sqrt(((E @ W) ** 2).sum(1))
    assert (norm[:256] >= true_norm * (1 - 1e-7)).all() and norm[256] == norm[:256].max()


def test_vec3_folded_scores_equal_projection_then_distance():
    # the same fold for the vec3 model's 128 -> 128 proj (csrc/encode_tc128_host.cpp), and the three constants of its
    # error bound: max |M_k|, an upper bound of |proj.weight|_2, |proj.bias| + max |e_k|
    from vqvdb_b200 import build
    from vqvdb_b200.codec import fold_encoder_vq
    from weights_pack import read_pack
    build.build()
    pack = os.path.join(REPO, "vqvdb_b200", "weights", "vqvae_vec3_seed0.vqw")
    _, T = read_pack(pack)
    W = np.asarray(T["encoder.proj.weight"], dtype=np.float64).reshape(128, 128)
    b = np.asarray(T["encoder.proj.bias"], dtype=np.float64)
    E = np.asarray(T["quantizer.embedding"], dtype=np.float64)          # [256][128]
    m, esq, norm = fold_encoder_vq(pack, channels=128)
    rng = np.random.default_rng(1)
    x = rng.standard_normal((512, 128)) * rng.uniform(0.01, 10.0, size=(512, 1))
    z = x @ W.T + b
    dist = (z * z).sum(1, keepdims=True) + (E * E).sum(1)[None, :] - 2.0 * z @ E.T
    folded = esq.astype(np.float64)[None, :] - 2.0 * x @ m.astype(np.float64).T
    want = dist - (z * z).sum(1, keepdims=True)
    assert np.abs(folded - want).max() <= 1e-6 * np.abs(want).max()
    assert (folded.argmin(1) == dist.argmin(1)).mean() >= 0.999
    assert norm[0] >= np.sqrt(((E @ W) ** 2).sum(1)).max() * (1 - 1e-7)
    assert norm[1] >= np.linalg.norm(W, 2) * (1 - 1e-7)
    assert norm[2] >= (np.linalg.norm(b) + np.sqrt((E * E).sum(1)).max()) * (1 - 1e-7)
    assert (norm[3:] == 0).all()
