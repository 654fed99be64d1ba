"""The folded decoder tail (VQVDB_B200_DECODE_BF16_TC2_FOLD) is the same linear map as the reference's
up_conv -> PixelShuffle3D(2) -> final (python/VQVAE_v2.py:172-187, 266-275), zero padding at both resolutions included.

Host-only: the product's fold (csrc/decode_tc_host.cpp, through vqvdb_b200_debug_fold_decoder_tail) against the three
layers evaluated one after the other in float64 with the pack's own weights."""
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tools"))


def _tail_layer_by_layer(a, up_w, up_b, fin_w, fin_b):
    import torch
    import torch.nn.functional as F
    u = F.conv3d(a, up_w, up_b, padding=1)
    b, c, d, h, w = u.shape
    ps = u.view(b, c // 8, 2, 2, 2, d, h, w).permute(0, 1, 5, 2, 6, 3, 7, 4).contiguous().view(b, c // 8, 2 * d, 2 * h, 2 * w)
    return F.conv3d(ps, fin_w, fin_b, padding=1)[:, 0]


def _tail_folded(a, wg, bg, fin_b):
    import torch
    import torch.nn.functional as F
    G = F.conv3d(a, wg, bg, padding=1)                     # [b, 64 = r*8 + eps, 4, 4, 4]
    out = torch.zeros(a.shape[0], 8, 8, 8, dtype=a.dtype)
    for D in range(8):
        for H in range(8):
            for W in range(8):
                p, r = (D >> 1, H >> 1, W >> 1), (D & 1, H & 1, W & 1)
                s = fin_b[0].expand(a.shape[0]).clone()
                for eps in range(8):
                    e = ((eps >> 2) & 1, (eps >> 1) & 1, eps & 1)
                    q = [p[i] + ((1 if r[i] else -1) if e[i] else 0) for i in range(3)]
                    if all(0 <= x < 4 for x in q):
                        s = s + G[:, (r[0] * 4 + r[1] * 2 + r[2]) * 8 + eps, q[0], q[1], q[2]]
                out[:, D, H, W] = s
    return out


@pytest.mark.parametrize("seed", [0, 1])
def test_folded_tail_equals_the_three_layers(seed):
    import torch
    from vqvdb_b200 import build
    from vqvdb_b200.codec import fold_decoder_tail
    from weights_pack import read_pack
    build.build()
    _, T = read_pack(os.path.join(REPO, "vqvdb_b200", "weights", "vqvae_float.vqw"))
    f64 = lambda name: torch.tensor(np.asarray(T[name]), dtype=torch.float64)
    up_w, up_b = f64("decoder.up_conv.weight"), f64("decoder.up_conv.bias")
    fin_w, fin_b = f64("decoder.final.weight"), f64("decoder.final.bias")
    wg, bg = fold_decoder_tail()                            # the embedded pack == vqvae_float.vqw
    torch.manual_seed(seed)
    a = torch.randn(4, 64, 4, 4, 4, dtype=torch.float64)
    a[0] = 0.0                                              # bias-only case: the boundary terms of the bias must drop too
    want = _tail_layer_by_layer(a, up_w, up_b, fin_w, fin_b)
    got = _tail_folded(a, torch.tensor(wg, dtype=torch.float64), torch.tensor(bg, dtype=torch.float64), fin_b)
    scale = float(want.abs().max())
    # the fold is computed in double and stored as fp32: relative error ~1e-7
    assert float((got - want).abs().max()) <= 2e-6 * scale
