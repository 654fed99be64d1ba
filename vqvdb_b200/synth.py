"""Synthetic 8x8x8 leaf generators shared by the tests and bench.py.

Pure numpy: deterministic from the seed, identical here and on the GPU box.
Leaf buffer layout is OpenVDB's: offset = (x << 6) | (y << 3) | z, i.e. the
[8, 8, 8] array is indexed [x][y][z] with z fastest
(reference: orchestrator/VQVAECodec.cpp:54 copies leaf.buffer().data() verbatim).
"""
from __future__ import annotations

import numpy as np


def kat_leaves(n: int = 256) -> np.ndarray:
    """Known-answer input of SURVEY Appendix C: x[i] = ((i * 2654435761) mod 1000) / 1000."""
    i = np.arange(n * 512, dtype=np.uint64)
    v = (i * np.uint64(2654435761)) % np.uint64(1000)
    return (v.astype(np.float32) / np.float32(1000.0)).reshape(n, 1, 8, 8, 8)


def _trilinear_3_to_8(ctrl: np.ndarray) -> np.ndarray:
    """align_corners=True trilinear upsample of [..., 3, 3, 3] control grids to [..., 8, 8, 8]."""
    t = np.linspace(0.0, 2.0, 8, dtype=np.float64)
    i0 = np.minimum(np.floor(t).astype(np.int64), 1)
    f = (t - i0).astype(np.float32)
    w = np.zeros((8, 3), dtype=np.float32)
    w[np.arange(8), i0] = 1.0 - f
    w[np.arange(8), i0 + 1] += f
    out = np.einsum("...abc,xa,yb,zc->...xyz", ctrl.astype(np.float32), w, w, w, optimize=True)
    return out.astype(np.float32)


def smoke_leaves(n: int, seed: int = 0, channels: int = 1, sparse: bool = False,
                 chunk: int = 65536) -> np.ndarray:
    """Smooth 'smoke' leaves in [0,1] (float model) or [-1,1] (vec3 model).

    sparse=True applies clamp(2*(v-0.5), 0, 1), which produces many exactly-zero voxels
    and exercises all 256 codes.
    """
    rng = np.random.default_rng(seed)
    out = np.empty((n, channels, 8, 8, 8), dtype=np.float32)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        ctrl = rng.random((hi - lo, channels, 3, 3, 3), dtype=np.float32)
        v = np.clip(_trilinear_3_to_8(ctrl), 0.0, 1.0)
        if sparse:
            v = np.clip(2.0 * (v - 0.5), 0.0, 1.0)
        if channels != 1:
            v = 2.0 * v - 1.0
        out[lo:hi] = v
    return out


def noise_leaves(n: int, seed: int = 0, channels: int = 1) -> np.ndarray:
    rng = np.random.default_rng(seed)
    v = rng.random((n, channels, 8, 8, 8), dtype=np.float32)
    return v if channels == 1 else 2.0 * v - 1.0


def nonfinite_leaves(n: int = 8, seed: int = 11, channels: int = 1) -> np.ndarray:
    """Smoke leaves with a NaN voxel in leaf 1, +inf in leaf 2 and -inf in leaf n-3: the reference turns such a leaf into
    code 0 everywhere (torch.relu keeps the NaN, every distance is NaN, torch.argmin answers 0)."""
    x = smoke_leaves(n, seed=seed, channels=channels)
    x[1, 0, 2, 3, 4] = np.nan
    x[2, channels - 1, 7, 7, 7] = np.inf
    x[n - 3, 0, 0, 0, 0] = -np.inf
    return x


def random_indices(n: int, seed: int = 1234) -> np.ndarray:
    """uint8 i.i.d. uniform[0,255] indices for decode-only throughput."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(n, 4, 4, 4), dtype=np.uint8)


def fog_sphere_grid(dim: int = 64, radius: float = 28.0, width: float = 8.0):
    """Config 1: dense [dim]^3 fog sphere rho = clamp((R - |p - c|) / w, 0, 1) cut into 8^3 leaves.

    Returns (origins [L,3] int32, leaves [L,1,8,8,8] float32) for every 8^3 block holding any
    rho > 0, in ascending (x, y, z) origin order.
    """
    c = dim / 2.0
    ax = np.arange(dim, dtype=np.float32)
    x, y, z = np.meshgrid(ax, ax, ax, indexing="ij")
    r = np.sqrt((x - c) ** 2 + (y - c) ** 2 + (z - c) ** 2)
    rho = np.clip((radius - r) / width, 0.0, 1.0).astype(np.float32)
    nb = dim // 8
    blocks = rho.reshape(nb, 8, nb, 8, nb, 8).transpose(0, 2, 4, 1, 3, 5).reshape(-1, 8, 8, 8)
    og = np.stack(np.meshgrid(np.arange(nb), np.arange(nb), np.arange(nb), indexing="ij"), -1).reshape(-1, 3) * 8
    active = blocks.reshape(len(blocks), -1).max(axis=1) > 0
    return og[active].astype(np.int32), np.ascontiguousarray(blocks[active][:, None])


def psnr(a: np.ndarray, b: np.ndarray, peak: float = 1.0) -> float:
    """Whole-array PSNR with the notebook's formula (20*log10(peak) - 10*log10(mse + 1e-12))."""
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return 20.0 * np.log10(peak) - 10.0 * np.log10(mse + 1e-12)
