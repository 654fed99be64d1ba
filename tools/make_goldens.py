"""Generate tests/golden/*.npz from the reference itself (dev container only).

The reference has no tests, fixtures or known-answer vectors for the encode/decode
path, so the pin is the reference's own shipped model run on CPU in fp32:
the TorchScript blob of src/Bin/bin_model.h, loaded with torch.jit.load and driven
through its exported encode()/decode() methods (python/save_for_inference.py:74-104) —
the same two methods TorchBackend.cpp:149,180 call.  Inputs come from the seeded
generators in vqvdb_b200/synth.py, so only hashes of the inputs are stored.

Each golden holds:
    input_sha256      hash of the float32 input bytes (guards generator drift)
    indices           uint8 [n,4,4,4]   reference encode() -> uint8 (TorchBackend.cpp:150)
    margins           float32 [n,4,4,4] second-best minus best VQ distance (reference formula,
                      fp32) — lets a test classify an index mismatch as a near-tie
    recon             float32 [m,1,8,8,8] reference decode(indices[:m])
    recon_sum         float64 sum of decode over all n leaves

Run:  python tools/make_goldens.py [name ...]   (needs /root/reference; ~3 min on 8 cores; names restrict what is rewritten)
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, REPO)

import weights_pack as wp  # noqa: E402
from vqvdb_b200 import synth  # noqa: E402


def reference_outputs(mod, x: np.ndarray, n_recon: int):
    with torch.no_grad():
        xt = torch.from_numpy(x)
        z = mod.encoder(xt)                                   # [n,128,4,4,4]
        flat = z.permute(0, 2, 3, 4, 1).contiguous().view(-1, z.shape[1])
        emb = mod.quantizer.embedding
        dist = (torch.sum(flat ** 2, dim=1, keepdim=True) + torch.sum(emb ** 2, dim=1)
                - 2 * torch.matmul(flat, emb.t()))            # save_for_inference.py:56-60
        top2 = torch.topk(dist, 2, dim=1, largest=False).values
        margins = (top2[:, 1] - top2[:, 0]).view(-1, 4, 4, 4).numpy()
        idx64 = mod.encode(xt)
        assert torch.equal(idx64.view(-1), torch.argmin(dist, dim=1))
        idx = idx64.to(torch.uint8).numpy()
        recon = mod.decode(idx64).numpy()
    return idx, margins.astype(np.float32), recon[:n_recon].copy(), float(recon.astype(np.float64).sum())


def main():
    only = set(sys.argv[1:])
    torch.set_num_threads(os.cpu_count() or 1)
    mod = wp.load_reference_module()
    out_dir = os.path.join(REPO, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    og, fog = synth.fog_sphere_grid()
    cases = {
        "kat256": (synth.kat_leaves(256), 256),
        "smoke1024_seed0": (synth.smoke_leaves(1024, seed=0), 64),
        "sparse1024_seed1": (synth.smoke_leaves(1024, seed=1, sparse=True), 64),
        "noise256_seed2": (synth.noise_leaves(256, seed=2), 32),
        "fogsphere64": (fog, 32),
        "zeros4": (np.zeros((4, 1, 8, 8, 8), np.float32), 4),
        "nonfinite8_seed11": (synth.nonfinite_leaves(8, seed=11), 8),
    }
    for name, (x, n_recon) in cases.items():
        if only and name not in only:
            continue
        idx, margins, recon, rsum = reference_outputs(mod, x, n_recon)
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, input_sha256=np.array(hashlib.sha256(x.tobytes()).hexdigest()),
                            indices=idx, margins=margins, recon=recon, recon_sum=np.array(rsum),
                            n=np.array(x.shape[0]))
        print("%-20s n=%5d codes=%3d min_margin=%.3e  %s" % (
            name, x.shape[0], len(np.unique(idx)), float(margins.min()), path))
    # config 4: the reference's vec3 architecture (python/VQVAE_v2.py:278-325) with the seeded weights of
    # tools/weights_pack.py vec3 — the reference classes themselves, imported, run on CPU fp32.
    vec3_pack = os.path.join(REPO, "vqvdb_b200", "weights", "vqvae_vec3_seed0.vqw")
    if not os.path.exists(vec3_pack):
        wp.main(["vec3"])
    vmod = wp.load_vec3_reference_module(vec3_pack)
    vmeta, _ = wp.read_pack(vec3_pack)
    for name, gen, n_recon in (("vec3_smoke256_seed5", lambda: synth.smoke_leaves(256, seed=5, channels=3), 32),
                               ("vec3_noise64_seed6", lambda: synth.noise_leaves(64, seed=6, channels=3), 16),
                               ("vec3_sparse1024_seed7", lambda: synth.smoke_leaves(1024, seed=7, channels=3, sparse=True), 16),
                               ("vec3_nonfinite8_seed12", lambda: synth.nonfinite_leaves(8, seed=12, channels=3), 8)):
        if only and name not in only:
            continue
        x = gen()
        idx, margins, recon, rsum = reference_outputs(vmod, x, n_recon)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), input_sha256=np.array(hashlib.sha256(x.tobytes()).hexdigest()),
                            indices=idx, margins=margins, recon=recon, recon_sum=np.array(rsum), n=np.array(x.shape[0]),
                            pack_sha256=np.array(vmeta["sha256"]))
        print("%-20s n=%5d codes=%3d min_margin=%.3e" % (name, x.shape[0], len(np.unique(idx)), float(margins.min())))
    if only and "decode_random128_seed1234" not in only:
        return
    # decode-only golden: random indices straight into the reference decoder (config 2)
    ridx = synth.random_indices(128, seed=1234)
    with torch.no_grad():
        rrec = mod.decode(torch.from_numpy(ridx).long()).numpy()
    np.savez_compressed(os.path.join(out_dir, "decode_random128_seed1234.npz"),
                        input_sha256=np.array(hashlib.sha256(ridx.tobytes()).hexdigest()),
                        recon=rrec, recon_sum=np.array(float(rrec.astype(np.float64).sum())))
    print("decode_random128_seed1234 done; fog sphere leaves = %d" % len(og))


if __name__ == "__main__":
    main()
