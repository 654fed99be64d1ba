"""Index / reconstruction parity of the CUDA paths against the reference's golden outputs, as numbers.

    python tools/parity_report.py [out.json]          (on a B200 box; default gpurun_out/r2_parity_report.json)

For every golden of tests/golden/ (generated from the reference's shipped TorchScript blob by tools/make_goldens.py)
and every encoder path: how many uint8 indices differ from the reference's, the reference's own fp32 top-2 margin at
the worst mismatch, and how many latents of the set are near-ties at all (margin <= 1e-4: the only places where a
mismatch is tolerated, tests/conftest.py).  For every decoder path: max |d| and PSNR against the reference's
reconstruction.  The same for the vec3 goldens (reference module classes, seeded weights).  The committed copy lives
in profiles/; tests/test_gpu_parity.py asserts the bounds derived from it.
"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

import numpy as np  # noqa: E402

from vqvdb_b200 import BackendType, CodecConfig, DataType, IVQVAECodec, TensorView, synth  # noqa: E402

TIE = 1e-4
FLOAT_CASES = {
    "kat256": lambda: synth.kat_leaves(256),
    "smoke1024_seed0": lambda: synth.smoke_leaves(1024, seed=0),
    "sparse1024_seed1": lambda: synth.smoke_leaves(1024, seed=1, sparse=True),
    "noise256_seed2": lambda: synth.noise_leaves(256, seed=2),
    "fogsphere64": lambda: synth.fog_sphere_grid()[1],
    "zeros4": lambda: np.zeros((4, 1, 8, 8, 8), np.float32),
}
VEC3_CASES = {
    "vec3_smoke256_seed5": lambda: synth.smoke_leaves(256, seed=5, channels=3),
    "vec3_noise64_seed6": lambda: synth.noise_leaves(64, seed=6, channels=3),
    "vec3_sparse1024_seed7": lambda: synth.smoke_leaves(1024, seed=7, channels=3, sparse=True),
}


def golden(name):
    return np.load(os.path.join(REPO, "tests", "golden", name + ".npz"))


def enc(c, x):
    return c.encode(TensorView(np.ascontiguousarray(x), list(x.shape), DataType.FLOAT32)).buffer


def dec(c, i):
    return c.decode(TensorView(np.ascontiguousarray(i), list(i.shape), DataType.UINT8)).buffer


def index_row(got, g):
    mm = got != g["indices"]
    return {"latents": int(got.size), "mismatches": int(mm.sum()),
            "max_reference_margin_at_mismatch": float(g["margins"][mm].max()) if mm.any() else 0.0,
            "near_tie_latents": int((g["margins"] <= TIE).sum()),
            "mismatch_frac": float(mm.sum()) / max(1, got.size)}


def recon_row(rec, ref):
    return {"leaves": int(rec.shape[0]), "max_abs_diff": float(np.abs(rec - ref).max()), "psnr_db_vs_reference_recon": float(synth.psnr(rec, ref))}


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "gpurun_out", "r2_parity_report.json")
    rep = {"tie_margin": TIE, "oracle": "reference TorchScript blob, CPU fp32 (tools/make_goldens.py)", "encode": {}, "decode": {}, "vec3": {}}
    for prec in ("fp16x2_tc", "fp32"):
        c = IVQVAECodec.create(CodecConfig(device=CodecConfig.Device.CUDA, encode_precision=prec), BackendType.B200)
        rep["encode"][c.encode_path] = {name: index_row(enc(c, gen()), golden(name)) for name, gen in FLOAT_CASES.items()}
        c.close()
    for prec in ("default", "fp32"):
        c = IVQVAECodec.create(CodecConfig(device=CodecConfig.Device.CUDA, decode_precision=prec), BackendType.B200)
        rows = {}
        for name in list(FLOAT_CASES) + ["decode_random128_seed1234"]:
            g = golden(name)
            idx = synth.random_indices(128, seed=1234) if name.startswith("decode_random") else g["indices"]
            m = len(g["recon"])
            rows[name] = recon_row(dec(c, idx[:m]), g["recon"])
        rep["decode"][c.decode_path] = rows
        c.close()
    pack = os.path.join(REPO, "vqvdb_b200", "weights", "vqvae_vec3_seed0.vqw")
    if os.path.exists(pack):
        for prec in ("default", "fp32"):
            c = IVQVAECodec.create(CodecConfig(device=CodecConfig.Device.CUDA, source=pack, decode_precision=prec, encode_precision=prec), BackendType.B200)
            for name, gen in VEC3_CASES.items():
                g = golden(name)
                m = len(g["recon"])
                rr = recon_row(dec(c, g["indices"][:m]), g["recon"])
                rr["psnr_db_vs_reference_recon"] += 20.0 * np.log10(2.0)   # tanh outputs span (-1, 1): peak 2
                if prec == "default":
                    row = index_row(enc(c, gen()), g)
                    row["encode_path"] = c.encode_path
                    row["decode"] = {}
                    row["encoders"] = {}
                    rep["vec3"][name] = row
                rep["vec3"][name]["encoders"][c.encode_path] = index_row(enc(c, gen()), g)
                rep["vec3"][name]["decode"][c.decode_path] = rr
            c.close()
    worst = max(r["mismatch_frac"] for p in rep["encode"].values() for n, r in p.items() if r["latents"] >= 4096)
    rep["summary"] = {
        "kat256_mismatches": {p: r["kat256"]["mismatches"] for p, r in rep["encode"].items()},
        "worst_mismatch_frac_on_sets_of_4096_or_more": worst,
        "every_mismatch_is_a_reference_near_tie": all(r["max_reference_margin_at_mismatch"] <= TIE for p in rep["encode"].values() for r in p.values()),
    }
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as f:
        json.dump(rep, f, indent=1)
    print(json.dumps(rep["summary"]))
    for p, rows in rep["encode"].items():
        for n, r in rows.items():
            print("%-16s %-18s %6d / %7d differ   worst margin %.2e   near-ties in set %d" % (p, n, r["mismatches"], r["latents"], r["max_reference_margin_at_mismatch"], r["near_tie_latents"]))


if __name__ == "__main__":
    main()
