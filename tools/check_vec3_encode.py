"""Bring-up check of the tensor-core vec3 encoder (encode_tc128*.cu): stage taps against the C oracle, indices against
the goldens and the fp32 generic kernel, and a throughput line.  Runs on the GPU box (reads nothing outside the repo).

    python tools/check_vec3_encode.py [n_timing_leaves]
"""
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle.pyoracle import COracle  # noqa: E402
from vqvdb_b200 import BackendType, CodecConfig, IVQVAECodec, synth  # noqa: E402

PACK = os.path.join(REPO, "vqvdb_b200", "weights", "vqvae_vec3_seed0.vqw")


def make(enc):
    return IVQVAECodec.create(CodecConfig(device=CodecConfig.Device.CUDA, source=PACK, encode_precision=enc), BackendType.B200)


def encode_dev(c, x):
    xd = torch.from_numpy(x).cuda()
    out = torch.empty((x.shape[0], 64), dtype=torch.uint8, device="cuda")
    c.encode_device(xd, x.shape[0], out, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return out.cpu().numpy().reshape(-1, 4, 4, 4)


def main():
    n_time = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    tc, f32 = make("default"), make("fp32")
    print("paths:", tc.encode_path, "|", f32.encode_path)
    o = COracle(PACK)
    x = synth.smoke_leaves(1024, seed=7, channels=3, sparse=True)
    g = np.load(os.path.join(REPO, "tests", "golden", "vec3_sparse1024_seed7.npz"))

    n = 37
    xd = torch.from_numpy(x[:n]).cuda()
    idx_d = torch.empty((n, 64), dtype=torch.uint8, device="cuda")
    for stage, want in ((4, o.encode_tap(x[:n], 0, 64, 128)), (5, o.encode_tap(x[:n], 1, 64, 128)), (6, o.encode_tap(x[:n], 2, 64, 128)),
                        (1, o.encode_tap(x[:n], 3, 64, 128)), (2, o.encode_tap(x[:n], 4, 64, 128)), (3, o.latents(x[:n]))):
        tap = torch.zeros((n, 64 * 512 if stage in (4, 5) else 128 * 64), dtype=torch.float32, device="cuda")
        tc.debug_encode_tap(xd, n, stage, tap, idx_d, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        got = tap.cpu().numpy().reshape(want.shape)
        print("tap %d: max |err| %.3e  rms %.3e  scale %.3e" % (stage, np.abs(got - want).max(), np.sqrt(((got - want) ** 2).mean()), np.abs(want).max()))

    # more leaves than CTAs: from a CTA's second leaf on, pre.0 is computed ahead by the epilogue warps (encode_tc128_front.cu)
    n2 = 600
    xd2 = torch.from_numpy(x[:n2]).cuda()
    idx2 = torch.empty((n2, 64), dtype=torch.uint8, device="cuda")
    for stage, want in ((4, o.encode_tap(x[:n2], 0, 64, 128)), (3, o.latents(x[:n2]))):
        tap = torch.zeros((n2, 64 * 512 if stage == 4 else 128 * 64), dtype=torch.float32, device="cuda")
        tc.debug_encode_tap(xd2, n2, stage, tap, idx2, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        got = tap.cpu().numpy().reshape(want.shape)
        per_leaf = np.abs(got - want).reshape(n2, -1).max(axis=1)
        print("tap %d on %d leaves: max |err| %.3e (first leaf of a CTA: %.3e, later leaves: %.3e)" % (
            stage, n2, per_leaf.max(), per_leaf[:148].max(), per_leaf[148:].max()))

    # phase timestamps of the second leaf / pair of CTA 0 (cycles since the leaf / pair started)
    nprof = 148 * 6
    xp = torch.from_numpy(np.tile(x, (1, 1, 1, 1, 1))[:nprof]).cuda()
    ip = torch.empty((nprof, 64), dtype=torch.uint8, device="cuda")
    tap = torch.zeros((64,), dtype=torch.float32, device="cuda")
    tc.debug_encode_tap(xp, nprof, 100, tap, ip, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    t = tap.cpu().numpy()
    print("front stamps [pre conv | pre end | conv1 tiles | gn2 sweeps | conv2 tiles | split sweep | down]:", [int(v) for v in t[:8]])
    print("back stamps [prep | r0c1 | r0c2 | r1c1 | r1c2 | fc0 | fc2 | scale + split | score GEMM | scores read, decided | (near-tie rows: z | re-scored) | done]:",
          [int(v) for v in t[16:30]])

    for name, xs in (("sparse1024", x), ("smoke256", synth.smoke_leaves(256, seed=5, channels=3)), ("noise64", synth.noise_leaves(64, seed=6, channels=3))):
        gg = np.load(os.path.join(REPO, "tests", "golden", "vec3_%s_seed%d.npz" % (name, {"sparse1024": 7, "smoke256": 5, "noise64": 6}[name])))
        m = gg["indices"].shape[0]
        a, b = encode_dev(tc, xs[:m]), encode_dev(f32, xs[:m])
        mm = a != gg["indices"]
        print("%s: tc vs golden %d / %d differ (worst margin %.3e), fp32 vs golden %d, tc vs fp32 %d" % (
            name, mm.sum(), a.size, gg["margins"][mm].max() if mm.any() else 0.0, (b != gg["indices"]).sum(), (a != b).sum()))
    # census on mixed data: the tensor-core encoder against the fp32 kernel, oracle margins where they differ
    y = np.concatenate([synth.smoke_leaves(4096, seed=41, channels=3), synth.noise_leaves(1024, seed=42, channels=3),
                        synth.smoke_leaves(3072, seed=43, channels=3, sparse=True)])
    a, b = encode_dev(tc, y), encode_dev(f32, y)
    bad = np.argwhere((a != b).reshape(len(y), -1).any(axis=1)).ravel()
    line = "census: %d of %d latents differ (tc vs fp32) in %d leaves" % (int((a != b).sum()), a.size, len(bad))
    if len(bad):
        io, mo = o.encode(y[bad], with_margins=True)
        line += "; vs oracle: tc %d, fp32 %d; worst oracle margin at a tc mismatch %.3e" % (
            int((a[bad] != io).sum()), int((b[bad] != io).sum()), float(mo[a[bad] != io].max()) if (a[bad] != io).any() else 0.0)
    print(line)
    # the codebook search's two paths: stage 3 sends EVERY row through the exact fp32 re-scoring (and taps z); the default
    # decides ~99.9 % of the rows from the tensor-core scores alone.  Same indices, or the shortlist bound is wrong.
    nf = 4096
    yd = torch.from_numpy(y[:nf]).cuda()
    forced = torch.empty((nf, 64), dtype=torch.uint8, device="cuda")
    ztap = torch.zeros((nf, 128 * 64), dtype=torch.float32, device="cuda")
    tc.debug_encode_tap(yd, nf, 3, ztap, forced, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    print("exact path forced on every row vs shortlist: %d of %d latents differ" % (
        int((forced.cpu().numpy().reshape(-1, 4, 4, 4) != a[:nf]).sum()), nf * 64))
    for nn in (1, 2, 3, 297, 1023):
        assert np.array_equal(encode_dev(tc, x[:nn]), encode_dev(tc, x)[:nn]), nn
    print("ragged counts agree")

    big = np.tile(x, (max(1, n_time // 1024), 1, 1, 1, 1))
    bd = torch.from_numpy(big).cuda()
    out = torch.empty((big.shape[0], 64), dtype=torch.uint8, device="cuda")
    for c, label in ((tc, "tc"), (f32, "fp32")):
        c.encode_device(bd, big.shape[0], out, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        c.encode_device(bd, big.shape[0], out, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("%s: %d leaves in %.2f ms = %.1f k leaves/s" % (label, big.shape[0], dt * 1e3, big.shape[0] / dt / 1e3))


if __name__ == "__main__":
    main()
