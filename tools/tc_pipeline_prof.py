"""Cycle breakdown of decode_tc_kernel (kProf build): run on a B200 box.

    python tools/tc_pipeline_prof.py [leaves]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from vqvdb_b200 import BackendType, CodecConfig, IVQVAECodec  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 4 * 50
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16_tc"
upg = 45
codec = IVQVAECodec.create(CodecConfig(device=CodecConfig.Device.CUDA, decode_precision=prec), BackendType.B200)
idx = torch.randint(0, 256, (n, 4, 4, 4), dtype=torch.uint8, device="cuda")
vox = torch.empty((n, 1, 8, 8, 8), dtype=torch.float32, device="cuda")
threads = 608
prof = torch.zeros((148 * threads * 4 + 148 * 8,), dtype=torch.float32, device="cuda")
sp = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    codec.debug_decode_tap(idx, n, 100, prof, vox, sp)
torch.cuda.synchronize()
pall = prof.cpu().numpy()
p = pall[:148 * threads * 4].reshape(148, threads, 4)
groups = n / 4 / 148
ep = pall[148 * threads * 4:].reshape(148, 8).mean(axis=0) / groups
units = groups * upg
w = p[:, :512, :]
iss = p[:, 512:576:32, :]
print("%s: groups per CTA: %.0f (4 leaves each, %d units)" % (prec, groups, upg))
print("worker per unit: wait a_empty %.0f  stage %.0f  | wait d_full per group (4 layers) %.0f" % (
    w[..., 0].mean() / units, w[..., 1].mean() / units, w[..., 2].mean() / groups))
print("issuer per unit: wait w_full %.0f  wait a_full %.0f  issue %.0f  total %.0f" % (
    iss[..., 0].mean() / units, iss[..., 1].mean() / units, iss[..., 2].mean() / units, iss[..., 3].mean() / units))
names = ["gather (codebook rows)", "stem epilogue (GN, x, gn1)", "res conv1 epilogue", "res conv2 epilogue + attention",
         "folded tail: accumulator -> G planes", "folded tail: 8-term gather, sigmoid, store", "staging + waiting for the MMAs", "store"]
print("thread 0, cycles per group of 4 leaves:")
for nm, v in zip(names, ep):
    print("  %-44s %9.0f" % (nm, v))
print("  %-44s %9.0f  (= %.0f cycles per leaf per SM)" % ("total", ep.sum(), ep.sum() / 4))
