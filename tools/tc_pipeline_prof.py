"""Cycle breakdown of the tcgen05 decoder's pipeline (debug stage 100 of vqvdb_b200_debug_decode_tap).

Per thread the kernel records: workers  -> [cycles waiting for a free A buffer, cycles staging, cycles waiting for d_full]
                               issuers  -> [waiting w_full, waiting a_full, issuing MMAs+commits, total]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from vqvdb_b200 import BackendType, CodecConfig, IVQVAECodec

n = int(sys.argv[1]) if len(sys.argv) > 1 else 59200
codec = IVQVAECodec.create(CodecConfig(device=CodecConfig.Device.CUDA, decode_precision="bf16_tc"), BackendType.B200)
idx = torch.randint(0, 256, (n, 4, 4, 4), dtype=torch.uint8, device="cuda")
vox = torch.empty((n, 1, 8, 8, 8), dtype=torch.float32, device="cuda")
threads = 640
prof = torch.zeros((148 * threads * 4 + 148 * 8,), dtype=torch.float32, device="cuda")
sp = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    codec.debug_decode_tap(idx, n, 100, prof, vox, sp)
torch.cuda.synchronize()
pall = prof.cpu().numpy()
p = pall[:148 * threads * 4].reshape(148, threads, 4)
ep = pall[148 * threads * 4:].reshape(148, 8).mean(axis=0) / (n / 8 / 148)
units = (n / 8 / 148) * 216
w = p[:, :512, :]
lane0 = w[:, ::32, :]
iss = p[:, 512::32, :][:, :4, :]
print("units per CTA: %.0f" % units)
print("worker lane0 per unit: wait a_empty %.0f  stage %.0f  | wait d_full per group(7 passes) %.0f" % (
    lane0[..., 0].mean() / units, lane0[..., 1].mean() / units, lane0[..., 2].mean() / (units / 216)))
print("worker other lanes per unit: wait(incl syncwarp) %.0f stage %.0f" % (w[..., 0].mean() / units, w[..., 1].mean() / units))
print("issuer per unit: wait w_full %.0f  wait a_full %.0f  issue %.0f  total %.0f" % (
    iss[..., 0].mean() / units, iss[..., 1].mean() / units, iss[..., 2].mean() / units, iss[..., 3].mean() / units))
names = ["gather (codebook rows)", "stem epilogue (GN, x, gn1)", "res conv1 epilogue", "res conv2 epilogue + attention",
         "up_conv pass: accumulator -> P (x4)", "up_conv pass: final conv on FFMA (x4)", "staging + waiting for the MMAs", "sigmoid + store"]
print("thread 0, cycles per group of 8 leaves:")
for nm, v in zip(names, ep):
    print("  %-44s %9.0f" % (nm, v))
print("  %-44s %9.0f" % ("total", ep.sum()))
