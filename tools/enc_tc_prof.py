"""Cycle breakdown of the tensor-core encoder (debug stage 100 of vqvdb_b200_debug_encode_tap).

Per CTA the kernel records, for row thread 0, the cycles spent in each phase of a leaf, and for the MMA issuer the cycles
waiting for operands (a_ready), for weights (w_full) and issuing.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vqvdb_b200 import BackendType, CodecConfig, IVQVAECodec

n = int(sys.argv[1]) if len(sys.argv) > 1 else 59200
codec = IVQVAECodec.create(CodecConfig(device=CodecConfig.Device.CUDA, encode_precision="fp16x2_tc"), BackendType.B200)
x = torch.rand((n, 1, 8, 8, 8), dtype=torch.float32, device="cuda")
idx = torch.empty((n, 64), dtype=torch.uint8, device="cuda")
prof = torch.zeros((148, 64), dtype=torch.float32, device="cuda")
for _ in range(2):
    prof.zero_()
    codec.debug_encode_tap(x, n, 100, prof, idx)
torch.cuda.synchronize()
p = prof.cpu().numpy()
leaves = n / 148.0
names = ["stage next leaf + its pre.0 (inside the conv2 MMA wait)", "GN pre.1 of the next leaf", "gn1 + split -> A8", "wait conv1 MMA", "conv1 epilogue: normalise, split, store",
         "wait conv2 MMA", "conv2 epilogue (residual, -> Y)", "wait down MMA", "down epilogue (GN, -> H32)", "wait res32.c1 MMA",
         "res32.c1 epilogue", "wait res32.c2 MMA", "res32.c2 epilogue + attention", "  conv1 epilogue: accumulator reads (incl. wait for tile group 2)", "  conv1 epilogue: GroupNorm statistics",
         "wait VQ MMA", "VQ scores -> bounds, two smallest", "VQ decision (+ near-tie rows: z, shortlist, exact re-scoring)", "clear Y"]
tot = 0.0
for i, nm in enumerate(names):
    c = p[:, i].mean() / leaves
    tot += c
    print("%-58s %8.0f cyc/leaf" % (nm, c))
print("%-58s %8.0f cyc/leaf" % ("row thread total", tot))
print("issuer: wait a_ready %.0f  wait w_full %.0f  issue+commit %.0f  total %.0f cyc/leaf" % tuple(p[:, 32 + i].mean() / leaves for i in range(4)))
print("issuer weight waits by phase: 8^3 convs %.0f  down %.0f  4^3 convs %.0f  VQ %.0f cyc/leaf" % tuple(p[:, 36 + i].mean() / leaves for i in range(4)))
print("issuer, operands ready -> layer's MMAs complete: 8^3 convs (2) %.0f  down %.0f  4^3 convs (2) %.0f  VQ %.0f cyc/leaf" % tuple(p[:, 40 + i].mean() / leaves for i in range(4)))
print("issuer, time inside the issue loops: 8^3 convs (180 MMAs) %.0f  down (64) %.0f  4^3 convs (72) %.0f  VQ (6) %.0f cyc/leaf" % tuple(p[:, 44 + i].mean() / leaves for i in range(4)))
