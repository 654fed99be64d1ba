"""Bring-up check of the tensor-core encoder (encode_tc.cu): every activation tap against the C oracle, then the
indices against the oracle (with its top-2 margins) and against the fp32 FFMA kernel.

    python tools/enc_tc_bringup.py [n_leaves]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle.pyoracle import COracle
from vqvdb_b200 import BackendType, CodecConfig, DataType, IVQVAECodec, TensorView, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
x = np.concatenate([synth.smoke_leaves(n // 3, seed=11), synth.smoke_leaves(n // 3, seed=12, sparse=True),
                    synth.noise_leaves(n - 2 * (n // 3), seed=13)])
n = x.shape[0]
oracle = COracle()
tc = IVQVAECodec.create(CodecConfig(device=CodecConfig.Device.CUDA, encode_precision="fp16x2_tc"), BackendType.B200)
ff = IVQVAECodec.create(CodecConfig(device=CodecConfig.Device.CUDA, encode_precision="fp32"), BackendType.B200)
assert tc is not None and ff is not None
print("encode paths:", tc.encode_path, ff.encode_path)
xd = torch.from_numpy(x).cuda()
idx = torch.zeros((n, 64), dtype=torch.uint8, device="cuda")
rc = 0
shapes = {0: (16, 512), 6: (16, 512), 1: (16, 512), 2: (32, 64), 7: (32, 64), 3: (32, 64), 4: (32, 64), 5: (128, 64)}
for stage in (0, 6, 1, 2, 7, 3, 4, 5):
    c, p = shapes[stage]
    tap = torch.full((n, c, p), float("nan"), dtype=torch.float32, device="cuda")
    tc.debug_encode_tap(xd, n, stage, tap, idx)
    torch.cuda.synchronize()
    got = tap.cpu().numpy()
    ref = (oracle.latents(x) if stage == 5 else oracle.encode_tap(x, stage)).reshape(n, c, p)
    bad = ~np.isfinite(got)
    err = np.abs(np.where(bad, 0, got) - ref)
    print("stage %d: max|ref| %.3f  max err %.3e  rms err %.3e  non-finite %d" % (stage, np.abs(ref).max(), err.max(), np.sqrt((err ** 2).mean()), bad.sum()))
    if bad.any() or err.max() > 2e-4:
        rc = 1
        w = np.argwhere((err > 2e-4) | bad)
        print("   first bad entries (leaf, ch, pos):", w[:6].tolist(), " got", got[tuple(w[0])], "ref", ref[tuple(w[0])])
# who is right at `down`?  fp64 convolution of the oracle's stage-1 activation, weights from the pack
try:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    import weights_pack as wp
    _meta, tens = wp.read_pack(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vqvdb_b200", "weights", "vqvae_float.vqw"))
    wd = torch.from_numpy(np.asarray(tens["encoder.down.weight"])).double()
    bd = torch.from_numpy(np.asarray(tens["encoder.down.bias"])).double()
    x1 = torch.from_numpy(oracle.encode_tap(x, 1)).double()
    ref64 = torch.nn.functional.conv3d(x1, wd, bd, stride=2, padding=1).numpy().reshape(n, 32, 64)
    tap = torch.zeros((n, 32, 64), dtype=torch.float32, device="cuda")
    tc.debug_encode_tap(xd, n, 2, tap, idx)
    torch.cuda.synchronize()
    got = tap.cpu().numpy()
    orc = oracle.encode_tap(x, 2).reshape(n, 32, 64)
    for nm, a in (("tcgen05", got), ("C oracle", orc)):
        e = a - ref64
        print("down vs fp64: %-8s max err %.3e rms %.3e mean (bias) %.3e" % (nm, np.abs(e).max(), np.sqrt((e ** 2).mean()), e.mean()))
    sgn = np.sign(ref64)
    print("   tcgen05 error projected on sign(ref) (negative = truncation toward zero): %.3e" % float(((got - ref64) * sgn).mean()))
except Exception as ex:  # diagnostic only
    print("fp64 down check skipped:", ex)
idx_o, margins = oracle.encode(x, with_margins=True)
idx_tc = tc.encode(TensorView(x, list(x.shape), DataType.FLOAT32)).buffer
idx_ff = ff.encode(TensorView(x, list(x.shape), DataType.FLOAT32)).buffer
for name, got in (("tcgen05", idx_tc), ("ffma", idx_ff)):
    mm = got != idx_o
    print("%s vs oracle: %d of %d indices differ; largest oracle margin at a mismatch %.3e" % (
        name, mm.sum(), mm.size, float(margins[mm].max()) if mm.any() else 0.0))
    if mm.any() and float(margins[mm].max()) > 1e-4:
        rc = 1
print("tcgen05 vs ffma: %d differ" % int((idx_tc != idx_ff).sum()))
print("codes used:", len(np.unique(idx_tc)))
sys.exit(rc)
