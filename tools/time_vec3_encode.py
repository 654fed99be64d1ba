"""One vec3 encode of n leaves through the device-pointer entry point (for ncu launch lists: front / back kernel times)."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from vqvdb_b200 import BackendType, CodecConfig, IVQVAECodec, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
pack = os.path.join(REPO, "vqvdb_b200", "weights", "vqvae_vec3_seed0.vqw")
c = IVQVAECodec.create(CodecConfig(device=CodecConfig.Device.CUDA, source=pack), BackendType.B200)
x = np.tile(synth.smoke_leaves(1024, seed=7, channels=3, sparse=True), (max(1, n // 1024), 1, 1, 1, 1))[:n]
xd = torch.from_numpy(x).cuda()
out = torch.empty((n, 64), dtype=torch.uint8, device="cuda")
for _ in range(reps):
    c.encode_device(xd, n, out, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print(c.encode_path, n, int(out.sum()))
