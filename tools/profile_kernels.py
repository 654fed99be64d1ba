"""Small driver for ncu captures: runs the encode and decode kernels a few times on synthetic leaves.

    ncu --set full --clock-control none --import-source on -k regex:encode_fp32 -s 2 -c 1 \
        -o gpurun_out/prof_encode python tools/profile_kernels.py --leaves 59200
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from bench import gen_leaves_gpu  # noqa: E402
from vqvdb_b200 import BackendType, CodecConfig, IVQVAECodec  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--leaves", type=int, default=59200)   # 148 SMs x 400
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--decode-precision", default="default")
ap.add_argument("--encode-precision", default="default")
ap.add_argument("--vec3-decode", action="store_true", help="profile the vec3 model's decode_tc128_kernel instead (random indices)")
a = ap.parse_args()
if a.vec3_decode:
    pack = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vqvdb_b200", "weights", "vqvae_vec3_seed0.vqw")
    codec = IVQVAECodec.create(CodecConfig(device=CodecConfig.Device.CUDA, source=pack), BackendType.B200)
    assert codec is not None and codec.decode_path == "bf16_tcgen05_c128_fold"
    idx = torch.randint(0, 256, (a.leaves, 4, 4, 4), dtype=torch.uint8, device="cuda")
    vox = torch.empty((a.leaves, 3, 8, 8, 8), dtype=torch.float32, device="cuda")
    for _ in range(a.iters):
        codec.decode_device(idx, a.leaves, vox, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    print("profiled %d vec3 leaves x %d iters, decode path %s" % (a.leaves, a.iters, codec.decode_path))
    codec.close()
    sys.exit(0)
dev = torch.device("cuda", 0)
codec = IVQVAECodec.create(CodecConfig(device=CodecConfig.Device.CUDA, decode_precision=a.decode_precision, encode_precision=a.encode_precision), BackendType.B200)
assert codec is not None
x = gen_leaves_gpu(a.leaves, dev, seed=0)
idx = torch.empty((a.leaves, 4, 4, 4), dtype=torch.uint8, device=dev)
vox = torch.empty((a.leaves, 1, 8, 8, 8), dtype=torch.float32, device=dev)
sp = torch.cuda.current_stream().cuda_stream
for _ in range(a.iters):
    codec.encode_device(x, a.leaves, idx, sp)
    codec.decode_device(idx, a.leaves, vox, sp)
torch.cuda.synchronize()
print("profiled %d leaves x %d iters, decode path %s" % (a.leaves, a.iters, codec.decode_path))
codec.close()
