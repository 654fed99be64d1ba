"""ncu driver for the vec3 model's fp32 checking encoder (encode_generic_kernel); the tensor-core encoder: tools/time_vec3_encode.py."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from bench import gen_leaves_gpu
from vqvdb_b200 import BackendType, CodecConfig, IVQVAECodec
pack = os.path.join(os.getcwd(), "vqvdb_b200", "weights", "vqvae_vec3_seed0.vqw")
c = IVQVAECodec.create(CodecConfig(device=CodecConfig.Device.CUDA, source=pack, encode_precision="fp32"), BackendType.B200)
n = 444 * 4
x = gen_leaves_gpu(n, torch.device("cuda", 0), 0, channels=3)
idx = torch.empty((n, 4, 4, 4), dtype=torch.uint8, device="cuda")
for _ in range(3):
    c.encode_device(x, n, idx, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("done")
