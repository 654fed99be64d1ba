"""Throughput of the Vec3f model (BASELINE config 4) through the device-pointer API: leaves/s encode, decode."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vqvdb_b200 import BackendType, CodecConfig, IVQVAECodec

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
pack = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vqvdb_b200", "weights", "vqvae_vec3_seed0.vqw")
c = IVQVAECodec.create(CodecConfig(device=CodecConfig.Device.CUDA, source=pack), BackendType.B200)
x = torch.rand((n, 3, 8, 8, 8), device="cuda") * 2 - 1
idx = torch.empty((n, 64), dtype=torch.uint8, device="cuda")
vox = torch.empty((n, 3, 8, 8, 8), device="cuda")
sp = torch.cuda.current_stream().cuda_stream
for name, fn in (("encode", lambda: c.encode_device(x, n, idx, sp)), ("decode", lambda: c.decode_device(idx, n, vox, sp))):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print("vec3 %s: %.1f ms for %d leaves = %.0f leaves/s" % (name, ms, n, n / ms * 1e3))
c.close()
