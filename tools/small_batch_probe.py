"""Where a small host-pointer call spends its time (B200 box): kernel time by CUDA events through the device API vs the
wall time of the synchronous host API with raw addresses (no Python tensor slicing inside the loop)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from vqvdb_b200 import BackendType, CodecConfig, IVQVAECodec  # noqa: E402

c = IVQVAECodec.create(CodecConfig(device=CodecConfig.Device.CUDA), BackendType.B200)
L = c._L
sp = torch.cuda.current_stream().cuda_stream
for n in (1, 64, 148, 296, 1024, 8192):
    x = torch.rand((n, 1, 8, 8, 8), device="cuda")
    idx = torch.empty((n, 4, 4, 4), dtype=torch.uint8, device="cuda")
    vox = torch.empty((n, 1, 8, 8, 8), device="cuda")
    hx = x.cpu().pin_memory()
    hidx = torch.empty((n, 4, 4, 4), dtype=torch.uint8).pin_memory()
    hvox = torch.empty((n, 1, 8, 8, 8)).pin_memory()
    res = {}
    for name, fn in (("enc", lambda: c.encode_device(x, n, idx, sp)), ("dec", lambda: c.decode_device(idx, n, vox, sp))):
        for _ in range(5):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        reps = 50
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        res[name + "_kernel_us"] = a.elapsed_time(b) / reps * 1e3
    h = c._h
    ax, ai, av = hx.data_ptr(), hidx.data_ptr(), hvox.data_ptr()
    for name, fn in (("enc", lambda: L.vqvdb_b200_encode(h, ax, n, ai)), ("dec", lambda: L.vqvdb_b200_decode(h, ai, n, av))):
        for _ in range(5):
            fn()
        reps = 200
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        res[name + "_host_call_us"] = (time.perf_counter() - t0) / reps * 1e6
    rt = res["enc_host_call_us"] + res["dec_host_call_us"]
    print("n=%5d  kernel enc %7.1f dec %7.1f us | host call enc %7.1f dec %7.1f us | roundtrip %8.0f leaves/s" % (
        n, res["enc_kernel_us"], res["dec_kernel_us"], res["enc_host_call_us"], res["dec_host_call_us"], n / rt * 1e6), flush=True)
c.close()
