"""A/B timing of tuning builds of libvqvdb_b200.so (vqvdb_b200/build.py: VQVDB_B200_VARIANT / VQVDB_B200_DEFINES).

    python tools/ab_encode.py [--leaves N] [--decode] name=path/to/lib.so ...      (on a B200 box)

Every library runs in its own process on the same seeded leaves; prints ms per pass, leaves/s and a SHA-256 of the
result (index parity across variants must be exact: only the schedule changes, never the arithmetic).
"""
import hashlib
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(n, decode):
    sys.path.insert(0, REPO)
    import torch
    from vqvdb_b200 import BackendType, CodecConfig, IVQVAECodec
    c = IVQVAECodec.create(CodecConfig(device=CodecConfig.Device.CUDA), BackendType.B200)
    g = torch.Generator(device="cuda")
    g.manual_seed(7)
    x = torch.empty((n, 1, 8, 8, 8), dtype=torch.float32, device="cuda")
    for lo in range(0, n, 131072):
        hi = min(n, lo + 131072)
        ctrl = torch.rand((hi - lo, 1, 3, 3, 3), generator=g, device="cuda")
        x[lo:hi] = torch.nn.functional.interpolate(ctrl, size=(8, 8, 8), mode="trilinear", align_corners=True).clamp_(0, 1)
    idx = torch.empty((n, 4, 4, 4), dtype=torch.uint8, device="cuda")
    vox = torch.empty((n, 1, 8, 8, 8), dtype=torch.float32, device="cuda")
    sp = torch.cuda.current_stream().cuda_stream
    c.encode_device(x, n, idx, sp)
    fn = (lambda: c.decode_device(idx, n, vox, sp)) if decode else (lambda: c.encode_device(x, n, idx, sp))
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(5):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    out = vox if decode else idx
    print("RESULT %.3f %.0f %s" % (ms, n / ms * 1e3, hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()[:16]))


def main():
    args = sys.argv[1:]
    if args and args[0] == "--child":
        return child(int(args[1]), args[2] == "1")
    n, decode, libs = 592000, False, []
    while args:
        a = args.pop(0)
        if a == "--leaves":
            n = int(args.pop(0))
        elif a == "--decode":
            decode = True
        else:
            libs.append(a.split("=", 1))
    for name, path in libs:
        env = dict(os.environ, VQVDB_B200_LIB=os.path.abspath(path))
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(n), "1" if decode else "0"], env=env,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        res = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT")]
        if not res:
            print("%-24s FAILED\n%s" % (name, r.stdout[-800:]))
            continue
        ms, rate, sha = res[0].split()[1:]
        print("%-24s %9s ms  %12s leaves/s  sha %s" % (name, ms, rate, sha), flush=True)


if __name__ == "__main__":
    main()
