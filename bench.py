#!/usr/bin/env python
"""bench.py — 8^3 leaves/sec, encode + decode roundtrip (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--leaves L]

One "step" = one pass of the hot path over one batch: encode L leaves to uint8 indices, then decode
those indices back to voxels.  At N=1 the workload is BASELINE.json configs[2] (1 M-leaf fp32
FloatGrid roundtrip; configs[1], decode-only, is the second half of the same step and is reported
in `parts`).  For N>1 every rank processes its own L leaves (leaf-range sharding, weak scaling) and
the decoded blocks are gathered to rank 0 over NCCL for grid reassembly inside the timed region.

Keys follow the driver contract; see DESIGN.md §Measurement for how each number is produced.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

# Algorithmic work per leaf: dense MACs x 2.
FLOP_ENCODE = 26.40e6 + 4.19e6      # encoder + VQ distance GEMM
FLOP_DECODE = 114.14e6
# decoder with the linear tail folded (decode_mma.cuh): stem 28.31 + res convs 2 x 14.16 + folded tail conv 14.16 MFLOP;
# what the tensor pipe actually executes on the *_fold path (the reference's algorithm stays the yardstick for `achieved`)
FLOP_DECODE_FOLDED = (128 * 64 + 3 * 64 * 64) * 27 * 64 * 2.0


# ncu --set full, 59 200 leaves: encode_tc_kernel read 122.51 MB + wrote 7.90 MB; decode_tc2_kernel<fold> read 5.05 MB + wrote 69.74 MB
NCU_DRAM_BYTES_PER_LEAF = {"encode": (122.509312e6 + 7.897856e6) / 59200, "decode": (5.050880e6 + 69.744896e6) / 59200}
NCU_DRAM_SOURCE = {"encode": "profiles/r1e_encode_tc_ncu_summary.txt", "decode": "profiles/r1d_decode_tc2_ncu_summary.txt"}


def dec_kernel_name(path: str) -> str:
    return {"bf16_tcgen05_n192_fold": "decode_tc2_kernel<fold>", "bf16_tcgen05_n192": "decode_tc2_kernel", "bf16_tcgen05": "decode_tc_kernel",
            "bf16_mma": "decode_mma_kernel"}.get(path, "decode_fp32_kernel")
BYTES_ENCODE = 2048 + 64
BYTES_DECODE = 64 + 2048


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p["bf16_tflops_sustained"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower() == "active":
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


def gen_leaves_gpu(n, device, seed):
    """Synthetic smoke: trilinear align_corners upsample of U[0,1] 3^3 control
    grids to 8^3, clamped to [0,1]."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((n, 1, 8, 8, 8), dtype=torch.float32, device=device)
    step = 131072
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        ctrl = torch.rand((hi - lo, 1, 3, 3, 3), generator=g, device=device)
        out[lo:hi] = torch.nn.functional.interpolate(ctrl, size=(8, 8, 8), mode="trilinear", align_corners=True).clamp_(0, 1)
    return out


def bind_to_gpu_numa_node(local_rank: int):
    """One process per GPU: run this rank's host threads — and therefore allocate its pinned staging buffers — on the
    NUMA node the GPU's PCIe root hangs off, so the H2D / D2H streams of the 8 ranks do not cross the socket
    interconnect.  Returns a short description for the JSON line, or None when the topology cannot be read."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return "node %d (%d cpus)" % (node, len(cpus))
    except Exception:
        return None


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (its LibTorch backend,
    compiled unmodified into oracle/_ref), all host threads, batch 512, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    from oracle.pyoracle import COracle, RefCodec, ref_available
    from vqvdb_b200 import synth
    cores = os.cpu_count() or 1
    sample = args.ref_sample
    x = synth.smoke_leaves(sample, seed=0)
    steps, warmup = args.steps, max(1, min(args.warmup, 2))
    if ref_available():
        kind = "reference"
        ref = RefCodec("cpu", threads=cores)
        tmp = tempfile.mkdtemp(prefix="vqvdb_bench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        path = os.path.join(tmp, "x.bin")
        x.tofile(path)
        ref.proc.stdin.write("bench roundtrip %s %d %d %d %d\n" % (path, sample, 512, steps, warmup))
        ref.proc.stdin.flush()
        resp = ref._readline().split()
        assert resp[0] == "ok", resp
        total = float(resp[1])
        threads = ref.threads
        ref.close()
        os.remove(path); os.rmdir(tmp)
    else:
        kind = "port"
        o = COracle(threads=cores)
        threads = o.threads
        for _ in range(warmup):
            o.decode(o.encode(x))
        t0 = time.perf_counter()
        for _ in range(steps):
            o.decode(o.encode(x))
        total = time.perf_counter() - t0
    value = sample * steps / total
    line = {
        "impl": "reference", "metric": "leaves_per_sec_encode_decode", "value": value, "unit": "leaves/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * total / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "roundtrip_1M_float_leaves", "leaves_per_gpu": args.leaves, "model": None,
                   "note": "reference CPU backend (LibTorch, TorchBackend.cpp) timed on a %d-leaf sample per step, batch 512" % sample},
        "cpu_baseline": {"value": value, "unit": "leaves/s", "cores": threads, "kind": kind,
                         "sample": "%d smoke leaves x %d steps, batch 512, encode+decode" % (sample, steps)},
        "e2e": {"value": value, "unit": "leaves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def cpu_baseline(sample, reference_gpu=True):
    """Reference backend on the box's host cores (rank 0, N=1 only), plus — informational — the same
    reference backend with Device::CUDA, i.e. 'the reference's own GPU backend' of the 10x target."""
    from oracle.pyoracle import COracle, RefCodec, ref_available
    from vqvdb_b200 import synth
    cores = os.cpu_count() or 1
    x = synth.smoke_leaves(sample, seed=0)
    out = {}
    if ref_available():
        ref = RefCodec("cpu", threads=cores)
        tmp = tempfile.mkdtemp(prefix="vqvdb_bench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        path = os.path.join(tmp, "x.bin")
        x.tofile(path)
        try:
            ref.proc.stdin.write("bench roundtrip %s %d %d %d %d\n" % (path, sample, 512, 1, 1))
            ref.proc.stdin.flush()
            resp = ref._readline().split()
            total = float(resp[1])
            out["cpu_baseline"] = {"value": sample / total, "unit": "leaves/s", "cores": ref.threads, "kind": "reference",
                                   "sample": "%d smoke leaves, batch 512, encode+decode, 1 warm-up pass" % sample}
        finally:
            ref.close()
        if reference_gpu:
            try:
                big = 8 * sample
                xb = synth.smoke_leaves(big, seed=1)
                pb = os.path.join(tmp, "xb.bin")
                xb.tofile(pb)
                rg = RefCodec("cuda")
                res = {}
                for batch in (64, 8192):
                    rg.proc.stdin.write("bench roundtrip %s %d %d %d %d\n" % (pb, big, batch, 2, 1))
                    rg.proc.stdin.flush()
                    resp = rg._readline().split()
                    if resp[0] == "ok":
                        res["batch_%d" % batch] = 2 * big / float(resp[1])
                rg.close()
                out["reference_gpu_backend"] = {"unit": "leaves/s", "leaves": big, **res,
                                                "note": "reference TorchBackend Device::CUDA through its host-pointer API, torch defaults (TF32 conv allowed)"}
                os.remove(pb)
            except Exception as e:  # noqa: BLE001 — informational only
                out["reference_gpu_backend"] = {"unavailable": str(e)[:200]}
        os.remove(path); os.rmdir(tmp)
    else:
        o = COracle(threads=cores)
        o.decode(o.encode(x[:256]))
        t0 = time.perf_counter()
        o.decode(o.encode(x))
        out["cpu_baseline"] = {"value": sample / (time.perf_counter() - t0), "unit": "leaves/s", "cores": o.threads,
                               "kind": "port", "sample": "%d smoke leaves, encode+decode" % sample}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--leaves", type=int, default=1_000_000, help="leaves per GPU per step")
    ap.add_argument("--ref-sample", type=int, default=16384, help="leaves per step for the CPU reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--decode-precision", default="default")
    ap.add_argument("--encode-precision", default="default", help="default | fp32 (FFMA) | fp16x2_tc (tcgen05)")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="N>1: decode straight into rank 0's buffer over NVLink (CUDA IPC peer stores), or decode locally + NCCL gather")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from vqvdb_b200 import BackendType, CodecConfig, IVQVAECodec
    from vqvdb_b200.sharding import PeerGather, gather_blocks

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 backend has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    json_fd = None
    if world > 1:
        # NCCL prints its version banner on stdout when NCCL_DEBUG is set: park fd 1 on stderr for the run and print
        # the one JSON line through the saved descriptor, so stdout carries that line only
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    W = max(3, args.warmup)
    K = max(1, args.steps)
    L = args.leaves

    codec = IVQVAECodec.create(CodecConfig(device=CodecConfig.Device.CUDA, device_index=local,
                                           decode_precision=args.decode_precision,
                                           encode_precision=args.encode_precision), BackendType.B200)
    if codec is None:
        raise SystemExit("B200 backend failed to initialise")

    x = gen_leaves_gpu(L, dev, seed=rank)
    idx = torch.empty((L, 4, 4, 4), dtype=torch.uint8, device=dev)
    vox = torch.empty((L, 1, 8, 8, 8), dtype=torch.float32, device=dev)
    gathered, peer = None, None
    if world > 1 and args.gather == "peer":
        peer = PeerGather(codec, L * world, dst=0)     # rank 0 owns [L*world, 512] fp32; the others map it over NVLink
    elif world > 1 and rank == 0:
        gathered = torch.empty((L * world, 1, 8, 8, 8), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream

    def step():
        codec.encode_device(x, L, idx, sp)
        if peer is not None:  # grid reassembly on rank 0: the decode kernel's stores land in rank 0's HBM over NVLink
            codec.decode_device(idx, L, peer.my_slice_ptr, sp)
        else:
            codec.decode_device(idx, L, vox, sp)
            if world > 1:     # ... or decode locally and gather the blocks with NCCL
                gather_blocks(vox, L * world, dst=0, out=gathered)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = codec.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(K):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = codec.kernel_launches - launches0
    clocks = sampler.stop() if rank == 0 else None

    # per-kernel durations (same stream, CUDA events), for the roofline of the dominant kernel
    def time_kernel(fn, reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream)
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    enc_ms = time_kernel(lambda: codec.encode_device(x, L, idx, sp), K)
    dec_ms = time_kernel(lambda: codec.decode_device(idx, L, vox, sp), K)

    # end to end through the C-ABI with HOST buffers (pinned), H2D + D2H inside the timed region
    hx = torch.empty((L, 1, 8, 8, 8), dtype=torch.float32, pin_memory=True)
    hx.copy_(x)
    hidx = torch.empty((L, 4, 4, 4), dtype=torch.uint8, pin_memory=True)
    hvox = torch.empty((L, 1, 8, 8, 8), dtype=torch.float32, pin_memory=True)
    torch.cuda.synchronize()

    def e2e_step():
        codec.encode_into(hx, L, hidx)
        codec.decode_into(hidx, L, hvox)
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_checksum = float(hvox[:: max(1, L // 1024)].double().sum())

    # the reference's SOPs call the backend with 64 leaves at a time (their default batch): the same host-pointer calls at
    # that size, synchronous, one after the other — what a drop-in replacement sees before its caller batches more
    small = {}
    if rank == 0:
        for nb in (64, 8192):
            reps = max(4, min(400, 131072 // nb))
            for _ in range(3):
                codec.encode_into(hx[:nb], nb, hidx[:nb])
                codec.decode_into(hidx[:nb], nb, hvox[:nb])
            t0 = time.perf_counter()
            for i in range(reps):
                lo = (i * nb) % (L - nb)
                codec.encode_into(hx[lo:lo + nb], nb, hidx[lo:lo + nb])
                codec.decode_into(hidx[lo:lo + nb], nb, hvox[lo:lo + nb])
            small["batch_%d" % nb] = reps * nb / (time.perf_counter() - t0)

    t = torch.tensor([ms, e2e_s * 1e3, enc_ms, dec_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, enc_ms, dec_ms = [float(v) for v in t.tolist()]

    if rank == 0:
        peaks = measured_peaks()
        value = L * world * K / (ms / 1e3)
        dom = "decode" if dec_ms >= enc_ms else "encode"
        dom_ms = max(dec_ms, enc_ms)
        flop = FLOP_DECODE if dom == "decode" else FLOP_ENCODE
        tensor_path = codec.decode_path != "fp32"
        dom_on_tensor = (dom == "decode" and tensor_path) or (dom == "encode" and codec.encode_path == "fp16x2_tcgen05")
        achieved = flop * L / (dom_ms / 1e3) / 1e12
        peak = peaks["bf16_tflops_sustained"]
        ffma_peak = 71.0   # TFLOP/s, measured on this part with tools/microbench/pipe_rates.cu (nominal 74.4)
        hmma_peak = 555.0  # TFLOP/s, legacy mma.sync bf16 path, same micro-benchmark
        enc_tf = FLOP_ENCODE * L / (enc_ms / 1e3) / 1e12
        dec_tf = FLOP_DECODE * L / (dec_ms / 1e3) / 1e12
        enc_tc = codec.encode_path == "fp16x2_tcgen05"
        dec_tc = codec.decode_path.startswith("bf16_tcgen05")
        dec_peak = peak if dec_tc else hmma_peak if tensor_path else ffma_peak
        dec_issued_flop = FLOP_DECODE_FOLDED if codec.decode_path.endswith("_fold") else FLOP_DECODE
        dec_issued_tf = dec_issued_flop * L / (dec_ms / 1e3) / 1e12
        ncu_applies = (dom == "encode" and enc_tc) or (dom == "decode" and codec.decode_path == "bf16_tcgen05_n192_fold")
        # tensor-core encoder: every GEMM is three fp16 products (hi*hi, hi*lo, lo*hi); pre.0 (221 184 MAC) stays on FFMA and
        # proj (262 144 MAC) is folded into the codebook, whose score GEMM shrinks from 64x128x256 to 64x32x256 (524 288 MAC):
        # MACs issued to the tensor pipe per leaf = 3 * (13 197 824 - 221 184 - 262 144 + 524 288)
        enc_issued_tf = 3 * (13197824 - 221184 - 262144 + 524288) * 2 * L / (enc_ms / 1e3) / 1e12
        kernels = {
            ("encode_tc_kernel" if enc_tc else "encode_fp32_kernel"): {
                "ms": enc_ms, "share_of_step": enc_ms / (enc_ms + dec_ms), "achieved_tflops": enc_tf,
                "pipe": "tensor (tcgen05.mma, fp16 2-way split operands = 3 products, fp32 accumulate in TMEM)" if enc_tc else "fp32 FFMA",
                "pipe_peak_tflops": peak if enc_tc else ffma_peak,
                "frac_of_pipe_peak": (enc_issued_tf if enc_tc else enc_tf) / (peak if enc_tc else ffma_peak),
                "issued_tflops": enc_issued_tf if enc_tc else enc_tf,
                "frac_of_bf16_tensor_peak": enc_tf / peak, "algorithmic_mflop_per_leaf": FLOP_ENCODE / 1e6,
                "hbm_gbs": BYTES_ENCODE * L / (enc_ms / 1e3) / 1e9},
            dec_kernel_name(codec.decode_path): {
                "ms": dec_ms, "share_of_step": dec_ms / (enc_ms + dec_ms), "achieved_tflops": dec_tf,
                "pipe": ("tensor (tcgen05.mma bf16, TMEM accumulators)" if dec_tc else "tensor (mma.sync bf16)" if tensor_path else "fp32 FFMA"),
                "pipe_peak_tflops": dec_peak,
                # folded tail: up_conv -> PixelShuffle3D -> final run as one 64->64 conv, so fewer flops are ISSUED than the
                # reference's algorithm counts; the pipe fraction uses the issued ones
                "issued_tflops": dec_issued_tf, "issued_mflop_per_leaf": dec_issued_flop / 1e6,
                "frac_of_pipe_peak": dec_issued_tf / dec_peak,
                "frac_of_bf16_tensor_peak": dec_tf / peak, "algorithmic_mflop_per_leaf": FLOP_DECODE / 1e6,
                "hbm_gbs": BYTES_DECODE * L / (dec_ms / 1e3) / 1e9},
        }
        line = {
            "metric": "leaves_per_sec_encode_decode", "value": value, "unit": "leaves/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "%s encode+VQ, %s decode" % (
                "f16x2-split operands/f32 accumulate (f32-level)" if codec.encode_path == "fp16x2_tcgen05" else "f32",
                "bf16 operands/f32 accumulate" if codec.decode_path != "fp32" else "f32"),
            "data": "synthetic",
            "config": {"workload": "roundtrip_1M_float_leaves", "leaves_per_gpu": L, "weights": "shipped float model C=1 D=128 K=256",
                       "sharding": "leaf ranges, one rank per GPU" + ((", decode kernels store straight into rank 0's buffer over NVLink (CUDA IPC)" if peer is not None else ", NCCL gather of decoded blocks to rank 0") if world > 1 else ""),
                       "l2": "inputs (%.2f GB/step) exceed the 126 MB L2; no explicit flush" % (L * 2048 / 1e9),
                       "numa": ("rank 0 bound to its GPU's NUMA " + numa) if numa else "not bound",
                       "encode_path": codec.encode_path, "decode_path": codec.decode_path},
            "parts": {"encode_ms": enc_ms, "decode_ms": dec_ms,
                      "encode_leaves_per_s": L / (enc_ms / 1e3), "decode_leaves_per_s": L / (dec_ms / 1e3)},
            "roofline": {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak,
                         # DRAM bytes of one launch of the dominant kernel, from the committed ncu --set full captures
                         # (dram__bytes_read.sum + dram__bytes_write.sum per leaf, scaled to this launch's L leaves)
                         "traffic": (NCU_DRAM_BYTES_PER_LEAF[dom] * L) if ncu_applies else None, "traffic_source": NCU_DRAM_SOURCE[dom] if ncu_applies else None,
                         "algorithmic_bytes": (BYTES_DECODE if dom == "decode" else BYTES_ENCODE) * L,
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (%s)" % peaks["source"],
                         "hbm_gbs_nonbinding": (BYTES_DECODE if dom == "decode" else BYTES_ENCODE) * L / (dom_ms / 1e3) / 1e9,
                         "note": "compute-bound path (34 kFLOP/B); achieved = ALGORITHMIC flops / time; %s kernel runs on %s" % (dom, "tensor cores (the encoder issues 3 fp16 products per algorithmic MAC, see kernels.*.issued_tflops)" if dom_on_tensor else "fp32 FFMA (measured CUDA-core peak 71 TFLOP/s, so frac of ITS pipe is %.2f)" % (achieved / ffma_peak)),
                         "kernels": kernels},
            "e2e": {"value": L * world * K / (e2e_ms / 1e3), "unit": "leaves/s",
                    "h2d_bytes_per_step": L * (2048 + 64), "d2h_bytes_per_step": L * (64 + 2048),
                    "api": "vqvdb_b200_encode + vqvdb_b200_decode on pinned host buffers", "checksum": e2e_checksum},
            "gpu_launches": launches,
            "clocks": clocks,
            "e2e_small_batches": {"unit": "leaves/s", **small,
                                  "note": "roundtrip through the same host-pointer calls, 64 / 8192 leaves per call (the reference SOPs' default batch is 64); compare reference_gpu_backend"},
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                line.update(cpu_baseline(args.ref_sample))
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": "leaves/s", "cores": 0, "kind": "reference",
                                        "sample": "failed: %s" % str(e)[:200]}
        if json_fd is not None:
            os.write(json_fd, (json.dumps(line) + "\n").encode())
        else:
            print(json.dumps(line))
    if peer is not None:
        peer.close()
    codec.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
