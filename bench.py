#!/usr/bin/env python
"""bench.py — 8^3 leaves/sec, encode + decode roundtrip (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--leaves L | --scaling strong --leaves-total T] [--workload float|vec3]

One "step" = one pass of the hot path over one batch: encode the leaves to uint8 indices, then decode those
indices back to voxels.  At N=1 the workload is BASELINE.json configs[2] (1 M-leaf fp32 FloatGrid roundtrip;
configs[1], decode-only, is the second half of the same step and is reported in `parts`).  For N>1 the leaf
array is sharded by contiguous leaf ranges, one rank per GPU: `--scaling weak` (default) gives every rank its
own L leaves; `--scaling strong --leaves-total 10000000` is BASELINE.json configs[4] (a 10 M-leaf grid split
over the ranks, N=1 included).  Decoded blocks are gathered to rank 0 for grid reassembly inside the timed
region — by the decode kernels' own stores over NVLink (default) or by an NCCL gather (`--gather nccl`) — and
after the timed region rank 0 checks the gathered buffer bit for bit (`gather_verified`).
`--workload vec3` is BASELINE.json configs[3] (500 k three-channel leaves, seeded weights).

Keys follow the driver contract; see DESIGN.md §5 for how each number is produced.
"""
from __future__ import annotations

import argparse
import hashlib
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
# decoder with the linear tail folded (decode_tc.cuh): stem 28.31 + res convs 2 x 14.16 + folded tail conv 14.16 MFLOP;
# what the tensor pipe actually executes (the reference's algorithm stays the yardstick for `achieved`)
FLOP_DECODE_FOLDED = (128 * 64 + 3 * 64 * 64) * 27 * 64 * 2.0
FLOP_VEC3_ENCODE = 488.72e6 + 4.19e6
FLOP_VEC3_DECODE = 399.03e6
# vec3 decoder with the folded tail (decode_tc128.cuh): five 128 -> 128 convs + three 128 -> 64 tail convs on the 4^3 grid
FLOP_VEC3_DECODE_FOLDED = (5 * 128 * 128 + 3 * 128 * 64) * 27 * 64 * 2.0
BYTES_ENCODE = 2048 + 64
BYTES_DECODE = 64 + 2048

# DRAM traffic of one launch from `ncu --set full` (dram__bytes_read.sum + dram__bytes_write.sum, per leaf).  Each entry
# is tied to the git blob hash of the kernel source it was captured from: a changed kernel reports traffic = null
# ("stale") instead of a number that no longer describes it.
NCU_DRAM = {
    "encode": {"bytes_per_leaf": (122.116608e6 + 6.356736e6) / 59200, "source": "profiles/r2b_encode_tc_ncu_summary.txt",
               "file": "vqvdb_b200/csrc/encode_tc.cu", "blob": "c2b96d5a4d10c266ef6a30518121785e48b555f7"},
    "decode": {"bytes_per_leaf": (5.199616e6 + 64.842240e6) / 59200, "source": "profiles/r2b_decode_tc_ncu_summary.txt",
               "file": "vqvdb_b200/csrc/decode_tc.cu", "blob": "6c350ada0246889c8a60decaef8450ae63d90b86"},
    # vec3 encoder = two kernels per batch (front: 31.5 MB read + 1 150.7 MB written, back: 202.1 + 111.0, per 4 144 leaves —
    # the 32 KB-per-leaf hand-over array and the write-backs of the front kernel's per-CTA scratch (x, the look-ahead
    # pre.0's partial sums, conv1's output: 384 KB per CTA, rewritten ~2 MB per leaf in L2) are what reaches DRAM)
    "encode_vec3": {"bytes_per_leaf": (31.483136e6 + 1150.738e6 + 202.092288e6 + 110.995e6) / 4144,
                    "source": "profiles/r2b_encode_tc128_ncu_summary.txt",
                    "file": ["vqvdb_b200/csrc/encode_tc128_front.cu", "vqvdb_b200/csrc/encode_tc128.cu"],
                    "blob": ["a02744ea43039dd728f757eee8d2b4db1f603502", "ff73f49e251ea4b22012bb906024aebdb4eec5ce"]},
}


def git_blob_hash(path: str) -> str:
    data = open(os.path.join(REPO, path), "rb").read()
    return hashlib.sha1(b"blob %d\0" % len(data) + data).hexdigest()


def ncu_traffic(kernel: str, leaves: int):
    """(bytes per launch, source) from the committed capture, or (None, why) when it does not describe this build."""
    e = NCU_DRAM[kernel]
    files = e["file"] if isinstance(e["file"], list) else [e["file"]]
    blobs = e["blob"] if isinstance(e["blob"], list) else [e["blob"]]
    try:
        for f, b in zip(files, blobs):
            if git_blob_hash(f) != b:
                return None, "stale: %s changed since %s was captured" % (f, e["source"])
    except OSError:
        return None, "kernel source not found"
    return e["bytes_per_leaf"] * leaves, e["source"]


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


def gen_leaves_gpu(n, device, seed, channels=1):
    """Synthetic smoke: trilinear align_corners upsample of U[0,1] 3^3 control
    grids to 8^3, clamped to [0,1]; the vec3 workload (config 4) maps it to [-1,1] per channel."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((n, channels, 8, 8, 8), dtype=torch.float32, device=device)
    step = 131072 // channels
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        ctrl = torch.rand((hi - lo, channels, 3, 3, 3), generator=g, device=device)
        v = torch.nn.functional.interpolate(ctrl, size=(8, 8, 8), mode="trilinear", align_corners=True).clamp_(0, 1)
        out[lo:hi] = v if channels == 1 else v.mul_(2).sub_(1)
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


def workload_config(args, world):
    """The `config` object shared by both arms (the driver compares them)."""
    if args.workload == "vec3":
        name = "roundtrip_500k_vec3_leaves"
        weights = "seeded EncoderVec3/DecoderVec3 (torch.manual_seed(0)) C=3 D=128 K=256"
    elif args.scaling == "strong":
        name = "roundtrip_%s_float_leaves_sharded" % ("10M" if args.leaves_total == 10_000_000 else str(args.leaves_total))
        weights = "shipped float model C=1 D=128 K=256"
    else:
        name = "roundtrip_1M_float_leaves" if args.leaves == 1_000_000 else "roundtrip_%d_float_leaves" % args.leaves
        weights = "shipped float model C=1 D=128 K=256"
    total = args.leaves_total if args.scaling == "strong" else args.leaves * world
    return {"workload": name, "leaves_per_gpu": total // world if args.scaling == "strong" else args.leaves,
            "leaves_total": total, "weights": weights}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (its LibTorch backend,
    compiled unmodified into oracle/_ref), all host threads, batch 512, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle.pyoracle import COracle, RefCodec, ref_available
    from vqvdb_b200 import synth
    cores = os.cpu_count() or 1
    sample = args.ref_sample
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    if args.workload == "vec3" or not ref_available():
        # the reference has no vec3 weights and no C++ vec3 path: its Python module classes are restated in
        # the C oracle, which is what runs here; likewise when oracle/_ref could not be built
        kind = "port"
        pack = os.path.join(REPO, "vqvdb_b200", "weights", "vqvae_vec3_seed0.vqw") if args.workload == "vec3" else None
        o = COracle(pack, threads=cores) if pack else COracle(threads=cores)
        x = synth.smoke_leaves(sample, seed=0, channels=3) if args.workload == "vec3" else synth.smoke_leaves(sample, seed=0)
        threads = o.threads
        for _ in range(warmup):
            o.decode(o.encode(x))
        t0 = time.perf_counter()
        for _ in range(steps):
            o.decode(o.encode(x))
        total = time.perf_counter() - t0
    else:
        kind = "reference"
        x = synth.smoke_leaves(sample, seed=0)
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
    value = sample * steps / total
    cfg = workload_config(args, max(1, args.gpus))
    cfg["reference_sample"] = "%d leaves per step, batch 512 (the whole workload would take minutes per step on host cores)" % sample
    line = {
        "impl": "reference", "metric": "leaves_per_sec_encode_decode", "value": value, "unit": "leaves/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * total / steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
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
                for batch in (64, 1024, 8192):
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


def parity_block(codec, x, idx, n_local, sp, sample_leaves, channels):
    """BASELINE.json's metric names "reconstruction PSNR vs reference": after the timed region, a strided sample of this
    rank's leaves goes through the checker on the host — the reference's own LibTorch CPU backend (oracle/_ref) for
    the float model, the C restatement of the reference's module classes for vec3 — and the GPU results for the same
    leaves are compared with it: uint8 index mismatches (and the reference margin at each), PSNR of the GPU
    reconstruction against the reference's, and the difference of the two reconstructions' PSNR against the input."""
    import numpy as np
    import torch
    from oracle.pyoracle import COracle, RefCodec, ref_available
    from vqvdb_b200 import synth
    m = min(sample_leaves, n_local)
    pick = torch.arange(0, m, device=x.device) * (n_local // m)
    xs = x[pick].cpu().numpy()
    idx_gpu = idx[pick].contiguous()
    rec_gpu_d = torch.empty((m, channels, 8, 8, 8), dtype=torch.float32, device=x.device)
    codec.decode_device(idx_gpu, m, rec_gpu_d, sp)
    torch.cuda.synchronize()
    idx_gpu, rec_gpu = idx_gpu.cpu().numpy(), rec_gpu_d.cpu().numpy()
    pack = os.path.join(REPO, "vqvdb_b200", "weights", "vqvae_vec3_seed0.vqw") if channels == 3 else None
    port = COracle(pack) if pack else COracle()
    idx_port, margins = port.encode(xs, with_margins=True)
    if channels == 1 and ref_available():
        checker = "reference (TorchBackend.cpp, CPU fp32)"
        ref = RefCodec("cpu")
        idx_ref = ref.encode(xs)
        rec_ref = ref.decode(idx_ref)
        ref.close()
    else:
        checker = "port (oracle/vqvae_oracle.c)"
        idx_ref, rec_ref = idx_port, port.decode(idx_port)
    mm = idx_gpu != idx_ref
    p_gpu, p_ref = synth.psnr(xs, rec_gpu), synth.psnr(xs, rec_ref)
    return {"sample_leaves": int(m), "checker": checker, "index_total": int(idx_ref.size), "index_mismatches": int(mm.sum()),
            "max_margin_at_mismatch": float(margins[mm].max()) if mm.any() else 0.0,
            "near_tie_latents_in_sample": int((margins <= 1e-4).sum()),
            "psnr_db_vs_reference_recon": float(synth.psnr(rec_gpu, rec_ref)),
            "psnr_db_input_vs_recon": float(p_gpu), "psnr_db_input_vs_reference_recon": float(p_ref), "dpsnr_db": float(p_gpu - p_ref),
            "margin_note": "margin = the reference formula's own fp32 top-2 distance gap (C restatement); <= 1e-4 is a near-tie"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="float", choices=["float", "vec3"])
    ap.add_argument("--leaves", type=int, default=None, help="leaves per GPU per step (weak scaling); default 1 M (float) / 500 k (vec3)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--leaves-total", type=int, default=10_000_000, help="--scaling strong: leaves of the whole grid (configs[4]: 10 M)")
    ap.add_argument("--ref-sample", type=int, default=16384, help="leaves per step for the CPU reference")
    ap.add_argument("--parity-sample", type=int, default=2048, help="leaves checked against the reference after the timed region")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip e2e_pageable / small batches / link probes (profiling runs)")
    ap.add_argument("--decode-precision", default="default")
    ap.add_argument("--encode-precision", default="default", help="default | fp32 (FFMA) | fp16x2_tc (tcgen05)")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="N>1: decode straight into rank 0's buffer over NVLink (CUDA IPC peer stores), or decode locally + NCCL gather")
    args = ap.parse_args()
    if args.leaves is None:
        args.leaves = 500_000 if args.workload == "vec3" else 1_000_000
    if args.workload == "vec3":
        args.ref_sample = min(args.ref_sample, 256)    # 892 MFLOP per leaf on host cores (the C restatement runs ~20-80 leaves/s)
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from vqvdb_b200 import BackendType, CodecConfig, IVQVAECodec
    from vqvdb_b200.sharding import PeerGather, gather_blocks, leaf_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 backend has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    # NCCL prints its version banner on stdout when NCCL_DEBUG is set, and the C++ host layer logs like the reference's
    # backends do (std::cout): park fd 1 on stderr for the run and print the one JSON line through the saved
    # descriptor, so stdout carries that line only
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(3, args.warmup)
    K = max(1, args.steps)
    vec3 = args.workload == "vec3"
    CH = 3 if vec3 else 1
    cfg = workload_config(args, world)
    total = cfg["leaves_total"]
    lo, hi = leaf_range(rank, world, total)
    L = hi - lo                                      # this rank's leaves

    source = {}
    if vec3:
        pack = os.path.join(REPO, "vqvdb_b200", "weights", "vqvae_vec3_seed0.vqw")
        if not os.path.exists(pack):
            subprocess.check_call([sys.executable, os.path.join(REPO, "tools", "weights_pack.py"), "vec3"], stdout=subprocess.DEVNULL)
        source = {"source": pack}
    codec = IVQVAECodec.create(CodecConfig(device=CodecConfig.Device.CUDA, device_index=local,
                                           decode_precision=args.decode_precision,
                                           encode_precision=args.encode_precision, **source), BackendType.B200)
    if codec is None:
        raise SystemExit("B200 backend failed to initialise")

    x = gen_leaves_gpu(L, dev, seed=rank, channels=CH)
    idx = torch.empty((L, 4, 4, 4), dtype=torch.uint8, device=dev)
    vox = None
    gathered, peer = None, None
    if world > 1 and args.gather == "peer":
        peer = PeerGather(codec, total, dst=0, elem_floats=512 * CH)   # rank 0 owns [total, C*512] fp32; the others map it over NVLink
    else:
        vox = torch.empty((L, CH, 8, 8, 8), dtype=torch.float32, device=dev)
        if world > 1 and rank == 0:
            gathered = torch.empty((total, CH, 8, 8, 8), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream

    def step():
        codec.encode_device(x, L, idx, sp)
        if peer is not None:  # grid reassembly on rank 0: the decode kernel's stores land in rank 0's HBM over NVLink
            codec.decode_device(idx, L, peer.my_slice_ptr, sp)
        else:
            codec.decode_device(idx, L, vox, sp)
            if world > 1:     # ... or decode locally and gather the blocks with NCCL
                gather_blocks(vox, total, dst=0, out=gathered)

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

    # ---- the gathered grid, checked bit for bit: every rank's slice of rank 0's buffer == a local decode of that rank's indices ----
    gather_verified = None
    if world > 1:
        full = peer.finish(stream) if peer is not None else gathered   # barrier: every rank's stores have landed
        all_idx = gather_blocks(idx, total, dst=0)                      # 64 B per leaf over NCCL, outside the timed region
        if rank == 0:
            ok, chunk = True, 1 << 19
            tmp = torch.empty((min(chunk, total), CH, 8, 8, 8), dtype=torch.float32, device=dev)
            full_v = full.view(total, CH * 512)
            for c0 in range(0, total, chunk):
                c1 = min(total, c0 + chunk)
                codec.decode_device(all_idx[c0:c1], c1 - c0, tmp, sp)
                torch.cuda.synchronize()
                ok = ok and bool(torch.equal(tmp[: c1 - c0].view(c1 - c0, -1), full_v[c0:c1]))
            gather_verified = ok
            del tmp, all_idx
        barrier()

    # ---- parity against the reference on a strided sample of rank 0's leaves ----
    parity = None
    if rank == 0 and args.parity_sample > 0:
        try:
            parity = parity_block(codec, x, idx, L, sp, args.parity_sample if not vec3 else min(args.parity_sample, 1024), CH)
        except Exception as e:  # noqa: BLE001
            parity = {"unavailable": str(e)[:300]}
    barrier()

    # ---- per-kernel durations (same stream, CUDA events), for the roofline of the dominant kernel ----
    if vox is None:
        vox = torch.empty((L, CH, 8, 8, 8), dtype=torch.float32, device=dev)

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

    # ---- end to end through the C-ABI with HOST buffers (pinned), H2D + D2H inside the timed region ----
    # (a rank's share above 4 M leaves is measured on its first 4 M: 2 x 8 GB of pinned host memory is enough to be
    # bandwidth-, not latency-bound, and a 10 M-leaf single-GPU run would otherwise pin 41 GB)
    Le = min(L, 4_000_000)
    hx = torch.empty((Le, CH, 8, 8, 8), dtype=torch.float32, pin_memory=True)
    hx.copy_(x[:Le])
    hidx = torch.empty((Le, 4, 4, 4), dtype=torch.uint8, pin_memory=True)
    hvox = torch.empty((Le, CH, 8, 8, 8), dtype=torch.float32, pin_memory=True)
    torch.cuda.synchronize()

    for _ in range(2):
        codec.encode_into(hx, Le, hidx)
        codec.decode_into(hidx, Le, hvox)
    barrier()
    t_enc = t_dec = 0.0
    t0 = time.perf_counter()
    for _ in range(K):
        ta = time.perf_counter()
        codec.encode_into(hx, Le, hidx)
        tb = time.perf_counter()
        codec.decode_into(hidx, Le, hvox)
        t_enc += tb - ta
        t_dec += time.perf_counter() - tb
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_checksum = float(hvox[:: max(1, Le // 1024)].double().sum())

    # ---- what the links give when every rank copies at once (the host-fed ceiling on this box) ----
    link = None
    if not args.no_extras:
        nbytes = min(Le, 1 << 19) * CH * 2048
        src_h, dst_d = hvox.view(-1)[: nbytes // 4], vox.view(-1)[: nbytes // 4]
        res = []
        for direction in ("h2d", "d2h"):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            a.record(stream)
            for _ in range(3):
                if direction == "h2d":
                    dst_d.copy_(src_h, non_blocking=True)
                else:
                    src_h.copy_(dst_d, non_blocking=True)
            b.record(stream)
            barrier()
            res.append(3 * nbytes / (a.elapsed_time(b) / 1e3) / 1e9)
        # host-side copy rate of this rank while all ranks copy (pinned -> pinned, torch CPU copy_)
        barrier()
        tc = time.perf_counter()
        hvox.view(-1)[: nbytes // 4].copy_(hx.view(-1)[: nbytes // 4])
        res.append(nbytes / (time.perf_counter() - tc) / 1e9)
        lt = torch.tensor(res, dtype=torch.float64, device=dev)
        lmin, lsum = lt.clone(), lt.clone()
        if world > 1:
            dist.all_reduce(lmin, op=dist.ReduceOp.MIN)
            dist.all_reduce(lsum, op=dist.ReduceOp.SUM)
        # what host feeding allows on this box: a step cannot be shorter than its kernels, nor than its 2 KB/leaf transfers
        # at the rate every rank gets when all ranks copy at once
        t_enc_floor = max(enc_ms / 1e3 * Le / L, Le * CH * 2048 / (float(lmin[0]) * 1e9))
        t_dec_floor = max(dec_ms / 1e3 * Le / L, Le * CH * 2048 / (float(lmin[1]) * 1e9))
        link = {"unit": "GB/s", "host_fed_ceiling_leaves_per_s": Le * world / (t_enc_floor + t_dec_floor),
                "h2d_per_rank_min": float(lmin[0]), "d2h_per_rank_min": float(lmin[1]),
                "h2d_all_ranks": float(lsum[0]), "d2h_all_ranks": float(lsum[1]),
                "host_memcpy_per_rank_min": float(lmin[2]), "host_memcpy_all_ranks": float(lsum[2]),
                "note": "pinned-memory cudaMemcpyAsync of %.2f GB and a host memcpy, every rank at the same time" % (nbytes / 1e9)}

    # ---- the reference's callers hand over pageable memory and get an owning Tensor back (VQVAECodec.cpp:48,114-127,
    #      172-196): the same roundtrip through the C++ IVQVAECodec::encode/decode virtuals of libvqvdb_b200_host.so ----
    pageable = None
    if rank == 0 and world == 1 and not vec3 and not args.no_extras:
        try:
            from vqvdb_b200.hostlib import HostBackend
            hb = HostBackend(local)
            px = hx.numpy().copy()                    # pageable
            pvox = np.empty((Le, 1, 8, 8, 8), np.float32)
            pvox.fill(0)                              # pre-faulted, as a caller-owned grid buffer is
            pidx = np.empty((Le, 4, 4, 4), np.uint8)
            pidx.fill(0)
            hb.encode_into(px, pidx); hb.decode_into(pidx, pvox)        # warm-up (staging buffers, copy threads)
            reps = max(2, min(K, 5))
            ti = sum(hb.encode_into(px, pidx) + hb.decode_into(pidx, pvox) for _ in range(reps))
            tv = te = td = 0.0
            for _ in range(reps):
                iv, s_e = hb.encode(px)
                iv = np.array(iv)                     # the reference's loop copies the indices on to its file writer
                _, s_d = hb.decode(iv)
                te += s_e; td += s_d
            tv = te + td
            same = bool(np.array_equal(iv, pidx))
            pageable = {"value": Le * reps / tv, "unit": "leaves/s",
                        "api": "IVQVAECodec::encode/decode(TensorView over pageable memory) -> owning Tensor (libvqvdb_b200_host.so)",
                        "encode_leaves_per_s": Le * reps / te, "decode_leaves_per_s": Le * reps / td,
                        "into_value": Le * reps / ti,
                        "into_api": "B200Backend::encodeInto/decodeInto on caller-owned pageable buffers (the batch loop's form: no Tensor allocation)",
                        "copy_threads": int(os.environ.get("VQVDB_B200_COPY_THREADS", "0")) or min(8, max(1, (os.cpu_count() or 2) // 2)),
                        "indices_equal_pinned_path": same,
                        "note": "decode() returns a 2 KB/leaf std::vector<std::byte> by value (IVQVAECodec.hpp:61-80): its allocation + zero-fill is inside `value`, not inside `into_value`"}
            hb.close()
            del px, pvox, pidx
        except Exception as e:  # noqa: BLE001
            pageable = {"unavailable": str(e)[:300]}

    # the reference's SOPs call the backend with 64 leaves at a time (their default batch; UI maximum 1024 for the encoder
    # SOP, 8192 for the decoder SOP): the same host-pointer calls at those sizes, synchronous, one after the other
    small = {}
    small_native = {}
    if rank == 0 and not args.no_extras:
        for nb in (64, 1024, 8192):
            if nb >= Le:
                continue
            reps = max(4, min(400, 131072 // nb))
            # raw addresses into the pinned buffers: the loop times the C-ABI calls, not Python tensor slicing
            ax, ai, av = hx.data_ptr(), hidx.data_ptr(), hvox.data_ptr()
            offs = [(i * nb) % (Le - nb) for i in range(reps)]
            for _ in range(3):
                codec.encode_into(ax, nb, ai)
                codec.decode_into(ai, nb, av)
            t0 = time.perf_counter()
            for o in offs:
                codec.encode_into(ax + o * CH * 2048, nb, ai + o * 64)
                codec.decode_into(ai + o * 64, nb, av + o * CH * 2048)
            small["batch_%d" % nb] = reps * nb / (time.perf_counter() - t0)
        # the same calls from a native loop (libvqvdb_b200_host.so: B200Backend::encodeInto / decodeInto per batch) — what a SOP's
        # C++ loop sees; the reference_gpu_backend numbers below are taken from a native loop too (oracle/ref_worker.py)
        if not vec3 and world == 1:
            try:
                from vqvdb_b200.hostlib import HostBackend
                hb2 = HostBackend(local)
                for nb in (64, 1024, 8192):
                    if nb >= Le:
                        continue
                    tot = min(Le, max(4 * nb, 131072))
                    tot -= tot % nb
                    hb2.roundtrip_batched(hx.data_ptr(), 4 * nb, nb, hidx.data_ptr(), hvox.data_ptr())      # warm-up
                    small_native["batch_%d" % nb] = tot / hb2.roundtrip_batched(hx.data_ptr(), tot, nb, hidx.data_ptr(), hvox.data_ptr())
                hb2.close()
            except Exception as e:  # noqa: BLE001
                small_native = {"unavailable": str(e)[:300]}

    t = torch.tensor([ms, e2e_s * 1e3, enc_ms, dec_ms, t_enc * 1e3, t_dec * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, enc_ms, dec_ms, e2e_enc_ms, e2e_dec_ms = [float(v) for v in t.tolist()]

    if rank == 0:
        peaks = measured_peaks()
        value = total * K / (ms / 1e3)
        flop_enc = FLOP_VEC3_ENCODE if vec3 else FLOP_ENCODE
        flop_dec = FLOP_VEC3_DECODE if vec3 else FLOP_DECODE
        bytes_enc, bytes_dec = CH * 2048 + 64, 64 + CH * 2048
        dom = "decode" if dec_ms >= enc_ms else "encode"
        dom_ms = max(dec_ms, enc_ms)
        flop = flop_dec if dom == "decode" else flop_enc
        enc_tc = codec.encode_path.startswith("fp16x2_tcgen05")
        dec_tc = codec.decode_path.startswith("bf16_tcgen05")
        dom_on_tensor = (dom == "decode" and dec_tc) or (dom == "encode" and enc_tc)
        achieved = flop * L / (dom_ms / 1e3) / 1e12
        peak = peaks["bf16_tflops_sustained"]
        ffma_peak = 71.0   # TFLOP/s, measured on this part with tools/microbench/pipe_rates.cu (nominal 74.4)
        enc_tf = flop_enc * L / (enc_ms / 1e3) / 1e12
        dec_tf = flop_dec * L / (dec_ms / 1e3) / 1e12
        dec_peak = peak if dec_tc else ffma_peak
        dec_issued_flop = ((FLOP_VEC3_DECODE_FOLDED if vec3 else FLOP_DECODE_FOLDED) if codec.decode_path.endswith("_fold") else flop_dec)
        dec_issued_tf = dec_issued_flop * L / (dec_ms / 1e3) / 1e12
        # tensor-core encoder: every GEMM is three fp16 products (hi*hi, hi*lo, lo*hi); pre.0 (221 184 MAC) stays on FFMA and
        # proj (262 144 MAC) is folded into the codebook, whose score GEMM shrinks from 64x128x256 to 64x32x256 (524 288 MAC):
        # MACs issued to the tensor pipe per leaf = 3 * (13 197 824 - 221 184 - 262 144 + 524 288)
        # vec3 (encode_tc128*.cu): the 64 -> 64, 64 -> 128 (stride 2) and 128 -> 128 convs run as three fp16 products each;
        # pre.0, proj and the distances (5.8 M of the 246.4 M MACs) stay on FFMA
        enc_tc_macs = (2 * 64 * 64 * 27 * 512 + 64 * 128 * 27 * 64 + 4 * 128 * 128 * 27 * 64) if vec3 else (13197824 - 221184 - 262144 + 524288)
        enc_issued_tf = (3 * enc_tc_macs * 2 if enc_tc else flop_enc) * L / (enc_ms / 1e3) / 1e12
        traffic, traffic_source = (None, "no capture for this path")
        if (dom == "encode" and enc_tc) or (dom == "decode" and codec.decode_path == "bf16_tcgen05_n192_fold"):
            traffic, traffic_source = ncu_traffic("encode_vec3" if vec3 else dom, L)
        enc_name = ("encode_tc128_front_kernel + encode_tc128_back_kernel" if vec3 else "encode_tc_kernel") if enc_tc else ("generic_encode_kernel" if vec3 else "encode_fp32_kernel")
        dec_name = ("decode_tc128_kernel" if vec3 else "decode_tc_kernel") if dec_tc else ("generic_decode_kernel" if vec3 else "decode_fp32_kernel")
        kernels = {
            enc_name: {
                "ms": enc_ms, "share_of_step": enc_ms / (enc_ms + dec_ms), "achieved_tflops": enc_tf,
                "pipe": "tensor (tcgen05.mma, fp16 2-way split operands = 3 products, fp32 accumulate in TMEM)" if enc_tc else "fp32 FFMA",
                "pipe_peak_tflops": peak if enc_tc else ffma_peak,
                "frac_of_pipe_peak": enc_issued_tf / (peak if enc_tc else ffma_peak),
                "issued_tflops": enc_issued_tf,
                "frac_of_bf16_tensor_peak": enc_tf / peak, "algorithmic_mflop_per_leaf": flop_enc / 1e6,
                "hbm_gbs": bytes_enc * L / (enc_ms / 1e3) / 1e9},
            dec_name: {
                "ms": dec_ms, "share_of_step": dec_ms / (enc_ms + dec_ms), "achieved_tflops": dec_tf,
                "pipe": "tensor (tcgen05.mma bf16, TMEM accumulators)" if dec_tc else "fp32 FFMA",
                "pipe_peak_tflops": dec_peak,
                # folded tail: up_conv -> PixelShuffle3D -> final run as one 64->64 conv, so fewer flops are ISSUED than the
                # reference's algorithm counts; the pipe fraction uses the issued ones
                "issued_tflops": dec_issued_tf, "issued_mflop_per_leaf": dec_issued_flop / 1e6,
                "frac_of_pipe_peak": dec_issued_tf / dec_peak,
                "frac_of_bf16_tensor_peak": dec_tf / peak, "algorithmic_mflop_per_leaf": flop_dec / 1e6,
                "hbm_gbs": bytes_dec * L / (dec_ms / 1e3) / 1e9},
        }
        cfg.update({
            "sharding": "leaf ranges, one rank per GPU" + ((", decode kernels store straight into rank 0's buffer over NVLink (CUDA IPC)" if peer is not None else ", NCCL gather of decoded blocks to rank 0") if world > 1 else ""),
            "l2": "inputs (%.2f GB/step/GPU) exceed the 126 MB L2; no explicit flush" % (L * CH * 2048 / 1e9),
            "numa": ("rank 0 bound to its GPU's NUMA " + numa) if numa else "not bound",
            "encode_path": codec.encode_path, "decode_path": codec.decode_path})
        line = {
            "metric": "leaves_per_sec_encode_decode", "value": value, "unit": "leaves/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "%s encode+VQ, %s decode" % (
                "f16x2-split operands/f32 accumulate (f32-level)" if enc_tc else "f32",
                "bf16 operands/f32 accumulate" if dec_tc else "f32"),
            "data": "synthetic",
            "config": cfg,
            "parts": {"encode_ms": enc_ms, "decode_ms": dec_ms,
                      "encode_leaves_per_s": L / (enc_ms / 1e3), "decode_leaves_per_s": L / (dec_ms / 1e3)},
            "roofline": {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak,
                         # DRAM bytes of one launch of the dominant kernel, from the committed ncu --set full capture of the
                         # same kernel source (null when the source has changed since)
                         "traffic": traffic, "traffic_source": traffic_source,
                         "algorithmic_bytes": (bytes_dec if dom == "decode" else bytes_enc) * L,
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (%s)" % peaks["source"],
                         "hbm_gbs_nonbinding": (bytes_dec if dom == "decode" else bytes_enc) * L / (dom_ms / 1e3) / 1e9,
                         "note": "compute-bound path (34 kFLOP/B); achieved = ALGORITHMIC flops / time; %s kernel runs on %s" % (dom, "tensor cores (the encoder issues 3 fp16 products per algorithmic MAC, see kernels.*.issued_tflops)" if dom_on_tensor else "fp32 FFMA (measured CUDA-core peak 71 TFLOP/s, so frac of ITS pipe is %.2f)" % (achieved / ffma_peak)),
                         "kernels": kernels},
            # every rank times the same number of leaves (Le = its share, capped at 4 M): whole-job rate = ranks x Le / time
            "e2e": {"value": Le * world * K / (e2e_ms / 1e3), "unit": "leaves/s", "leaves_per_rank_per_step": Le,
                    "h2d_bytes_per_step": Le * (CH * 2048 + 64), "d2h_bytes_per_step": Le * (64 + CH * 2048),
                    "api": "vqvdb_b200_encode + vqvdb_b200_decode on pinned host buffers", "checksum": e2e_checksum,
                    "encode_ms_per_step": e2e_enc_ms / K, "decode_ms_per_step": e2e_dec_ms / K,
                    "h2d_gbs_per_rank_during_encode": Le * CH * 2048 / (e2e_enc_ms / K / 1e3) / 1e9,
                    "d2h_gbs_per_rank_during_decode": Le * CH * 2048 / (e2e_dec_ms / K / 1e3) / 1e9,
                    "link": link},
            "parity": parity,
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if gather_verified is not None:
            line["gather_verified"] = gather_verified
        if pageable is not None:
            line["e2e_pageable"] = pageable
        if small:
            line["e2e_small_batches"] = {"unit": "leaves/s", **small, **({"native_loop": small_native} if small_native else {}),
                                         "note": "roundtrip through the same host-pointer calls, 64 / 1024 / 8192 leaves per call (the reference SOPs' default batch is 64, their UI maxima 1024 and 8192); compare reference_gpu_backend"}
        if world == 1 and not args.no_cpu_baseline and not vec3:
            try:
                line.update(cpu_baseline(args.ref_sample))
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": "leaves/s", "cores": 0, "kind": "reference",
                                        "sample": "failed: %s" % str(e)[:200]}
        elif world == 1 and not args.no_cpu_baseline:
            try:
                from oracle.pyoracle import COracle
                from vqvdb_b200 import synth
                o = COracle(os.path.join(REPO, "vqvdb_b200", "weights", "vqvae_vec3_seed0.vqw"), threads=os.cpu_count() or 1)
                xs = synth.smoke_leaves(args.ref_sample, seed=0, channels=3)
                o.decode(o.encode(xs[:64]))
                t0 = time.perf_counter()
                o.decode(o.encode(xs))
                line["cpu_baseline"] = {"value": args.ref_sample / (time.perf_counter() - t0), "unit": "leaves/s", "cores": o.threads,
                                        "kind": "port", "sample": "%d vec3 smoke leaves, encode+decode (the reference ships no C++ vec3 path)" % args.ref_sample}
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": "leaves/s", "cores": 0, "kind": "port", "sample": "failed: %s" % str(e)[:200]}
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if peer is not None:
        peer.close()
    codec.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
