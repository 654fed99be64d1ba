"""Leaf-range sharding + gather, world_size 2 and 3 on CPU (gloo)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vqvdb_b200.sharding import gather_blocks, leaf_range, shard_sizes


@pytest.mark.parametrize("world,n", [(1, 10), (2, 1000), (8, 10_000_000), (8, 1_000_003), (3, 2), (4, 0)])
def test_ranges_partition_the_leaf_array(world, n):
    edges = [leaf_range(r, world, n) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == n
    for a, b in zip(edges, edges[1:]):
        assert a[1] == b[0]
    sizes = shard_sizes(world, n)
    assert sum(sizes) == n and max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = leaf_range(rank, world, n)
        # stand-in for "decode my shard": every leaf's 512 voxels encode its global leaf number
        local = (torch.arange(lo, hi, dtype=torch.float32)[:, None] + torch.arange(512, dtype=torch.float32)[None] / 1024)
        local = local.view(hi - lo, 1, 8, 8, 8)
        full = gather_blocks(local, n, dst=0)
        idx_local = (torch.arange(lo * 64, hi * 64) % 251).to(torch.uint8).view(hi - lo, 4, 4, 4)
        idx_full = gather_blocks(idx_local, n, dst=0)
        if rank == 0:
            want = (torch.arange(n, dtype=torch.float32)[:, None] + torch.arange(512, dtype=torch.float32)[None] / 1024)
            ok = torch.equal(full.view(n, 512), want)
            ok &= torch.equal(idx_full.view(-1), (torch.arange(n * 64) % 251).to(torch.uint8))
            q.put(bool(ok))
        else:
            assert full is None and idx_full is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 1001), (3, 64)])
def test_gather_reassembles_file_order(world, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def _gpu_worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import numpy as np
    from vqvdb_b200 import BackendType, CodecConfig, IVQVAECodec, synth
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        codec = IVQVAECodec.create(CodecConfig(device=CodecConfig.Device.CUDA, device_index=rank), BackendType.B200)
        x = torch.from_numpy(synth.smoke_leaves(n, seed=11))
        lo, hi = leaf_range(rank, world, n)
        xd = x[lo:hi].cuda()
        idx = torch.empty((hi - lo, 4, 4, 4), dtype=torch.uint8, device="cuda")
        vox = torch.empty((hi - lo, 1, 8, 8, 8), dtype=torch.float32, device="cuda")
        sp = torch.cuda.current_stream().cuda_stream
        codec.encode_device(xd, hi - lo, idx, sp)
        codec.decode_device(idx, hi - lo, vox, sp)
        full_idx = gather_blocks(idx, n, dst=0)
        full_vox = gather_blocks(vox, n, dst=0)
        torch.cuda.synchronize()
        # fused form: every rank's decode kernel stores straight into rank 0's buffer (CUDA IPC, NVLink)
        from vqvdb_b200.sharding import PeerGather
        pg = PeerGather(codec, n, dst=0)
        codec.decode_device(idx, hi - lo, pg.my_slice_ptr, sp)
        fused = pg.finish()
        if rank == 0:
            fused = fused.view(n, 1, 8, 8, 8).clone()
        pg.close()
        if rank == 0:
            assert torch.equal(fused, full_vox), "peer-store gather differs from the NCCL gather"
            # single-GPU answer for the whole array on this rank's device
            xi = x.cuda()
            ref_idx = torch.empty((n, 4, 4, 4), dtype=torch.uint8, device="cuda")
            ref_vox = torch.empty((n, 1, 8, 8, 8), dtype=torch.float32, device="cuda")
            codec.encode_device(xi, n, ref_idx, sp)
            codec.decode_device(ref_idx, n, ref_vox, sp)
            torch.cuda.synchronize()
            q.put(bool(torch.equal(full_idx, ref_idx) and torch.equal(full_vox, ref_vox)))
        codec.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_two_gpu_sharded_roundtrip_equals_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gpu_worker, args=(r, 2, port, 1001, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
