"""Leaf-range sharding across the GPUs of one box, one process per GPU.

Leaves are independent (GroupNorm and attention are per leaf), so rank r of G simply owns the contiguous
range [r*N/G, (r+1)*N/G): origins and indices stay in file order and reassembly is a concatenation.  The
only exchange step on the path is the gather of decoded blocks (2048 B/leaf) — or of indices (64 B/leaf) on
the encode side — to the rank that rebuilds the grid.  Two forms:

* `PeerGather` (GPUs): the reassembly rank owns the gathered buffer and shares it by CUDA IPC; every rank's decode
  kernel stores its blocks straight into its slice of that buffer over NVLink — compute and gather are one kernel,
  no staging copy, no separate collective.  torch.distributed only carries the 64-byte handle and the barrier.
* `gather_blocks`: a plain torch.distributed gather (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def leaf_range(rank: int, world: int, n_leaves: int) -> Tuple[int, int]:
    """Contiguous, balanced: sizes differ by at most one leaf and concatenate to [0, n)."""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return (rank * n_leaves) // world, ((rank + 1) * n_leaves) // world


def shard_sizes(world: int, n_leaves: int) -> List[int]:
    return [leaf_range(r, world, n_leaves)[1] - leaf_range(r, world, n_leaves)[0] for r in range(world)]


def gather_blocks(local: torch.Tensor, n_leaves: int, dst: int = 0, group=None,
                  out: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """Gathers per-rank shards (leading dim = this rank's leaf count) into file order on `dst`.

    Ragged shards are handled by padding to the largest shard for the collective and slicing on arrival,
    so a single gather call moves everything.  Returns the [n_leaves, ...] tensor on dst, None elsewhere.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(world, n_leaves)
    if local.shape[0] != sizes[rank]:
        raise ValueError("rank %d holds %d leaves, expected %d" % (rank, local.shape[0], sizes[rank]))
    biggest = max(sizes)
    if min(sizes) == biggest:  # equal shards: receive straight into slices of the result, no staging copy
        if rank == dst and out is None:
            out = local.new_empty((n_leaves,) + tuple(local.shape[1:]))
        bufs = [out[r * biggest:(r + 1) * biggest] for r in range(world)] if rank == dst else None
        dist.gather(local.contiguous(), bufs, dst=dst, group=group)
        return out if rank == dst else None
    send = local
    if local.shape[0] != biggest:
        send = local.new_zeros((biggest,) + tuple(local.shape[1:]))
        send[: local.shape[0]] = local
    bufs = None
    if rank == dst:
        bufs = [local.new_empty((biggest,) + tuple(local.shape[1:])) for _ in range(world)]
    dist.gather(send.contiguous(), bufs, dst=dst, group=group)
    if rank != dst:
        return None
    if out is None:
        out = local.new_empty((n_leaves,) + tuple(local.shape[1:]))
    for r in range(world):
        lo, hi = leaf_range(r, world, n_leaves)
        out[lo:hi] = bufs[r][: hi - lo]
    return out


class _DevArray:
    """Minimal __cuda_array_interface__ carrier so torch can view a raw device allocation."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2}


class PeerGather:
    """Decode-and-gather in one kernel: rank `dst` creates the [n_leaves, *elem_shape] float32 buffer on its GPU,
    the other ranks map it (CUDA IPC, NVLink peer access) and decode straight into their leaf range.

        pg = PeerGather(codec, n_leaves, dst=0)
        codec.decode_device(idx, hi - lo, pg.my_slice_ptr, stream)     # stores land in dst's HBM
        full = pg.finish(stream)                                      # barrier; [n_leaves,1,8,8,8] on dst, None elsewhere
    """

    def __init__(self, codec, n_leaves: int, dst: int = 0, elem_floats: int = 512, group=None):
        self.codec, self.group, self.dst = codec, group, dst
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.n_leaves, self.elem_floats = n_leaves, elem_floats
        self.lo, self.hi = leaf_range(self.rank, self.world, n_leaves)
        nbytes = max(1, n_leaves * elem_floats * 4)
        box = [None]
        if self.rank == dst:
            self.base, handle = codec.peer_buffer_create(nbytes)
            box[0] = handle
        dist.broadcast_object_list(box, src=dst, group=group)
        self.opened = self.rank != dst
        if self.opened:
            self.base = codec.peer_buffer_open(box[0])
        self.my_slice_ptr = self.base + self.lo * elem_floats * 4
        self.full = None
        if self.rank == dst:
            self.full = torch.as_tensor(_DevArray(self.base, (n_leaves, elem_floats), "<f4"), device=torch.device("cuda", torch.cuda.current_device()))

    def finish(self, stream=None) -> Optional[torch.Tensor]:
        """Every rank's decode on `stream` has completed and is visible on dst when this returns."""
        (stream or torch.cuda.current_stream()).synchronize()
        dist.barrier(group=self.group)
        return self.full

    def close(self):
        if getattr(self, "base", None):
            self.full = None
            dist.barrier(group=self.group)   # nobody still writes into a buffer that is about to go away
            self.codec.peer_buffer_close(self.base, self.opened)
            self.base = None
