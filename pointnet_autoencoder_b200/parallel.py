"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the box, gloo
in CPU tests).

The reconstruction-loss ops are a loop over the batch index with no cross-element data flow
(tf_nndistance_g.cu:8,133; tf_approxmatch_g.cu:13,187,231,271), so the path shards by contiguous
batch slices with NO data-path collective.  The only real exchange step is the data-parallel
gradient all-reduce of a training step (one flattened fp32 bucket per step).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(batch, rank, world):
    """Contiguous slice [lo, hi) of a batch of `batch` elements owned by `rank` (sizes differ by <= 1)."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(t, rank=None, world=None, dim=0):
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(t.shape[dim], rank, world)
    return t.narrow(dim, lo, hi - lo)


def all_gather_batch(local, batch, group=None):
    """Reassemble a batch-sharded result on every rank (parity checks only; not on the hot path)."""
    world = dist.get_world_size(group)
    sizes = [shard_bounds(batch, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)


class GradBucket:
    """One flat fp32 buffer viewing every parameter's gradient: a training step needs exactly one
    all-reduce (sum, then divide by the world size: the loss is a mean over the global batch)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off: off + p.numel()].view_as(p)     # gradients accumulate straight into the bucket
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce(self, group=None, async_op=False):
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if world == 1:
            return None
        self.flat.div_(world)
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
