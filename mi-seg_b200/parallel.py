"""Data-parallel plumbing for the instance_cond path.

The norm is per-sample, so batch / patch / window sharding needs NO forward collective; the only
exchange is the all-reduce of the per-style parameter gradients d(gamma[s]), d(beta[s]) - what DDP's
backward does for MI-Seg (tune.py:103-109).  One process per GPU, NCCL over NVLink on the GPU box;
the same code runs on `gloo` for the CPU tests.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of `total` independent units (samples, patches, sliding windows):
    the first total % world ranks get one extra unit."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def style_parameters(module: torch.nn.Module) -> List[torch.nn.Parameter]:
    """All norms.{s}.weight / norms.{s}.bias of the conditional norms under `module`, in module order."""
    out = []
    for m in module.modules():
        inner = getattr(m, "norms", None)
        if isinstance(inner, torch.nn.ModuleList) and hasattr(m, "num_styles"):
            for n in inner:
                out.extend([n.weight, n.bias])
    return out


def allreduce_style_grads(params: Iterable[torch.nn.Parameter], group=None, average: bool = True,
                          bucket: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One flat fp32 bucket for every per-style gradient: a style absent on this rank (grad None, as in
    the reference) contributes zeros plus a 0 presence flag; after the SUM all-reduce a parameter whose
    style was absent on EVERY rank keeps grad None, the others get the (averaged) sum.  Returns the bucket."""
    params = list(params)
    if not params:
        return torch.empty(0)
    dev = params[0].device
    sizes = [p.numel() for p in params]
    total = sum(sizes) + len(params)
    if bucket is None or bucket.numel() != total or bucket.device != dev:
        bucket = torch.empty(total, dtype=torch.float32, device=dev)
    off = 0
    flags = bucket[sum(sizes):]
    for i, (p, n) in enumerate(zip(params, sizes)):
        if p.grad is None:
            bucket[off:off + n].zero_()
            flags[i] = 0.0
        else:
            bucket[off:off + n].copy_(p.grad.reshape(-1))
            flags[i] = 1.0
        off += n
    world = 1
    if dist.is_available() and dist.is_initialized():
        world = dist.get_world_size(group)
        if world > 1:
            dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
    host_flags = flags.tolist()  # one small D2H per step, outside the kernels' critical path
    off = 0
    scale = 1.0 / world if average else 1.0
    for p, n, f in zip(params, sizes, host_flags):
        if f > 0:
            g = bucket[off:off + n].reshape(p.shape) * scale
            p.grad = g.to(p.dtype)
        else:
            p.grad = None
        off += n
    return bucket
