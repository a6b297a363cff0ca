"""Data-parallel plumbing for the instance_cond path.

The norm is per-sample, so batch / patch / window sharding needs NO forward collective; the only
exchange is the all-reduce of the per-style parameter gradients d(gamma[s]), d(beta[s]) - what DDP's
backward does for MI-Seg (tune.py:103-109).  One process per GPU, NCCL over NVLink on the GPU box;
the same code runs on `gloo` for the CPU tests.
"""
from __future__ import annotations

import os
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


class PeerExchange:
    """Exchange buffers for `micn_bwd_allreduce` (include/micn.h): the all-reduce of d(gamma) / d(beta) fused into the
    backward kernel over NVLink peer memory, one process per GPU of one box.

    Every rank allocates one zero-filled buffer, shares it through CUDA IPC (the handle travels with
    `dist.all_gather_object`), and maps every peer's buffer into its own address space.  `ptrs` is the ctypes array of
    `world` device pointers the C ABI takes (entry `rank` = the local buffer).  All ranks must then make the same sequence
    of `micn_bwd_allreduce` calls with this exchange (the record tag is a launch counter kept in the buffer)."""

    def __init__(self, channels: int, num_styles: int, device: torch.device, group=None):
        import ctypes

        from . import _lib

        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerExchange needs an initialised process group (one process per GPU)")
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > _lib.MAX_PEERS:
            raise ValueError(f"at most {_lib.MAX_PEERS} GPUs can take part in the fused exchange")
        nbytes = int(_lib.lib().micn_peer_buffer_bytes(channels, num_styles, self.world))
        self.channels, self.num_styles = channels, num_styles
        self._peers = []  # keeps mapped peer storages alive (CUDA IPC route)
        self.how = None
        how = os.environ.get("MICN_PEER_ALLOC", "auto")
        ptrs = None
        if how in ("auto", "symm"):
            try:  # symmetric memory: one allocation mapped into every rank with access for all peers (cuMemSetAccess)
                import torch.distributed._symmetric_memory as symm_mem

                with torch.cuda.device(device):
                    self.local = symm_mem.empty(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
                    self._hdl = symm_mem.rendezvous(self.local, dist.group.WORLD if group is None else group)
                self.local.zero_()
                torch.cuda.synchronize(device)
                ptrs = [int(p) for p in self._hdl.buffer_ptrs]
                self.how = "symmetric_memory"
            except Exception as e:  # noqa: BLE001
                if how == "symm":
                    raise
                self._symm_error = repr(e)[:300]
        if ptrs is None:
            ptrs = self._map_with_cuda_ipc(nbytes, device, group)
            self.how = "cuda_ipc"
        assert len(ptrs) == self.world and ptrs[self.rank] == self.local.data_ptr()
        self.ptrs = (ctypes.c_void_p * self.world)(*ptrs)
        dist.barrier(group=group)  # nobody writes into a buffer that is not mapped (and zeroed) everywhere yet

    def _map_with_cuda_ipc(self, nbytes, device, group):
        """Every rank exports its buffer with cudaIpcGetMemHandle (torch's storage sharing), the handles travel with
        all_gather_object, and every peer buffer is opened in this process; a first peer-to-peer copy makes PyTorch
        enable peer access device -> peer for the kernels."""
        self.local = torch.zeros(max(nbytes, 2 << 20), dtype=torch.uint8, device=device)
        torch.cuda.synchronize(device)
        handle = self.local.untyped_storage()._share_cuda_()
        handles = [None] * self.world
        dist.all_gather_object(handles, handle, group=group)
        ptrs = []
        for r, h in enumerate(handles):
            if r == self.rank:
                ptrs.append(self.local.data_ptr())
                continue
            if not torch.cuda.can_device_access_peer(device.index, int(h[0])):
                raise RuntimeError(f"PeerExchange: GPU {device.index} cannot access the memory of rank {r}'s GPU")
            with torch.cuda.device(device):
                st = torch.UntypedStorage._new_shared_cuda(*h)
            t = torch.empty(0, dtype=torch.uint8, device=st.device).set_(st)
            scratch = torch.empty(16, dtype=torch.uint8, device=device)
            scratch.copy_(t[:16])
            self._peers.append(t)
            ptrs.append(t.data_ptr())
        torch.cuda.synchronize(device)
        return ptrs

