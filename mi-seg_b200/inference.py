"""Sliding-window inference for the conditional nets, sharded across the GPUs of a box (SURVEY.md 8(e), 8(f) row 4;
BASELINE.json configs[4]).

The reference runs MONAI's `sliding_window_inference(inputs, roi_size, sw_batch_size, predictor, overlap,
modalities=...)` (predict_whs.py:72-100, networks/lightning_monai.py:86-93, test.py:153-159, tune.py:141-148) on one
GPU, and can only do it with `sw_batch_size=1`: the `modalities` keyword is forwarded unchanged to every predictor
call, so a batch of windows meets a modality tensor of the IMAGE batch size and the conditional norm refuses it
("Expected number of styles as batch size.", conditional_instance_norm.py:46-47).

This module restates the same window enumeration and constant-weight blending (MONAI 1.1.0
`monai/inferers/utils.py`: symmetric zero padding up to the roi, `scan_interval = int(roi * (1 - overlap))`,
window starts clamped to the volume, windows ordered first-dimension-slowest, output = sum / count) and adds what the
path needs on an 8-GPU box:

* windows are independent, so rank r takes windows r, r + world, ... (one process per GPU, no forward collective);
  each rank accumulates weighted logits and counts locally and ONE all-reduce (sum) of the two maps ends the volume;
* `modalities` is expanded to the window batch (every window inherits the modality of the image it was cut from),
  so `sw_batch_size > 1` works for the conditional models.

Host logic only: the predictor is any callable `predictor(windows, modalities=...) -> logits`; the norms inside it
run on the sm_100a kernels when the model was built through `install()`.
"""
from __future__ import annotations

import itertools
import math
from typing import Callable, List, Optional, Sequence, Tuple, Union

import torch
import torch.distributed as dist
import torch.nn.functional as F

__all__ = ["window_slices", "sliding_window_inference"]


def _tuple(v: Union[int, Sequence[int]], n: int) -> Tuple[int, ...]:
    if isinstance(v, int):
        return (v,) * n
    v = tuple(int(a) for a in v)
    if len(v) != n:
        raise ValueError(f"roi_size must have {n} entries, got {len(v)}")
    return v


def window_slices(image_size: Sequence[int], roi_size: Sequence[int], overlap: float) -> List[Tuple[slice, ...]]:
    """Window slices of a (padded) volume in MONAI's order (`dense_patch_slices`, first dimension slowest)."""
    if not 0.0 <= overlap < 1.0:
        raise ValueError("overlap must be >= 0 and < 1.")
    starts_per_dim = []
    for size, roi in zip(image_size, roi_size):
        interval = roi if roi == size else max(int(roi * (1 - overlap)), 1)
        num = int(math.ceil(float(size - roi) / interval)) + 1
        starts = []
        for i in range(num):
            st = i * interval
            st -= max(st + roi - size, 0)
            starts.append(st)
        starts_per_dim.append(starts)
    return [tuple(slice(s, s + r) for s, r in zip(st, roi_size)) for st in itertools.product(*starts_per_dim)]


def sliding_window_inference(inputs: torch.Tensor, roi_size: Union[int, Sequence[int]], sw_batch_size: int,
                             predictor: Callable[..., torch.Tensor], overlap: float = 0.25,
                             modalities: Optional[Union[torch.Tensor, Sequence[int]]] = None, cval: float = 0.0,
                             group=None, shard: bool = True) -> torch.Tensor:
    """Drop-in for the way MI-Seg calls MONAI's `sliding_window_inference` (same positional arguments and
    `overlap` / `modalities` keywords, mode="constant").  With an initialised process group and `shard=True`
    the windows are dealt round-robin to the ranks and every rank returns the full blended volume."""
    nd = inputs.dim() - 2
    if nd < 1:
        raise ValueError("inputs must be [B, C, *spatial]")
    batch = inputs.shape[0]
    orig = tuple(inputs.shape[2:])
    roi = tuple(o if r is None or r <= 0 else r for r, o in zip(_tuple(roi_size, nd), orig))  # fall_back_tuple
    size = tuple(max(o, r) for o, r in zip(orig, roi))
    pad: List[int] = []
    for k in range(nd - 1, -1, -1):  # F.pad wants the last dimension first
        diff = max(roi[k] - orig[k], 0)
        pad.extend([diff // 2, diff - diff // 2])
    if any(pad):
        inputs = F.pad(inputs, pad, mode="constant", value=cval)
    slices = window_slices(size, roi, overlap)
    total = len(slices) * batch  # window w of image b has index b * len(slices) + w  (MONAI's order)

    world, rank = 1, 0
    if shard and dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = list(range(rank, total, world))

    mods = None
    mods_host: Optional[List[int]] = None
    if modalities is not None:
        mods = modalities if isinstance(modalities, torch.Tensor) else torch.as_tensor(list(modalities))
        mods = mods.reshape(-1)
        if mods.numel() != batch:
            raise ValueError("Expected number of styles as batch size.")  # the norm's own message
        if not (mods.is_cuda and torch.cuda.is_current_stream_capturing()):
            mods_host = [int(v) for v in mods.tolist()]  # ONE read-back per volume (the fast norms would otherwise read
                                                         # back every window batch's modality tensor, norms.py)

    out_sum: Optional[torch.Tensor] = None
    count = torch.zeros((batch, 1) + size, dtype=torch.float32, device=inputs.device)
    for g0 in range(0, len(mine), sw_batch_size):
        idx = mine[g0:g0 + sw_batch_size]
        where = [(i // len(slices), slices[i % len(slices)]) for i in idx]
        windows = torch.cat([inputs[(slice(b, b + 1), slice(None)) + sl] for b, sl in where])
        if mods is not None:
            if mods_host is not None:  # one modality per WINDOW; the host copy rides along (norms._cuda_styles_on_host)
                hv = [mods_host[b] for b, _ in where]
                wmods = torch.tensor(hv, dtype=mods.dtype, device=mods.device)
                if wmods.is_cuda:
                    wmods._micn_host = (wmods._version, hv)
            else:
                wmods = mods[torch.as_tensor([b for b, _ in where], device=mods.device)]
            prob = predictor(windows, modalities=wmods)
        else:
            prob = predictor(windows)
        if out_sum is None:
            out_sum = torch.zeros((batch, prob.shape[1]) + size, dtype=torch.float32, device=inputs.device)
        for k, (b, sl) in enumerate(where):
            out_sum[(b, slice(None)) + sl] += prob[k].to(torch.float32)  # importance map of mode="constant" is 1
            count[(b, 0) + sl] += 1.0
    if world > 1:
        # a rank with no window still has to know the channel count: agree on it, then reduce the two maps
        ch = torch.tensor([0 if out_sum is None else out_sum.shape[1]], device=inputs.device)
        dist.all_reduce(ch, op=dist.ReduceOp.MAX, group=group)
        if out_sum is None:
            out_sum = torch.zeros((batch, int(ch.item())) + size, dtype=torch.float32, device=inputs.device)
        dist.all_reduce(out_sum, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(count, op=dist.ReduceOp.SUM, group=group)
    if out_sum is None:
        raise RuntimeError("sliding_window_inference: no window was evaluated")
    out = out_sum / count
    if any(pad):  # crop the symmetric padding away again
        crop = [slice(None), slice(None)]
        for k in range(nd):
            lo = pad[2 * (nd - 1 - k)]
            crop.append(slice(lo, lo + orig[k]))
        out = out[tuple(crop)]
    return out.to(inputs.dtype) if inputs.dtype.is_floating_point else out
