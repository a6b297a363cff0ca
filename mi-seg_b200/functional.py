"""Autograd entry point of the B200 instance_cond kernels.

`instance_cond(x, styles, weights, biases, ...)` computes, per sample n and channel c,

    y[n, c] = (x[n, c] - mean) * rstd * weights[styles[n]][c] + biases[styles[n]][c]

(+ optional fused epilogue) by calling `micn_fwd` / `micn_bwd` of libmicn.so on the current CUDA
stream.  It replaces the per-sample loop + torch.stack of the reference
(/root/reference/networks/norms/conditional_instance_norm.py:59-60) and the autograd graph behind
it; the epilogues replace dynunet_block.py:107-111 / :113-125.

No CPU path: CPU tensors raise.  All statistics / parameter math is fp32; I/O dtype = x.dtype.
"""
from __future__ import annotations

import ctypes
import importlib.util
import os
from typing import List, Optional, Sequence

import torch

from . import _lib

_DTYPES = {torch.float32: _lib.MICN_F32, torch.bfloat16: _lib.MICN_BF16, torch.float16: _lib.MICN_F16}
_EPILOGUES = {"none": _lib.EPI_NONE, "lrelu": _lib.EPI_LRELU, "add_lrelu": _lib.EPI_ADD_LRELU}

_workspaces = {}
_ptr_arrays = {}

# ---- host-side binding: the C++ autograd functions of csrc/micn_torch.cpp when _micn_torch.so was built, else the
#      torch.autograd.Function / ctypes classes below.  Both issue the same C-ABI calls into libmicn.so.
_binding_choice = os.environ.get("MICN_BINDING", "auto")  # auto | cpp | ctypes
_ext_module = None
_ext_tried = False


def set_binding(which: str) -> None:
    """"auto" (C++ binding when built), "cpp" (require it) or "ctypes" (the Python path)."""
    global _binding_choice
    if which not in ("auto", "cpp", "ctypes"):
        raise ValueError("binding must be 'auto', 'cpp' or 'ctypes'")
    _binding_choice = which


def _ext():
    global _ext_module, _ext_tried
    if _binding_choice == "ctypes":
        return None
    if not _ext_tried:
        _ext_tried = True
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_micn_torch.so")
        if os.path.exists(path):
            _lib.lib()  # libmicn.so first (the binding links against it)
            spec = importlib.util.spec_from_file_location("_micn_torch", path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            if mod.micn_version() != _lib.lib().micn_version():
                raise _lib.MicnError("_micn_torch.so and libmicn.so were built from different sources: rebuild both")
            _ext_module = mod
    if _ext_module is None and _binding_choice == "cpp":
        raise _lib.MicnError("the C++ binding mi-seg_b200/_micn_torch.so is not built (make -C mi-seg_b200/csrc)")
    return _ext_module


def binding_in_use() -> str:
    return "cpp" if _ext() is not None else "ctypes"


def _present_mask(present) -> int:
    if present is None:
        return -1
    mask = 0
    for i, p in enumerate(present):
        if p:
            mask |= 1 << i
    return mask


_ws_need = {}


def _workspace(device: torch.device, stream: int, n: int, c: int, m: int, dtype_code: int, num_styles: int) -> torch.Tensor:
    """Zero-filled device workspace, one per (device, stream), grown on demand (micn.h: must be
    zero-filled once when allocated; the kernels leave it reusable)."""
    shape_key = (n, c, m, dtype_code, num_styles)
    need = _ws_need.get(shape_key)
    if need is None:
        need = int(_lib.lib().micn_workspace_bytes(n, c, m, dtype_code, num_styles))
        if len(_ws_need) > 4096:
            _ws_need.clear()
        _ws_need[shape_key] = need
    key = (device.index, stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(max(need, 1 << 16), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


class _on_device:
    """`with torch.cuda.device(dev)` only when `dev` is not already current (the context manager costs ~10 us)."""

    __slots__ = ("ctx",)

    def __init__(self, dev: torch.device):
        self.ctx = None if dev.index == torch.cuda.current_device() else torch.cuda.device(dev)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def _raw_stream(dev: torch.device) -> int:
    """cudaStream_t of torch's current stream on `dev` (the raw getter: ~0.3 us instead of ~8 us for the Stream object)."""
    return torch._C._cuda_getCurrentRawStream(dev.index)


def reset_workspaces() -> None:
    """Drop every cached workspace (call after a CUDA error so stale control words cannot survive)."""
    _workspaces.clear()
    _cl_workspaces.clear()


def _ptr_array(tensors: Sequence[torch.Tensor]):
    """ctypes array of the tensors' device pointers (cached: parameters keep their storage between steps)."""
    key = tuple(t.data_ptr() for t in tensors)
    arr = _ptr_arrays.get(key)
    if arr is None:
        arr = (ctypes.c_void_p * len(key))(*key)
        if len(_ptr_arrays) > 4096:
            _ptr_arrays.clear()
        _ptr_arrays[key] = arr
    return arr


def _dense_spatial(x: torch.Tensor) -> bool:
    """True when dims 2.. form one dense block (so a slab is M consecutive elements)."""
    expect = 1
    for size, stride in zip(reversed(x.shape[2:]), reversed(x.stride()[2:])):
        if size != 1 and stride != expect:
            return False
        expect *= size
    return True


def _as_ncm(x: torch.Tensor):
    """Return (tensor, stride_n, stride_c) with dense slabs; copies only when the layout forces it."""
    if not _dense_spatial(x):
        x = x.contiguous()
    n, c = x.shape[0], x.shape[1]
    m = 1
    for s in x.shape[2:]:
        m *= s
    sn = x.stride(0) if n > 1 else c * m
    sc = x.stride(1) if c > 1 else m
    if sn < 0 or sc < 0 or (n > 1 and sn == 0) or (c > 1 and sc == 0):  # expanded / flipped views
        x = x.contiguous()
        sn, sc = c * m, m
    return x, sn, sc, m


def _f32_params(ts: Sequence[torch.Tensor], device) -> List[torch.Tensor]:
    out = []
    for t in ts:
        if t.device != device:
            raise RuntimeError(f"instance_cond: parameter on {t.device}, input on {device}")
        if t.dtype != torch.float32 or not t.is_contiguous():
            t = t.detach().float().contiguous()
        out.append(t)  # (only the address is used, and autograd does not record inside Function.forward)
    return out


def _slope_tensor(slope, dev) -> Optional[torch.Tensor]:
    """The PReLU weight as a one-element fp32 CUDA tensor, or None for a python float (LeakyReLU)."""
    if not isinstance(slope, torch.Tensor):
        return None
    if slope.numel() != 1:
        raise ValueError("instance_cond: only a single-parameter PReLU slope is supported (nn.PReLU(num_parameters=1))")
    if slope.device != dev:
        raise RuntimeError(f"instance_cond: slope on {slope.device}, input on {dev}")
    t = slope.detach().reshape(1)
    return t if t.dtype == torch.float32 else t.float()


class _InstanceCondFn(torch.autograd.Function):
    """forward(x, styles_dev, residual, eps, epilogue, slope, present, S, *weights, *biases); `slope` is a float
    (LeakyReLU) or the one-element weight tensor of an nn.PReLU (read on the device, gradient returned)."""

    @staticmethod
    def forward(ctx, x, styles_dev, residual, eps, epilogue, slope, present, num_styles, *params):
        if not x.is_cuda:
            raise RuntimeError("instance_cond (mi-seg_b200) runs on CUDA tensors only: there is no CPU fallback")
        if x.dtype not in _DTYPES:
            raise TypeError(f"instance_cond: unsupported dtype {x.dtype} (float32, bfloat16, float16)")
        lib = _lib.lib()
        dev = x.device
        affine = len(params) > 0
        weights = _f32_params(params[:num_styles], dev) if affine else []
        biases = _f32_params(params[num_styles:], dev) if affine else []
        xs, sn, sc, m = _as_ncm(x)
        n, c = xs.shape[0], xs.shape[1]
        if affine and any(w.numel() != c for w in weights + biases):
            raise ValueError("instance_cond: parameter length does not match the channel count")
        y = torch.empty(xs.shape, dtype=xs.dtype, device=dev)  # fresh contiguous NC* (as torch.stack gives)
        stats = torch.empty(2, n * c, dtype=torch.float32, device=dev)  # [0] mean, [1] rstd
        mean_p = stats.data_ptr()
        rstd_p = mean_p + 4 * n * c
        res = None
        if epilogue == _lib.EPI_ADD_LRELU:
            if residual is None or residual.shape != xs.shape:
                raise ValueError("instance_cond: add_lrelu needs a residual of the input's shape")
            # Under autocast the block input can be fp32 while conv2's output is 16-bit (layer_norm is on autocast's
            # fp32 list, so C-Swin-UNETR's hidden states reach encoder2/3/4/10 in fp32).  The reference's in-place
            # `out += residual` (dynunet_block.py:123) rounds the sum to out's dtype; so does this.
            res = residual.contiguous() if residual.dtype == xs.dtype else residual.to(xs.dtype).contiguous()
        stream = _raw_stream(dev)
        ws = _workspace(dev, stream, n, c, m, _DTYPES[xs.dtype], num_styles)
        with _on_device(dev):
            gp = _ptr_array(weights) if affine else None
            bp = _ptr_array(biases) if affine else None
            slope_t = _slope_tensor(slope, dev)
            if slope_t is not None and epilogue != _lib.EPI_LRELU:
                # (with a residual the sign of the pre-activation cannot be recovered from the output for a slope <= 0,
                # and MI-Seg has no prelu(norm(x) + r): C-UNet adds its residual AFTER the activation, convolutions.py:329)
                raise ValueError("instance_cond: a PReLU slope (tensor) needs the 'lrelu' epilogue")
            if slope_t is None:
                rc = lib.micn_fwd(xs.data_ptr(), y.data_ptr(), res.data_ptr() if res is not None else None, gp, bp,
                                  num_styles, styles_dev.data_ptr() if styles_dev is not None else None,
                                  mean_p, rstd_p, n, c, m, sn, sc, _DTYPES[xs.dtype], epilogue,
                                  float(slope), float(eps), ws.data_ptr(), ws.numel(), stream)
            else:
                rc = lib.micn_fwd_prelu(xs.data_ptr(), y.data_ptr(), res.data_ptr() if res is not None else None, gp, bp,
                                        num_styles, styles_dev.data_ptr() if styles_dev is not None else None,
                                        mean_p, rstd_p, n, c, m, sn, sc, _DTYPES[xs.dtype], epilogue,
                                        slope_t.data_ptr(), float(eps), ws.data_ptr(), ws.numel(), stream)
        _lib.check(rc, "micn_fwd")
        keep_y = epilogue == _lib.EPI_ADD_LRELU  # (its LeakyReLU mask, and with a PReLU slope its gradient, come from y)
        ctx.save_for_backward(xs, styles_dev, stats, y if keep_y else None, slope_t, *weights, *biases)
        ctx.meta = (n, c, m, sn, sc, epilogue, None if slope_t is not None else float(slope), num_styles, affine, present,
                    residual.dtype if (residual is not None and epilogue == _lib.EPI_ADD_LRELU) else None)
        ctx.slope_shape = tuple(slope.shape) if slope_t is not None else None
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        n, c, m, sn, sc, epilogue, slope, num_styles, affine, present, res_dtype = ctx.meta
        has_res = res_dtype is not None
        xs, styles_dev, stats, act_out, slope_t, *params = ctx.saved_tensors
        weights, biases = params[:num_styles], params[num_styles:]
        lib = _lib.lib()
        dev = xs.device
        dy = dy.contiguous()
        if dy.dtype != xs.dtype:
            dy = dy.to(xs.dtype)
        dx = torch.empty(dy.shape, dtype=xs.dtype, device=dev)
        dres = torch.empty_like(dx) if has_res else None
        need_param_grads = affine and any(ctx.needs_input_grad[8:])
        # [dgamma_0 .. dgamma_{S-1}, dbeta_0 .. dbeta_{S-1}], each [C]
        pgrads = torch.empty((2 * num_styles, c), dtype=torch.float32, device=dev) if need_param_grads else None
        dgamma_p = pgrads.data_ptr() if need_param_grads else None
        dbeta_p = dgamma_p + 4 * num_styles * c if need_param_grads else None
        mean_p = stats.data_ptr()
        rstd_p = mean_p + 4 * n * c
        stream = _raw_stream(dev)
        ws = _workspace(dev, stream, n, c, m, _DTYPES[xs.dtype], num_styles)
        with _on_device(dev):
            gp = _ptr_array(weights) if affine else None
            bp = _ptr_array(biases) if affine else None
            act_ptr = act_out.data_ptr() if (act_out is not None and epilogue == _lib.EPI_ADD_LRELU) else None
            common = (dy.data_ptr(), xs.data_ptr(), act_ptr, gp, bp, num_styles,
                      styles_dev.data_ptr() if styles_dev is not None else None, mean_p, rstd_p,
                      dx.data_ptr(), dres.data_ptr() if dres is not None else None, dgamma_p, dbeta_p,
                      n, c, m, sn, sc, _DTYPES[xs.dtype], epilogue)
            want_ds = slope_t is not None and ctx.needs_input_grad[5]
            ds_part = None
            if slope_t is None:
                rc = lib.micn_bwd(*common, slope, ws.data_ptr(), ws.numel(), stream)
            else:
                if want_ds and epilogue == _lib.EPI_LRELU:  # the kernels accumulate the slope gradient in the same pass
                    ds_part = torch.zeros(max(n * c, 1024), dtype=torch.float32, device=dev)
                rc = lib.micn_bwd_prelu(*common, slope_t.data_ptr(), ds_part.data_ptr() if ds_part is not None else None,
                                        ws.data_ptr(), ws.numel(), stream)
        _lib.check(rc, "micn_bwd")
        dslope = None
        if want_ds and ds_part is not None:
            dslope = ds_part.sum().reshape(ctx.slope_shape)
        if dres is not None and res_dtype != dres.dtype:
            dres = dres.to(res_dtype)
        grads: List[Optional[torch.Tensor]] = [dx, None, dres, None, None, dslope, None, None]
        if affine:
            if pgrads is None:
                grads.extend([None] * (2 * num_styles))
            elif present is None:
                grads.extend(pgrads.unbind(0))
            else:  # styles absent from the batch keep .grad None like the reference (when known on the host)
                grads.extend(g if present[i % num_styles] else None for i, g in enumerate(pgrads.unbind(0)))
        return tuple(grads)


# ------------------------------------------------------------------------------------------------ dual-norm epilogue
_dual_ok = {}


def dual_supported(n: int, c: int, m: int, dtype: torch.dtype, backward: bool) -> bool:
    """Whether micn_fwd_dual / micn_bwd_dual take a problem of this shape (cached per shape)."""
    key = (n, c, m, dtype, backward, torch.cuda.current_device(), _lib.option_generation)
    ok = _dual_ok.get(key)
    if ok is None:
        ok = bool(_lib.lib().micn_dual_supported(n, c, m, _DTYPES[dtype], 1 if backward else 0))
        if len(_dual_ok) > 4096:
            _dual_ok.clear()
        _dual_ok[key] = ok
    return ok


class _DualNormFn(torch.autograd.Function):
    """forward(a, b, styles_dev, eps, slope, present, S, *wa, *ba, *wb, *bb) = lrelu(norm_a(a) + norm_b(b)): the
    downsample branch of UnetResBlock (dynunet_block.py:113-125 with conv3 / norm3) in one kernel per direction."""

    @staticmethod
    def forward(ctx, a, b, styles_dev, eps, slope, present, num_styles, *params):
        lib = _lib.lib()
        dev = a.device
        affine = len(params) > 0
        S = num_styles
        ps = _f32_params(params, dev) if affine else []
        n, c = a.shape[0], a.shape[1]
        m = a.numel() // max(n * c, 1)
        if affine and any(w.numel() != c for w in ps):
            raise ValueError("instance_cond: parameter length does not match the channel count")
        y = torch.empty_like(a)
        stats = torch.empty(4, n * c, dtype=torch.float32, device=dev)  # mean_a, rstd_a, mean_b, rstd_b
        sp = stats.data_ptr()
        q = 4 * n * c
        stream = _raw_stream(dev)
        ws = _workspace(dev, stream, n, c, m, _DTYPES[a.dtype], S)
        with _on_device(dev):
            arrs = [_ptr_array(ps[i * S:(i + 1) * S]) for i in range(4)] if affine else [None] * 4
            rc = lib.micn_fwd_dual(a.data_ptr(), b.data_ptr(), y.data_ptr(), arrs[0], arrs[1], arrs[2], arrs[3], S,
                                   styles_dev.data_ptr() if styles_dev is not None else None, sp, sp + q, sp + 2 * q,
                                   sp + 3 * q, n, c, m, _DTYPES[a.dtype], float(slope), float(eps), ws.data_ptr(),
                                   ws.numel(), stream)
        _lib.check(rc, "micn_fwd_dual")
        ctx.save_for_backward(a, b, styles_dev, stats, *ps)
        ctx.meta = (n, c, m, float(slope), S, affine, present)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        n, c, m, slope, S, affine, present = ctx.meta
        a, b, styles_dev, stats, *ps = ctx.saved_tensors
        lib = _lib.lib()
        dev = a.device
        dy = dy.contiguous()
        if dy.dtype != a.dtype:
            dy = dy.to(a.dtype)
        da, db = torch.empty_like(a), torch.empty_like(b)
        need_pg = affine and any(ctx.needs_input_grad[7:])
        # rows: dgamma_a[S], dbeta_a[S], dgamma_b[S], dbeta_b[S]
        pg = torch.empty((4 * S, c), dtype=torch.float32, device=dev) if need_pg else None
        gp = pg.data_ptr() if need_pg else None
        q = 4 * S * c
        sp = stats.data_ptr()
        sq = 4 * n * c
        stream = _raw_stream(dev)
        ws = _workspace(dev, stream, n, c, m, _DTYPES[a.dtype], S)
        with _on_device(dev):
            arrs = [_ptr_array(ps[i * S:(i + 1) * S]) for i in range(4)] if affine else [None] * 4
            rc = lib.micn_bwd_dual(dy.data_ptr(), a.data_ptr(), b.data_ptr(), arrs[0], arrs[1], arrs[2], arrs[3], S,
                                   styles_dev.data_ptr() if styles_dev is not None else None, sp, sp + sq, sp + 2 * sq,
                                   sp + 3 * sq, da.data_ptr(), db.data_ptr(), gp, gp + q if need_pg else None,
                                   gp + 2 * q if need_pg else None, gp + 3 * q if need_pg else None, n, c, m,
                                   _DTYPES[a.dtype], slope, ws.data_ptr(), ws.numel(), stream)
        _lib.check(rc, "micn_bwd_dual")
        grads: List[Optional[torch.Tensor]] = [da, db, None, None, None, None, None]
        if affine:
            if pg is None:
                grads.extend([None] * (4 * S))
            elif present is None:
                grads.extend(pg.unbind(0))
            else:
                grads.extend(g if present[i % S] else None for i, g in enumerate(pg.unbind(0)))
        return tuple(grads)


def instance_cond_dual(a: torch.Tensor, b: torch.Tensor, styles_dev: Optional[torch.Tensor],
                       weights_a: Sequence[torch.Tensor], biases_a: Sequence[torch.Tensor],
                       weights_b: Sequence[torch.Tensor], biases_b: Sequence[torch.Tensor], eps: float = 1e-5,
                       slope: float = 0.01, present: Optional[Sequence[bool]] = None,
                       num_styles: Optional[int] = None) -> Optional[torch.Tensor]:
    """lrelu(norm_a(a) + norm_b(b)) in one pass, or None when the dual kernels do not take this problem (the caller
    then composes `instance_cond(b)` and `instance_cond(a, epilogue="add_lrelu", residual=...)`)."""
    if not (a.is_cuda and b.is_cuda) or a.dtype not in _DTYPES or a.dtype != b.dtype or a.shape != b.shape or a.dim() < 3:
        return None
    if not (len(weights_a) == len(biases_a) == len(weights_b) == len(biases_b)):
        return None
    s = len(weights_a) if len(weights_a) else (num_styles or 1)
    n, c = a.shape[0], a.shape[1]
    m = a.numel() // max(n * c, 1)
    if a.numel() == 0 or not a.is_contiguous() or not b.is_contiguous() or a.data_ptr() % 16 or b.data_ptr() % 16:
        return None
    need_bwd = torch.is_grad_enabled() and (a.requires_grad or b.requires_grad or any(
        t.requires_grad for t in list(weights_a) + list(biases_a) + list(weights_b) + list(biases_b)))
    with _on_device(a.device):
        if not dual_supported(n, c, m, a.dtype, False) or (need_bwd and not dual_supported(n, c, m, a.dtype, True)):
            return None
    if styles_dev is not None and (styles_dev.dim() != 1 or not styles_dev.is_contiguous()):
        styles_dev = styles_dev.reshape(-1).contiguous()
    params = list(weights_a) + list(biases_a) + list(weights_b) + list(biases_b)
    if present is not None and params:
        params = [t if present[i % s] else t.detach() for i, t in enumerate(params)]
    ext = _ext()
    if ext is not None:
        ws = _workspace(a.device, _raw_stream(a.device), n, c, m, _DTYPES[a.dtype], s)
        return ext.instance_cond_dual(a, b, styles_dev, ws, float(eps), float(slope), _present_mask(present), s, params)
    return _DualNormFn.apply(a, b, styles_dev, eps, slope, tuple(present) if present is not None else None, s, *params)


# ------------------------------------------------------------------------------------------------ channels-last
_cl_native = True
_cl_workspaces = {}


def set_channels_last_native(enabled: bool) -> None:
    """Channels-last (stride_C == 1) inputs are normalised in place of their layout by default (the output keeps
    the input's strides, as PyTorch's own memory-format propagation does).  `False` restores the reference's
    behaviour of returning a fresh NC*-contiguous tensor at the cost of two transposing copies."""
    global _cl_native
    _cl_native = bool(enabled)


def _channels_last_view(x: torch.Tensor) -> Optional[torch.Tensor]:
    """x as a dense [N, *spatial, C] tensor if it is laid out that way (and worth it), else None."""
    if not _cl_native or x.dim() < 3 or x.shape[1] < 2 or (x.shape[1] & 1) or x.stride(1) != 1 or x.shape[0] > 65535:
        return None
    if x.data_ptr() % (2 * x.element_size()):  # the kernels load channel pairs: an odd storage offset takes the copy route
        return None
    perm = [0] + list(range(2, x.dim())) + [1]
    v = x.permute(perm)
    return v if v.is_contiguous() else None


def _cl_workspace(device, stream: int, n: int, c: int, m: int) -> torch.Tensor:
    need = int(_lib.lib().micn_cl_workspace_bytes(n, c, m))
    key = (device.index, stream)
    ws = _cl_workspaces.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(need, dtype=torch.uint8, device=device)
        _cl_workspaces[key] = ws
    return ws


class _InstanceCondClFn(torch.autograd.Function):
    """Channels-last route: forward(x_cl [N, *spatial, C] dense, styles_dev, eps, present, S, *weights, *biases)
    -> y_cl in the same layout (micn_fwd_cl / micn_bwd_cl)."""

    @staticmethod
    def forward(ctx, x_cl, styles_dev, eps, present, num_styles, *params):
        lib = _lib.lib()
        dev = x_cl.device
        affine = len(params) > 0
        weights = _f32_params(params[:num_styles], dev) if affine else []
        biases = _f32_params(params[num_styles:], dev) if affine else []
        n, c = x_cl.shape[0], x_cl.shape[-1]
        m = x_cl.numel() // (n * c)
        if affine and any(w.numel() != c for w in weights + biases):
            raise ValueError("instance_cond: parameter length does not match the channel count")
        y = torch.empty_like(x_cl)
        stats = torch.empty(2, n * c, dtype=torch.float32, device=dev)
        stream = _raw_stream(dev)
        ws = _cl_workspace(dev, stream, n, c, m)
        mean_p = stats.data_ptr()
        with _on_device(dev):
            rc = lib.micn_fwd_cl(x_cl.data_ptr(), y.data_ptr(), _ptr_array(weights) if affine else None,
                                 _ptr_array(biases) if affine else None, num_styles,
                                 styles_dev.data_ptr() if styles_dev is not None else None, mean_p,
                                 mean_p + 4 * n * c, n, c, m, _DTYPES[x_cl.dtype], float(eps), ws.data_ptr(), ws.numel(),
                                 stream)
        _lib.check(rc, "micn_fwd_cl")
        ctx.save_for_backward(x_cl, styles_dev, stats, *weights, *biases)
        ctx.meta = (n, c, m, num_styles, affine, present)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        n, c, m, num_styles, affine, present = ctx.meta
        x_cl, styles_dev, stats, *params = ctx.saved_tensors
        weights, biases = params[:num_styles], params[num_styles:]
        lib = _lib.lib()
        dev = x_cl.device
        dy = dy.contiguous()
        if dy.dtype != x_cl.dtype:
            dy = dy.to(x_cl.dtype)
        dx = torch.empty_like(x_cl)
        need_param_grads = affine and any(ctx.needs_input_grad[5:])
        pgrads = torch.empty((2 * num_styles, c), dtype=torch.float32, device=dev) if need_param_grads else None
        dgamma_p = pgrads.data_ptr() if need_param_grads else None
        stream = _raw_stream(dev)
        ws = _cl_workspace(dev, stream, n, c, m)
        mean_p = stats.data_ptr()
        with _on_device(dev):
            rc = lib.micn_bwd_cl(dy.data_ptr(), x_cl.data_ptr(), _ptr_array(weights) if affine else None,
                                 _ptr_array(biases) if affine else None, num_styles,
                                 styles_dev.data_ptr() if styles_dev is not None else None, mean_p,
                                 mean_p + 4 * n * c, dx.data_ptr(), dgamma_p,
                                 dgamma_p + 4 * num_styles * c if need_param_grads else None, n, c, m, _DTYPES[x_cl.dtype],
                                 ws.data_ptr(), ws.numel(), stream)
        _lib.check(rc, "micn_bwd_cl")
        grads: List[Optional[torch.Tensor]] = [dx, None, None, None, None]
        if affine:
            if pgrads is None:
                grads.extend([None] * (2 * num_styles))
            elif present is None:
                grads.extend(pgrads.unbind(0))
            else:
                grads.extend(g if present[i % num_styles] else None for i, g in enumerate(pgrads.unbind(0)))
        return tuple(grads)


def instance_cond(x: torch.Tensor, styles_dev: Optional[torch.Tensor], weights: Sequence[torch.Tensor],
                  biases: Sequence[torch.Tensor], eps: float = 1e-5, epilogue: str = "none",
                  residual: Optional[torch.Tensor] = None, slope=0.01,
                  present: Optional[Sequence[bool]] = None, num_styles: Optional[int] = None) -> torch.Tensor:
    """Batched input [N, C, *spatial]; `styles_dev` an int64 CUDA tensor [N] (None = style 0 everywhere);
    `weights` / `biases` per-style lists of [C] tensors (norms[s].weight / .bias), or empty for a
    non-affine norm.  `epilogue`: "none" | "lrelu" (lrelu(norm(x))) | "add_lrelu" (lrelu(norm(x)+residual))."""
    if epilogue not in _EPILOGUES:
        raise ValueError(f"instance_cond: unknown epilogue {epilogue!r}")
    s = len(weights) if len(weights) else (num_styles or 1)
    if len(weights) != len(biases):
        raise ValueError("instance_cond: weights and biases must have one entry per style")
    if s > _lib.MAX_STYLES:
        raise ValueError(f"instance_cond: at most {_lib.MAX_STYLES} styles are supported")
    if x.dim() < 3:
        raise ValueError("instance_cond: expected [N, C, *spatial] input")
    if styles_dev is not None:
        if styles_dev.device != x.device or styles_dev.dtype != torch.int64 or styles_dev.numel() != x.shape[0]:
            raise ValueError("instance_cond: styles must be an int64 tensor [N] on the input's device")
        if styles_dev.dim() != 1 or not styles_dev.is_contiguous():
            styles_dev = styles_dev.reshape(-1).contiguous()
    if present is not None and len(weights):
        # Parameters of styles absent from the batch stay OUT of the autograd graph, as in the reference (whose loop never
        # calls their nn.InstanceNorm, conditional_instance_norm.py:60): `.grad` stays None, and DDP's
        # find_unused_parameters=True (tune.py:103-109) sees them as unused instead of waiting for a hook that never fires.
        weights = [w if p else w.detach() for w, p in zip(weights, present)]
        biases = [b if p else b.detach() for b, p in zip(biases, present)]
    ext = _ext() if (x.is_cuda and x.dtype in _DTYPES) else None
    if epilogue == "none" and x.is_cuda and x.dtype in _DTYPES:
        x_cl = _channels_last_view(x)
        if x_cl is not None:  # token-major input: reduce the strided columns in place, keep the layout
            if ext is not None:
                n, c = x_cl.shape[0], x_cl.shape[-1]
                ws = _cl_workspace(x.device, _raw_stream(x.device), n, c, x_cl.numel() // max(n * c, 1))
                y_cl = ext.instance_cond_cl(x_cl, styles_dev, ws, float(eps), _present_mask(present), s,
                                            list(weights) + list(biases))
            else:
                y_cl = _InstanceCondClFn.apply(x_cl, styles_dev, eps, tuple(present) if present is not None else None, s,
                                               *weights, *biases)
            return y_cl.permute([0, x.dim() - 1] + list(range(1, x.dim() - 1)))
    if ext is not None:
        n, c = x.shape[0], x.shape[1]
        m = x.numel() // max(n * c, 1)
        ws = _workspace(x.device, _raw_stream(x.device), n, c, m, _DTYPES[x.dtype], s)
        slope_t = None
        if isinstance(slope, torch.Tensor):  # nn.PReLU weight: stays in the autograd graph (its gradient comes back)
            _slope_tensor(slope, x.device)  # (validation only)
            slope_t = slope.reshape(1) if slope.dtype == torch.float32 else slope.float().reshape(1)
        return ext.instance_cond(x, styles_dev, residual, slope_t, ws, float(eps), _EPILOGUES[epilogue],
                                 0.0 if slope_t is not None else float(slope), _present_mask(present), s,
                                 list(weights) + list(biases))
    return _InstanceCondFn.apply(x, styles_dev, residual, eps, _EPILOGUES[epilogue], slope,
                                 tuple(present) if present is not None else None, s, *weights, *biases)
