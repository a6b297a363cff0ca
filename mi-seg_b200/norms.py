"""Drop-in `instance_cond` modules: same constructor, `forward(input, styles)` signature, validation
messages and `norms.{s}.weight/bias` state-dict layout as the reference classes
(/root/reference/networks/norms/conditional_instance_norm.py:11-107), with the per-sample Python
loop + torch.stack (:59-60) replaced by one call into the sm_100a kernels.

Two ways to use them:

* stand-alone: `FastConditionalInstanceNorm{1,2,3}d` below derive from a local mirror of the
  reference base class, so they work where MI-Seg is not importable (the GPU box, the tests);
* inside MI-Seg: `integration.install()` re-creates the three classes on top of MI-Seg's own
  `_ConditionalInstanceNorm` (every block gates on `isinstance(norm, _ConditionalInstanceNorm)`,
  SURVEY.md section 8b) and registers them under the factory key "instance_cond".
"""
from __future__ import annotations

import warnings
from typing import List, Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn
from torch import Tensor

from .functional import instance_cond, instance_cond_dual

__all__ = ["FastConditionalInstanceNorm1d", "FastConditionalInstanceNorm2d", "FastConditionalInstanceNorm3d",
           "FastForwardMixin", "make_dropin_classes", "FastInstanceNorm1d", "FastInstanceNorm2d", "FastInstanceNorm3d",
           "FastPlainForwardMixin", "fast_instance_norm", "set_sync_free_styles", "check_status", "forward_fused_dual"]

_STYLE_CACHE = {}
_STYLE_CACHE_MAX = 256
_sync_free_styles = False


def set_sync_free_styles(enabled: bool) -> None:
    """How a CUDA `styles` tensor is handled.

    Default (False): its values are read back ONCE per tensor object (one `.tolist()`, cached on the tensor and
    reused by every norm of the model that receives the same object - MI-Seg hands one `modality.to(device)` tensor
    to all of them, utils/trainer.py:31,35; the reference itself syncs B times per norm call,
    conditional_instance_norm.py:60).  With the values on the host the module behaves exactly like the reference:
    an out-of-range id raises IndexError, and styles absent from the batch keep `.grad is None` (so AdamW neither
    decays nor moves them).

    True: never sync (needed while capturing a CUDA graph, where it is also selected automatically).  Ids are then
    read by the kernels; an out-of-range id is clamped and flagged in the workspace status word (`check_status()`
    raises for it), and absent styles receive ZERO gradients instead of None."""
    global _sync_free_styles
    _sync_free_styles = bool(enabled)


def _cuda_styles_on_host(styles: Tensor) -> Optional[List[int]]:
    if _sync_free_styles or torch.cuda.is_current_stream_capturing():
        return None
    cached = getattr(styles, "_micn_host", None)
    if cached is not None and cached[0] == styles._version:
        return cached[1]
    vals = [int(v) for v in styles.reshape(-1).tolist()]  # the one device sync per modality tensor
    try:
        styles._micn_host = (styles._version, vals)
    except Exception:  # noqa: BLE001 - a tensor type without a __dict__: just do not cache
        pass
    return vals


def _host_styles(styles, num_styles: int) -> Optional[List[int]]:
    """Styles as a list of python ints (list / int / CPU tensor, or a CUDA tensor read back once - see
    set_sync_free_styles); None for CUDA tensors in sync-free mode.  Mirrors how the reference indexes its
    ModuleList (:57, :60): negative ids wrap, out of range -> IndexError, floats -> TypeError."""
    if isinstance(styles, Tensor):
        if styles.is_floating_point() or styles.is_complex() or styles.dtype == torch.bool:
            # ModuleList.__getitem__ -> operator.index(tensor) raises TypeError for non-integer tensors
            raise TypeError(f"only integer tensors of a single element can be converted to an index "
                            f"(got styles of dtype {styles.dtype})")
        if styles.is_cuda:
            vals = _cuda_styles_on_host(styles)
            if vals is None:
                return None
        else:
            vals = [int(v) for v in styles.reshape(-1).tolist()]
    elif isinstance(styles, int):
        vals = [styles]
    else:
        vals = []
        for v in styles:
            if isinstance(v, Tensor):
                if v.is_floating_point():
                    raise TypeError("only integer tensors of a single element can be converted to an index")
                v = int(v.item())
            elif isinstance(v, bool) or not isinstance(v, int):
                if hasattr(v, "__index__"):
                    v = v.__index__()
                else:
                    raise TypeError(f"list indices must be integers or slices, not {type(v).__name__}")
            vals.append(int(v))
    out = []
    for v in vals:
        if v < -num_styles or v >= num_styles:
            raise IndexError(f"index {v} is out of range")
        out.append(v + num_styles if v < 0 else v)
    return out


def _device_styles(host: Sequence[int], device: torch.device) -> Tensor:
    key = (tuple(host), device.index)
    t = _STYLE_CACHE.get(key)
    if t is None:
        if len(_STYLE_CACHE) >= _STYLE_CACHE_MAX:
            _STYLE_CACHE.clear()
        t = torch.tensor(list(host), dtype=torch.int64, device=device)
        _STYLE_CACHE[key] = t
    return t


def check_status(device=None) -> None:
    """Raise IndexError if a kernel met an out-of-range style id since the last check (sync-free mode only: with host-
    visible styles the id is validated before the launch).  One small device-to-host read per cached workspace;
    call it at a step / epoch boundary."""
    import ctypes

    from . import _lib, functional

    lib = _lib.lib()
    for (dev_index, stream), ws in list(functional._workspaces.items()) + list(functional._cl_workspaces.items()):
        if device is not None and torch.device(device).index not in (None, dev_index):
            continue
        status = ctypes.c_int(0)
        with torch.cuda.device(dev_index):
            _lib.check(lib.micn_read_status(ws.data_ptr(), stream, ctypes.byref(status)), "micn_read_status")
        if status.value & 1:
            raise IndexError("instance_cond: a style id on the device was out of range "
                             "(clamped by the kernel; the output of that call used the wrong gamma/beta)")


class FastForwardMixin:
    """forward() shared by the stand-alone and the MI-Seg-derived classes.  Relies on the host
    class for `_check_input_dim`, `_check_input_styles`, `_get_no_batch_dim`, `norms`, `num_styles`."""

    def _params(self) -> Tuple[List[Tensor], List[Tensor]]:
        # (straight from the registries: nn.Module.__getattr__ costs ~0.5 us per lookup, four to six per call here)
        mods = self._modules["norms"]._modules.values()
        return [n._parameters["weight"] for n in mods], [n._parameters["bias"] for n in mods]

    def forward(self, input: Tensor, styles: Union[List, Tensor, int]) -> Tensor:
        """y = IN_{styles[n]}(input[n]) for every sample: the reference's forward (:62-68)."""
        return self.forward_fused(input, styles, "none")

    def forward_fused(self, input: Tensor, styles: Union[List, Tensor, int], epilogue: str = "none",
                      residual: Optional[Tensor] = None, slope: float = 0.01) -> Tensor:
        """forward() with the activation / residual that follows the norm in MI-Seg's blocks fused into
        the same kernel: "lrelu" = lrelu(norm(x)) (dynunet_block.py:107-111, 188-202), "add_lrelu" =
        lrelu(norm(x) + residual) (:113-125).  Used by blocks.fuse_blocks()."""
        self._check_input_dim(input)
        self._check_input_styles(input, styles)
        unbatched = input.dim() == self._get_no_batch_dim()
        x = input.unsqueeze(0) if unbatched else input
        host = _host_styles(styles, self.num_styles)
        first = self._modules["norms"]._modules["0"]  # (norms[0] without three trips through __getattr__ / __getitem__)
        if x.shape[1] != first.num_features:
            # nn.InstanceNorm*d._check_input_dim's message (affine norms check the channel count)
            raise ValueError(f"expected input's size at dim=1 to match num_features "
                             f"({first.num_features}), but got: {x.shape[1]}.")
        m = 1
        for s in x.shape[2:]:
            m *= s
        if m <= 1 and x.numel() > 0:
            raise ValueError(f"Expected more than 1 spatial element when training, got input size {x.size()}")
        if host is not None:
            styles_dev = _device_styles(host, x.device)
            present = [s in host for s in range(self.num_styles)]
        else:
            styles_dev = styles if (styles.dim() == 1 and styles.dtype == torch.int64) else styles.reshape(-1).to(torch.int64)
            present = None  # sync-free mode: absent styles get zero grads instead of None (set_sync_free_styles)
        if residual is not None and unbatched:
            residual = residual.unsqueeze(0)
        w, b = self._params()
        y = instance_cond(x, styles_dev, w, b, eps=first.eps, epilogue=epilogue, residual=residual,
                          slope=slope, present=present)
        return y.squeeze(0) if unbatched else y


def forward_fused_dual(norm_a, a: Tensor, norm_b, b: Tensor, styles, slope: float = 0.01) -> Tensor:
    """lrelu(norm_a(a, styles) + norm_b(b, styles)): the downsample branch of UnetResBlock
    (dynunet_block.py:113-125, residual = norm3(conv3(inp)), :82-98).  One kernel per direction when both norms are fast
    norms of the same family (both conditional with the same number of styles, or both plain) and the dual kernels take
    the shape; otherwise the two-call composition (norm_b, then norm_a with the add_lrelu epilogue)."""
    fam_a = isinstance(norm_a, FastForwardMixin), isinstance(norm_a, FastPlainForwardMixin)
    fam_b = isinstance(norm_b, FastForwardMixin), isinstance(norm_b, FastPlainForwardMixin)
    y = None
    if fam_a == fam_b and a.dim() == b.dim() and a.dim() == norm_a._get_no_batch_dim() + 1 and a.shape == b.shape:
        if fam_a[0] and norm_a.num_styles == norm_b.num_styles:
            for nm, x in ((norm_a, a), (norm_b, b)):
                nm._check_input_dim(x)
                nm._check_input_styles(x, styles)
            host = _host_styles(styles, norm_a.num_styles)
            first_a = norm_a._modules["norms"]._modules["0"]
            first_b = norm_b._modules["norms"]._modules["0"]
            if a.shape[1] == first_a.num_features == first_b.num_features and first_a.eps == first_b.eps:
                if host is not None:
                    styles_dev = _device_styles(host, a.device)
                    present = [s in host for s in range(norm_a.num_styles)]
                else:
                    styles_dev = styles if (styles.dim() == 1 and styles.dtype == torch.int64) else styles.reshape(-1).to(torch.int64)
                    present = None
                wa, ba = norm_a._params()
                wb, bb = norm_b._params()
                y = instance_cond_dual(a, b, styles_dev, wa, ba, wb, bb, eps=first_a.eps, slope=slope, present=present)
        elif fam_a[1] and norm_a.affine == norm_b.affine and norm_a.eps == norm_b.eps \
                and not norm_a.track_running_stats and not norm_b.track_running_stats \
                and a.shape[1] == norm_a.num_features == norm_b.num_features:
            wa = [norm_a.weight] if norm_a.affine else []
            ba = [norm_a.bias] if norm_a.affine else []
            wb = [norm_b.weight] if norm_b.affine else []
            bb = [norm_b.bias] if norm_b.affine else []
            y = instance_cond_dual(a, b, None, wa, ba, wb, bb, eps=norm_a.eps, slope=slope, present=None, num_styles=1)
    if y is None:
        residual = norm_b.forward_fused(b, styles, "none")
        y = norm_a.forward_fused(a, styles, "add_lrelu", residual=residual, slope=slope)
    return y


def _init_checks(track_running_stats: bool) -> None:
    if track_running_stats:
        raise NotImplementedError("track_running_stats=True is unreachable from MI-Seg's CLI "
                                  "(parse_normalization never sets it) and is not supported")


class _ConditionalInstanceNormBase(nn.Module):
    """Local mirror of the reference `_ConditionalInstanceNorm` (:11-68): constructor arguments,
    the `norms` ModuleList of affine nn.InstanceNorm*d (one per style) and the validation helpers."""

    def __init__(self, num_styles: int, num_features: int, eps: float = 1e-5, momentum: float = 0.1,
                 affine: bool = True, track_running_stats: bool = False, device=None, dtype=None) -> None:
        super().__init__()
        _init_checks(track_running_stats)
        if not affine:
            warnings.warn("Ignored affine=False for ConditionalInstanceNorm1D, set to True")
        factory_kwargs = {"device": device, "dtype": dtype}
        self.num_styles = num_styles
        self.norms = nn.ModuleList([
            self._get_norm()(num_features, eps, momentum, True, track_running_stats, **factory_kwargs)
            for _ in range(num_styles)])

    def _get_norm(self):
        raise NotImplementedError

    def _get_no_batch_dim(self):
        raise NotImplementedError

    def _check_input_dim(self, input):
        raise NotImplementedError

    def _check_input_styles(self, input, styles):
        if input.dim() == self._get_no_batch_dim():
            if not isinstance(styles, (int, list, Tensor)) or (isinstance(styles, Tensor) and torch.numel(styles) != 1) \
                    or (isinstance(styles, list) and len(styles) != 1):
                raise ValueError("Expected one style when input is not a batch.")
        else:
            if not isinstance(styles, (list, Tensor)) or len(styles) != len(input):
                raise ValueError("Expected number of styles as batch size.")


class _Base1d(_ConditionalInstanceNormBase):
    def _get_norm(self):
        return nn.InstanceNorm1d

    def _get_no_batch_dim(self):
        return 2

    def _check_input_dim(self, input):
        if input.dim() not in (2, 3):
            raise ValueError("expected 2D or 3D input (got {}D input)".format(input.dim()))


class _Base2d(_ConditionalInstanceNormBase):
    def _get_norm(self):
        return nn.InstanceNorm2d

    def _get_no_batch_dim(self):
        return 3

    def _check_input_dim(self, input):
        if input.dim() not in (3, 4):  # the reference's 2d message says "2D or 3D" too (:93)
            raise ValueError("expected 2D or 3D input (got {}D input)".format(input.dim()))


class _Base3d(_ConditionalInstanceNormBase):
    def _get_norm(self):
        return nn.InstanceNorm3d

    def _get_no_batch_dim(self):
        return 4

    def _check_input_dim(self, input):
        if input.dim() not in (4, 5):
            raise ValueError("expected 4D or 5D input (got {}D input)".format(input.dim()))


class FastConditionalInstanceNorm1d(FastForwardMixin, _Base1d):
    pass


class FastConditionalInstanceNorm2d(FastForwardMixin, _Base2d):
    pass


class FastConditionalInstanceNorm3d(FastForwardMixin, _Base3d):
    pass


def make_dropin_classes(ref_module):
    """Build the three fast classes on top of MI-Seg's own classes (module
    `networks.norms.conditional_instance_norm`) so `isinstance(m, _ConditionalInstanceNorm)` holds."""

    def build(name, base):
        def __init__(self, num_styles: int, num_features: int, eps: float = 1e-5, momentum: float = 0.1,
                     affine: bool = True, track_running_stats: bool = False, device=None, dtype=None) -> None:
            _init_checks(track_running_stats)
            base.__init__(self, num_styles, num_features, eps, momentum, affine, track_running_stats, device, dtype)

        return type(name, (FastForwardMixin, base), {"__init__": __init__, "__module__": __name__})

    return (build("FastConditionalInstanceNorm1d", ref_module.ConditionalInstanceNorm1d),
            build("FastConditionalInstanceNorm2d", ref_module.ConditionalInstanceNorm2d),
            build("FastConditionalInstanceNorm3d", ref_module.ConditionalInstanceNorm3d))


# ---------------------------------------------------------------------------------------------------
# the plain (unconditional) instance norm of MI-Seg's decoders through the
# same kernels.  `("instance", {"affine": True})` is the default `norm_name` of every UnetrUpBlock /
# UnetResBlock of the decoders (networks/blocks/unetr_block.py:61-85, dynunet_block.py:100-126): as many
# elements per C-Swin-UNETR forward as all instance_cond calls together.
# ---------------------------------------------------------------------------------------------------
class FastPlainForwardMixin:
    """forward() of nn.InstanceNorm{1,2,3}d (torch/nn/modules/instancenorm.py:46-56, 104-125) on the
    sm_100a kernels: one style, gamma/beta = self.weight / self.bias (or none when affine=False)."""

    def forward(self, input: Tensor) -> Tensor:
        return self.forward_fused(input, None, "none")

    def forward_fused(self, input: Tensor, styles=None, epilogue: str = "none", residual: Optional[Tensor] = None,
                      slope: float = 0.01) -> Tensor:
        """`styles` is accepted (and ignored) so blocks.fuse_blocks() can treat both norm families alike."""
        if self.track_running_stats:
            raise NotImplementedError("track_running_stats=True is not supported by the fast instance norm "
                                      "(MI-Seg never sets it)")
        self._check_input_dim(input)
        feature_dim = input.dim() - self._get_no_batch_dim()
        if input.size(feature_dim) != self.num_features:
            if self.affine:
                raise ValueError(f"expected input's size at dim={feature_dim} to match num_features "
                                 f"({self.num_features}), but got: {input.size(feature_dim)}.")
            warnings.warn(f"input's size at dim={feature_dim} does not match num_features. "
                          "You can silence this warning by not passing in num_features, "
                          "which is not used because affine=False")
        unbatched = input.dim() == self._get_no_batch_dim()
        x = input.unsqueeze(0) if unbatched else input
        m = 1
        for s in x.shape[2:]:
            m *= s
        if m <= 1 and x.numel() > 0:
            raise ValueError(f"Expected more than 1 spatial element when training, got input size {x.size()}")
        if residual is not None and unbatched:
            residual = residual.unsqueeze(0)
        w = [self.weight] if self.affine else []
        b = [self.bias] if self.affine else []
        y = instance_cond(x, None, w, b, eps=self.eps, epilogue=epilogue, residual=residual, slope=slope,
                          present=None, num_styles=1)
        return y.squeeze(0) if unbatched else y


class FastInstanceNorm1d(FastPlainForwardMixin, nn.InstanceNorm1d):
    pass


class FastInstanceNorm2d(FastPlainForwardMixin, nn.InstanceNorm2d):
    pass


class FastInstanceNorm3d(FastPlainForwardMixin, nn.InstanceNorm3d):
    pass


def fast_instance_norm(input: Tensor, weight: Optional[Tensor] = None, bias: Optional[Tensor] = None,
                       eps: float = 1e-5) -> Tensor:
    """Drop-in for `F.instance_norm(x, weight=..., bias=..., use_input_stats=True)` on a batched [N, C, *]
    CUDA tensor (e.g. SwinTransformer.proj_out, networks/nets/swin_transformer.py:135-136)."""
    if (weight is None) != (bias is None):
        raise ValueError("fast_instance_norm: pass both weight and bias, or neither")
    return instance_cond(input, None, [weight] if weight is not None else [], [bias] if bias is not None else [],
                         eps=eps, num_styles=1)
