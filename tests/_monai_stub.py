"""Minimal stand-in for the few MONAI 1.1.0 symbols MI-Seg's `networks.layers` / `networks.blocks`
import.  MONAI is not installed in this image and there is no network.

TEST INFRASTRUCTURE ONLY: lets `tests/` and `tests/golden/make_golden.py` import the *unmodified*
reference blocks from /root/reference so the layer factory boundary and the block epilogues can
be exercised.  Written from the symbol descriptions, not from MONAI sources."""
from __future__ import annotations

import enum
import importlib
import inspect
import sys
import types

import torch.nn as nn


def _look_up_option(opt, supported, default="no_default"):
    if isinstance(supported, dict):
        if opt in supported:
            return supported[opt]
    elif isinstance(supported, type) and issubclass(supported, enum.Enum):
        for m in supported:
            if opt == m or opt == m.value:
                return m
    elif opt in supported:
        return opt
    if default != "no_default":
        return default
    raise ValueError(f"unsupported option {opt!r}; available: {list(supported)}")


def _optional_import(module, name="", **_kw):
    try:
        mod = importlib.import_module(module)
        return (getattr(mod, name) if name else mod), True
    except Exception:  # noqa: BLE001 - mirror "return a placeholder and False"
        return None, False


def _has_option(obj, keywords):
    if not callable(obj):
        return False
    params = inspect.signature(obj).parameters
    if isinstance(keywords, str):
        keywords = (keywords,)
    return all(k in params for k in keywords)


def _ensure_tuple_rep(val, dim):
    if isinstance(val, (tuple, list)):
        if len(val) == dim:
            return tuple(val)
        raise ValueError("sequence length mismatch")
    return (val,) * dim


class _BareConvolution(nn.Sequential):
    """`monai.networks.blocks.Convolution` as dynunet_block.get_conv_layer uses it
    (act=None, norm=None, dropout=None): a Sequential holding one `.conv`."""

    def __init__(self, spatial_dims, in_channels, out_channels, strides=1, kernel_size=3, act=None,
                 norm=None, dropout=None, bias=True, conv_only=False, is_transposed=False, padding=None,
                 output_padding=None, **_kw):
        super().__init__()
        if act is not None or norm is not None or dropout is not None:
            raise NotImplementedError("stub supports the bare-conv use only")
        conv_t = (nn.ConvTranspose1d, nn.ConvTranspose2d, nn.ConvTranspose3d) if is_transposed else (
            nn.Conv1d, nn.Conv2d, nn.Conv3d)
        kw = dict(kernel_size=kernel_size, stride=strides, padding=padding, bias=bias)
        if is_transposed:
            kw["output_padding"] = output_padding
        self.add_module("conv", conv_t[spatial_dims - 1](in_channels, out_channels, **kw))


def install(reference_root: str = "/root/reference") -> None:
    """Register fake `monai.*` modules (idempotent) and put the reference on sys.path."""
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    if "monai" in sys.modules and getattr(sys.modules["monai"], "__micn_stub__", False):
        return

    def mod(name):
        m = types.ModuleType(name)
        m.__micn_stub__ = True
        sys.modules[name] = m
        return m

    monai = mod("monai")
    utils = mod("monai.utils")
    utils.look_up_option = _look_up_option
    utils.optional_import = _optional_import
    utils.has_option = _has_option
    utils.ensure_tuple_rep = _ensure_tuple_rep
    monai.utils = utils

    networks = mod("monai.networks")
    layers = mod("monai.networks.layers")
    blocks = mod("monai.networks.blocks")
    monai.networks, networks.layers, networks.blocks = networks, layers, blocks

    # the reference's own fork provides the registries; alias them under the monai names
    ref_fact = importlib.import_module("networks.layers.factories")
    ref_utils = importlib.import_module("networks.layers.utils")
    fact = mod("monai.networks.layers.factories")
    for k in ("Act", "Norm", "Conv", "Dropout", "Pool", "Pad", "split_args", "LayerFactory"):
        setattr(fact, k, getattr(ref_fact, k))
    lutils = mod("monai.networks.layers.utils")
    lutils.get_act_layer = ref_utils.get_act_layer
    lutils.get_norm_layer = ref_utils.get_norm_layer
    layers.factories, layers.utils = fact, lutils

    convutils = mod("monai.networks.layers.convutils")
    convutils.same_padding = lambda k, d=1: tuple(((kk - 1) // 2) * d for kk in k) if isinstance(k, (tuple, list)) else ((k - 1) // 2) * d
    convutils.stride_minus_kernel_padding = lambda k, s: tuple(ss - kk for kk, ss in zip(_ensure_tuple_rep(k, len(s)), s)) if isinstance(s, (tuple, list)) else s - k
    layers.convutils = convutils

    bconv = mod("monai.networks.blocks.convolutions")
    bconv.Convolution = _BareConvolution
    blocks.convolutions = bconv
    blocks.Convolution = _BareConvolution
