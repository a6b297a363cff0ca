"""Minimal stand-in for the few MONAI 1.1.0 symbols MI-Seg's `networks/` package imports.
MONAI is not installed in this image and there is no network.

TEST / BENCH INFRASTRUCTURE ONLY: lets `tests/`, `tests/golden/make_golden.py` and the model-level legs of `bench.py`
import the *unmodified* reference package (from /root/reference in the build container, or from the git-ignored copy
`baseline/_ref/networks` that `baseline/make_ref.py` makes so that it travels to the GPU box) and so exercise the layer
factory boundary, the block epilogues and whole C-UNet / C-UNETR / C-Swin-UNETR steps.  Written from the symbol
descriptions of SURVEY.md appendix A, not from MONAI sources.  `mi-seg_b200/` never imports this file.
"""
from __future__ import annotations

import enum
import importlib
import inspect
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------ monai.utils
def _look_up_option(opt, supported, default="no_default", print_all_options=True):
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


def _noop_decorator(*_a, **_kw):
    """alias(...), export(...), deprecated_arg(...): decorators the nets use for bookkeeping only."""
    def wrap(obj):
        return obj
    return wrap


class _SkipMode(enum.Enum):
    CAT = "cat"
    ADD = "add"
    MUL = "mul"


# ------------------------------------------------------------------------------------------------ monai.networks.blocks
class _Convolution(nn.Sequential):
    """`monai.networks.blocks.Convolution` as dynunet_block.get_conv_layer uses it (act=None, norm=None, dropout=None):
    a Sequential holding one `.conv` (the attribute name is part of the checkpoint keys, `...conv1.conv.weight`)."""

    def __init__(self, spatial_dims, in_channels, out_channels, strides=1, kernel_size=3, act=None,
                 norm=None, dropout=None, bias=True, conv_only=False, is_transposed=False, padding=None,
                 output_padding=None, **_kw):
        super().__init__()
        if not conv_only and (act is not None or norm is not None or dropout is not None):
            raise NotImplementedError("stub supports the bare-conv use only")
        conv_t = (nn.ConvTranspose1d, nn.ConvTranspose2d, nn.ConvTranspose3d) if is_transposed else (
            nn.Conv1d, nn.Conv2d, nn.Conv3d)
        if padding is None:
            padding = (kernel_size - 1) // 2 if isinstance(kernel_size, int) else tuple((k - 1) // 2 for k in kernel_size)
        kw = dict(kernel_size=kernel_size, stride=strides, padding=padding, bias=bias)
        if is_transposed:
            kw["output_padding"] = output_padding if output_padding is not None else 0
        self.add_module("conv", conv_t[spatial_dims - 1](in_channels, out_channels, **kw))


class _MLPBlock(nn.Module):
    """linear1 -> act -> drop -> linear2 -> drop (the attribute names linear1 / linear2 are used by the nets' `load_from`)."""

    def __init__(self, hidden_size, mlp_dim, dropout_rate=0.0, act="GELU", dropout_mode="vit"):
        super().__init__()
        mlp_dim = mlp_dim or hidden_size
        self.linear1 = nn.Linear(hidden_size, mlp_dim)
        self.linear2 = nn.Linear(mlp_dim, hidden_size)
        self.fn = nn.GELU() if (isinstance(act, str) and act.upper() == "GELU") else (act() if isinstance(act, type) else nn.GELU())
        self.drop1 = nn.Dropout(dropout_rate)
        self.drop2 = nn.Dropout(dropout_rate)

    def forward(self, x):
        return self.drop2(self.linear2(self.drop1(self.fn(self.linear1(x)))))


class _SABlock(nn.Module):
    """Standard multi-head self-attention over [B, L, hidden]: qkv Linear, softmax(QK^T / sqrt(d)) V, out_proj."""

    def __init__(self, hidden_size, num_heads, dropout_rate=0.0, qkv_bias=False, **_kw):
        super().__init__()
        if hidden_size % num_heads:
            raise ValueError("hidden size should be divisible by num_heads.")
        self.num_heads = num_heads
        self.out_proj = nn.Linear(hidden_size, hidden_size)
        self.qkv = nn.Linear(hidden_size, hidden_size * 3, bias=qkv_bias)
        self.drop_output = nn.Dropout(dropout_rate)
        self.dropout_rate = dropout_rate
        self.head_dim = hidden_size // num_heads
        self.scale = self.head_dim ** -0.5

    def forward(self, x):
        b, l, h = x.shape
        qkv = self.qkv(x).reshape(b, l, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        o = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2],
                                           dropout_p=self.dropout_rate if self.training else 0.0)
        return self.drop_output(self.out_proj(o.transpose(1, 2).reshape(b, l, h)))


class _DropPath(nn.Module):
    """Stochastic depth per sample (rate 0 by default in MI-Seg, parser.py:38)."""

    def __init__(self, drop_prob=0.0, scale_by_keep=True):
        super().__init__()
        self.drop_prob, self.scale_by_keep = drop_prob, scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask


def _trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


def install(reference_root: str = "/root/reference") -> None:
    """Register fake `monai.*` modules (idempotent) and put the reference checkout on sys.path."""
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
    utils.alias = utils.export = utils.deprecated_arg = _noop_decorator
    utils.SkipMode = _SkipMode
    umod = mod("monai.utils.module")
    umod.look_up_option, umod.optional_import = _look_up_option, _optional_import
    utils.module = umod
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
        setattr(layers, k, getattr(ref_fact, k))
    lutils = mod("monai.networks.layers.utils")
    for k in ("get_act_layer", "get_norm_layer", "get_dropout_layer", "get_pool_layer"):
        if hasattr(ref_utils, k):
            setattr(lutils, k, getattr(ref_utils, k))
            setattr(layers, k, getattr(ref_utils, k))
    layers.factories, layers.utils = fact, lutils
    layers.DropPath = _DropPath
    layers.trunc_normal_ = _trunc_normal_

    convutils = mod("monai.networks.layers.convutils")
    convutils.same_padding = lambda k, d=1: tuple(((kk - 1) // 2) * d for kk in k) if isinstance(k, (tuple, list)) else ((k - 1) // 2) * d
    convutils.stride_minus_kernel_padding = lambda k, s: tuple(ss - kk for kk, ss in zip(_ensure_tuple_rep(k, len(s)), s)) if isinstance(s, (tuple, list)) else s - k
    layers.convutils = convutils

    bconv = mod("monai.networks.blocks.convolutions")
    bconv.Convolution = _Convolution
    blocks.convolutions = bconv
    blocks.Convolution = _Convolution
    bmlp = mod("monai.networks.blocks.mlp")
    bmlp.MLPBlock = _MLPBlock
    blocks.mlp, blocks.MLPBlock = bmlp, _MLPBlock
    bsa = mod("monai.networks.blocks.selfattention")
    bsa.SABlock = _SABlock
    blocks.selfattention, blocks.SABlock = bsa, _SABlock
    # MI-Seg carries its own copy of the patch-embedding block (networks/blocks/patch_embedding.py:32-123) but vit.py:19
    # still imports MONAI's: hand it the local one
    bpe = mod("monai.networks.blocks.patchembedding")
    local_pe = importlib.import_module("networks.blocks.patch_embedding")
    bpe.PatchEmbeddingBlock = local_pe.PatchEmbeddingBlock
    blocks.patchembedding, blocks.PatchEmbeddingBlock = bpe, local_pe.PatchEmbeddingBlock


def reference_root() -> str | None:
    """Where an importable copy of the reference's `networks/` package is: the read-only checkout in the build
    container, else the git-ignored `baseline/_ref` copy that travels to the GPU box; None when neither exists."""
    import os

    here = os.path.dirname(os.path.abspath(__file__))
    for root in ("/root/reference", os.path.join(here, "_ref")):
        if os.path.isfile(os.path.join(root, "networks", "norms", "conditional_instance_norm.py")):
            return root
    return None
