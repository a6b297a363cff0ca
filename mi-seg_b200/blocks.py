"""Fused epilogues for MI-Seg's DynUNet blocks: the LeakyReLU / residual add that follows every
`instance_cond` norm in `UnetResBlock.forward` (networks/blocks/dynunet_block.py:100-126) and
`UnetBasicBlock.forward` (:187-203) is folded into the norm kernel, so each activation is read once
and written once per norm instead of three to five times.

    reference                                   fused
    out = lrelu(norm1(conv1(x)))                norm1.forward_fused(conv1(x), m, "lrelu")
    out = norm2(conv2(out)); out += residual;   norm2.forward_fused(conv2(out), m, "add_lrelu", residual)
    out = lrelu(out)
    (downsample branch) residual = norm3(conv3(x))  forward_fused_dual(norm2, conv2(out), norm3, conv3(x), m)

C-UNet's `ADN` blocks (acti_norm.py:104-110, "NDA": norm -> dropout(0) -> PReLU) fuse the same way, with the PReLU
weight read on the device (`adn_forward`).

`fuse_blocks(model)` rebinds `forward` on the block INSTANCES whose norms are the fast classes; class
definitions, parameters and state-dict keys of `networks/` stay untouched.  Convolutions remain
PyTorch/cuDNN (north_star).
"""
from __future__ import annotations

import types

import torch.nn as nn

from .norms import FastForwardMixin, FastPlainForwardMixin, forward_fused_dual

_MISSING = "Modalities must be passed to the forward step when encoder_norm_type is 'instance_cond'."


def _slope(block) -> float:
    return float(getattr(block.lrelu, "negative_slope", 0.01))


def _needs_modalities(block) -> bool:
    return any(isinstance(getattr(block, n, None), FastForwardMixin) for n in ("norm1", "norm2", "norm3"))


def unet_res_block_forward(self, inp, modalities=None):
    """Drop-in for UnetResBlock.forward (dynunet_block.py:100-126) with fused epilogues.  Works for the
    conditional norms (modalities required, as in the reference) and for the decoders' plain instance norms."""
    if modalities is None and _needs_modalities(self):
        raise ValueError(_MISSING)
    slope = _slope(self)
    out = self.conv1(inp)
    out = self.norm1.forward_fused(out, modalities, "lrelu", slope=slope)
    out = self.conv2(out)
    residual = inp
    if hasattr(self, "conv3"):
        residual = self.conv3(residual)
    if hasattr(self, "norm3"):
        # downsample branch (:82-98, :113-118): lrelu(norm2(out) + norm3(residual)), both norms in ONE kernel per direction
        return forward_fused_dual(self.norm2, out, self.norm3, residual, modalities, slope=slope)
    return self.norm2.forward_fused(out, modalities, "add_lrelu", residual=residual, slope=slope)


def unet_basic_block_forward(self, inp, modalities=None):
    """Drop-in for UnetBasicBlock.forward (dynunet_block.py:187-203) with fused epilogues."""
    if modalities is None and _needs_modalities(self):
        raise ValueError(_MISSING)
    slope = _slope(self)
    out = self.conv1(inp)
    out = self.norm1.forward_fused(out, modalities, "lrelu", slope=slope)
    out = self.conv2(out)
    return self.norm2.forward_fused(out, modalities, "lrelu", slope=slope)


def adn_forward(self, input, modalities=None):
    """Drop-in for ADN.forward (networks/blocks/acti_norm.py:104-110) in the "NDA" ordering C-UNet uses
    (norm -> dropout(p=0) -> PReLU, convolutions.py:173-179): prelu(norm(x)) in one kernel, the PReLU weight read
    on the device.  A conditional norm validates `modalities` exactly as it does in the reference."""
    act = self.A
    slope = act.weight if isinstance(act, nn.PReLU) else float(act.negative_slope)
    return self.N.forward_fused(input, modalities, "lrelu", slope=slope)


def _adn_fusable(block) -> bool:
    names = [n for n, _ in block.named_children()]
    if names not in (["N", "A"], ["N", "D", "A"]):
        return False
    if not isinstance(block.N, (FastForwardMixin, FastPlainForwardMixin)):
        return False
    if "D" in names and not (isinstance(block.D, (nn.Dropout, nn.Dropout2d, nn.Dropout3d)) and block.D.p == 0):
        return False
    act = block.A
    return (isinstance(act, nn.PReLU) and act.weight.numel() == 1) or isinstance(act, nn.LeakyReLU)


def _fusable(block) -> bool:
    fast = (FastForwardMixin, FastPlainForwardMixin)
    norms = [getattr(block, n, None) for n in ("norm1", "norm2")]
    if not all(isinstance(n, fast) for n in norms):
        return False
    if hasattr(block, "norm3") and not isinstance(block.norm3, fast):
        return False
    return isinstance(getattr(block, "lrelu", None), nn.LeakyReLU) and hasattr(block, "conv1") and hasattr(block, "conv2")


def fuse_blocks(model: nn.Module) -> int:
    """Rebind forward() of every UnetResBlock / UnetBasicBlock under `model` whose norms are fast
    instance_cond modules and whose activation is LeakyReLU.  Returns the number of blocks fused."""
    count = 0
    for m in model.modules():
        name = type(m).__name__
        if name == "UnetResBlock" and _fusable(m):
            m.forward = types.MethodType(unet_res_block_forward, m)
            count += 1
        elif name == "UnetBasicBlock" and _fusable(m):
            m.forward = types.MethodType(unet_basic_block_forward, m)
            count += 1
        elif name == "ADN" and _adn_fusable(m):
            m.forward = types.MethodType(adn_forward, m)
            count += 1
    return count
