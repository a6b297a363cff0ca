"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py

* norm_*.npz  - `networks.norms.conditional_instance_norm.ConditionalInstanceNorm{1,2,3}d`
                forward + autograd backward on seeded inputs (needs only torch).
* block_*.npz - the real `networks.blocks.dynunet_block.UnetResBlock` / `UnetBasicBlock`
                (imported through tests/_monai_stub.py because MONAI is absent) with hooks that
                record the tensors entering every norm, the block output and all gradients.

The GPU box has no /root/reference: tests read only the committed .npz files."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")

from networks.norms.conditional_instance_norm import (  # noqa: E402
    ConditionalInstanceNorm1d, ConditionalInstanceNorm2d, ConditionalInstanceNorm3d)

CLS = {1: ConditionalInstanceNorm1d, 2: ConditionalInstanceNorm2d, 3: ConditionalInstanceNorm3d}


def _np(t):
    return None if t is None else t.detach().to(torch.float32).cpu().numpy()


def norm_case(name, dim, shape, styles, num_styles, dtype=torch.float32, seed=0, mean=1.0, std=2.0,
              styles_as="list", batched=True):
    g = torch.Generator().manual_seed(seed)
    c = shape[1] if batched else shape[0]
    mod = CLS[dim](num_styles=num_styles, num_features=c)
    with torch.no_grad():
        for s in range(num_styles):
            mod.norms[s].weight.copy_(1.0 + 0.3 * torch.randn(c, generator=g))
            mod.norms[s].bias.copy_(0.3 * torch.randn(c, generator=g))
    x32 = torch.randn(*shape, generator=g) * std + mean
    dy32 = torch.randn(*shape, generator=g)
    x = x32.to(dtype).requires_grad_(True)
    dy = dy32.to(dtype)
    if styles_as == "list":
        st = list(styles)
    elif styles_as == "tensor":
        st = torch.tensor(styles, dtype=torch.int64)
    elif styles_as == "tensor_b1":
        st = torch.tensor(styles, dtype=torch.int64).reshape(-1, 1)
    elif styles_as == "int":
        st = int(styles)
    else:
        raise ValueError(styles_as)
    y = mod(x, st)
    y.backward(dy)
    out = dict(
        x=_np(x), dy=_np(dy), y=_np(y), dx=_np(x.grad),
        styles=np.asarray(styles, dtype=np.int64).reshape(-1), num_styles=np.int64(num_styles),
        dtype=str(dtype).replace("torch.", ""), dim=np.int64(dim), batched=np.bool_(batched),
        styles_as=styles_as,
        gamma=np.stack([_np(n.weight) for n in mod.norms]), beta=np.stack([_np(n.bias) for n in mod.norms]),
        present=np.array([n.weight.grad is not None for n in mod.norms]),
        dgamma=np.stack([_np(n.weight.grad) if n.weight.grad is not None else np.zeros(c, np.float32) for n in mod.norms]),
        dbeta=np.stack([_np(n.bias.grad) if n.bias.grad is not None else np.zeros(c, np.float32) for n in mod.norms]),
        y_is_contiguous=np.bool_(y.is_contiguous()),
    )
    np.savez_compressed(os.path.join(HERE, f"norm_{name}.npz"), **out)
    print("wrote", name, tuple(shape), out["dtype"])


def block_case(name, kind, cin, cout, stride, spatial, styles, num_styles=2, seed=0):
    import _monai_stub
    _monai_stub.install()
    from networks.blocks.dynunet_block import UnetBasicBlock, UnetResBlock

    torch.manual_seed(seed)
    cls = UnetResBlock if kind == "res" else UnetBasicBlock
    blk = cls(3, cin, cout, kernel_size=3, stride=stride,
              norm_name=("instance_cond", {"num_styles": num_styles, "affine": True}))
    with torch.no_grad():
        for nm, m in blk.named_modules():
            if nm in ("norm1", "norm2", "norm3"):
                for s in range(num_styles):
                    m.norms[s].weight.copy_(1.0 + 0.3 * torch.randn(cout))
                    m.norms[s].bias.copy_(0.3 * torch.randn(cout))
    b = len(styles)
    x = (torch.randn(b, cin, *spatial) * 1.5 + 0.5).requires_grad_(True)
    st = torch.tensor(styles, dtype=torch.int64)
    cap = {}

    def hook(nm):
        def f(_m, _i, o):
            o.retain_grad()
            cap[nm] = o
        return f

    hs = [getattr(blk, nm).register_forward_hook(hook(nm)) for nm in ("conv1", "conv2", "conv3", "norm3")
          if hasattr(blk, nm)]
    out = blk(x, st)
    dout = torch.randn_like(out)
    out.backward(dout)
    for h in hs:
        h.remove()
    rec = dict(kind=kind, styles=np.asarray(styles, np.int64), num_styles=np.int64(num_styles),
               x=_np(x), dx=_np(x.grad), out=_np(out), dout=_np(dout), stride=np.int64(stride))
    for nm, t in cap.items():
        rec[f"{nm}_out"] = _np(t)
        rec[f"{nm}_out_grad"] = _np(t.grad)
    for nm in ("norm1", "norm2", "norm3"):
        if hasattr(blk, nm):
            m = getattr(blk, nm)
            rec[f"{nm}_gamma"] = np.stack([_np(n.weight) for n in m.norms])
            rec[f"{nm}_beta"] = np.stack([_np(n.bias) for n in m.norms])
            rec[f"{nm}_present"] = np.array([n.weight.grad is not None for n in m.norms])
            rec[f"{nm}_dgamma"] = np.stack([_np(n.weight.grad) if n.weight.grad is not None else np.zeros(cout, np.float32) for n in m.norms])
            rec[f"{nm}_dbeta"] = np.stack([_np(n.bias.grad) if n.bias.grad is not None else np.zeros(cout, np.float32) for n in m.norms])
    for nm in ("conv1", "conv2", "conv3"):
        if hasattr(blk, nm):
            rec[f"{nm}_weight"] = _np(getattr(blk, nm).conv.weight)
            rec[f"{nm}_weight_grad"] = _np(getattr(blk, nm).conv.weight.grad)
    np.savez_compressed(os.path.join(HERE, f"block_{name}.npz"), **rec)
    print("wrote block", name)


def main():
    norm_case("3d_mixed", 3, (3, 4, 3, 4, 5), [0, 1, 0], 2)
    norm_case("3d_tensor_b1_neg", 3, (4, 5, 4, 4, 4), [1, -1, 0, -2], 2, styles_as="tensor_b1", seed=1)
    norm_case("3d_vec", 3, (2, 6, 8, 8, 8), [1, 0], 2, styles_as="tensor", seed=2)
    norm_case("3d_odd", 3, (2, 5, 3, 3, 3), [0, 1], 2, styles_as="tensor", seed=3)
    norm_case("3d_bigmean", 3, (2, 3, 6, 6, 6), [1, 1], 2, seed=4, mean=50.0, std=0.1)
    norm_case("3d_absent_style", 3, (2, 4, 4, 4, 4), [2, 2], 3, seed=5)
    norm_case("3d_unbatched_int", 3, (4, 3, 4, 5), 1, 2, styles_as="int", batched=False, seed=6)
    norm_case("3d_bf16", 3, (2, 8, 4, 8, 8), [0, 1], 2, dtype=torch.bfloat16, styles_as="tensor", seed=7)
    norm_case("1d_tokens", 1, (3, 12, 27), [2, 0, 1], 3, styles_as="tensor", seed=8)
    norm_case("2d", 2, (2, 3, 9, 7), [0, 1], 2, seed=9)
    norm_case("3d_mid", 3, (2, 8, 16, 16, 16), [0, 1], 2, styles_as="tensor", seed=10)
    block_case("res_down", "res", 2, 4, 1, (6, 6, 6), [0, 1])
    block_case("res_identity", "res", 4, 4, 1, (6, 6, 6), [1, 0], seed=1)
    block_case("res_stride2", "res", 3, 6, 2, (8, 8, 8), [1, 1], seed=2)
    block_case("basic", "basic", 2, 4, 1, (6, 6, 6), [0, 1], seed=3)


if __name__ == "__main__":
    main()
