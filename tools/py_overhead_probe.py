"""Host-side cost of one drop-in module call (tiny tensor, GPU time negligible): where do the microseconds go?"""
import cProfile
import importlib
import io
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("mi-seg_b200")
dev = torch.device("cuda:0")
mod = pkg.FastConditionalInstanceNorm3d(num_styles=2, num_features=48).to(dev)
ref = torch.nn.InstanceNorm3d(48, affine=True).to(dev)
x = torch.randn(1, 48, 8, 8, 8, device=dev, dtype=torch.bfloat16, requires_grad=True)
dy = torch.randn_like(x)
styles = torch.tensor([1], device=dev)
styles_list = [1]


def timeit(fn, reps=2000):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e6


def fwd_nograd():
    with torch.no_grad():
        mod(x, styles)


def fwd_grad():
    mod(x, styles)


def fwd_grad_list():
    mod(x, styles_list)


def fwd_bwd():
    torch.autograd.grad(mod(x, styles), [x, mod.norms[1].weight], dy)


def ref_fwd_bwd():
    torch.autograd.grad(ref(x), [x, ref.weight], dy)


def ref_fwd():
    ref(x)


for name, fn in (("ours fwd no_grad", fwd_nograd), ("ours fwd grad (device styles)", fwd_grad),
                 ("ours fwd grad (list styles)", fwd_grad_list), ("ours fwd+bwd", fwd_bwd),
                 ("torch InstanceNorm3d fwd", ref_fwd), ("torch InstanceNorm3d fwd+bwd", ref_fwd_bwd)):
    print(f"{name:36s} {timeit(fn):8.1f} us", flush=True)

for name, fn in (("fwd_grad", fwd_grad), ("fwd_bwd", fwd_bwd)):
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(2000):
        fn()
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
    print(f"==== {name} (2000 calls)")
    print("\n".join(l[:150] for l in s.getvalue().splitlines()[4:40]))
