"""bring-up: per-call GPU time of the C-Swin-UNETR norm call list (ours vs PyTorch/ATen), fwd+bwd via the module."""
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import mi_seg_b200 as pkg

dev = torch.device("cuda", 0)
tdt = torch.bfloat16
styles = torch.tensor([1], device=dev)
seen = set()
print(f"{'call':22s} {'C':>5s} {'S':>3s} {'chlast':>6s} {'ours us':>9s} {'torch us':>9s} {'host us':>8s} path")
for name, c, sp, chlast in bench.SWIN_UNETR_CALLS:
    key = (c, sp, chlast)
    if key in seen:
        continue
    seen.add(key)
    mod = pkg.FastConditionalInstanceNorm3d(num_styles=2, num_features=c).to(dev)
    if chlast:
        x = (torch.randn(1, sp, sp, sp, c, device=dev) * 2 + 1).to(tdt).permute(0, 4, 1, 2, 3)
    else:
        x = (torch.randn(1, c, sp, sp, sp, device=dev) * 2 + 1).to(tdt)
    x.requires_grad_(True)
    dy = torch.randn(1, c, sp, sp, sp, device=dev).to(tdt)

    def ours():
        mod(x, styles).backward(dy)

    def ref():
        w, b = mod.norms[1].weight, mod.norms[1].bias
        torch.stack([F.instance_norm(x[i].unsqueeze(0), None, None, w, b, True, 0.1, 1e-5).squeeze(0)
                     for i in range(1)]).backward(dy)

    res = []
    for fn in (ours, ref):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        for _ in range(20):
            fn()
        b_.record()
        host = (time.perf_counter() - t0) / 20 * 1e6
        torch.cuda.synchronize()
        res.append((a.elapsed_time(b_) / 20 * 1e3, host))
    print(f"{name:22s} {c:5d} {sp:3d} {str(chlast):>6s} {res[0][0]:9.1f} {res[1][0]:9.1f} {res[0][1]:8.1f} {pkg._lib.get_option('last_path')}")

# host profile of one small call (where does the Python time go?)
import cProfile
import pstats

mod = pkg.FastConditionalInstanceNorm3d(num_styles=2, num_features=96).to(dev)
x = torch.randn(1, 96, 24, 24, 24, device=dev).to(tdt).requires_grad_(True)
dy = torch.randn(1, 96, 24, 24, 24, device=dev).to(tdt)
for _ in range(5):
    mod(x, styles).backward(dy)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    mod(x, styles).backward(dy)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
