"""Bring-up: per-phase timeline of the resident kernels (micn_res.cuh, option res_trace) for one shape, plus the
duration of single launches and of an empty-ish launch of the same kind.  MICN_SHAPE=N,C,S  MICN_DTYPE=bf16|fp32"""
import ctypes
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("mi-seg_b200")
lib = pkg._lib.lib()
dev = torch.device("cuda", 0)
n, c, sp = (int(v) for v in os.environ.get("MICN_SHAPE", "1,48,48").split(","))
tdt, code = (torch.bfloat16, 1) if os.environ.get("MICN_DTYPE", "bf16") == "bf16" else (torch.float32, 0)
for kv in os.environ.get("MICN_OPTS", "").split(","):
    if "=" in kv:
        pkg._lib.set_option(kv.split("=")[0], int(kv.split("=")[1]))
m = sp ** 3
R = 6
xs = [(torch.randn(n, c, m, device=dev) * 2 + 1).to(tdt) for _ in range(R)]
dys = [torch.randn(n, c, m, device=dev).to(tdt) for _ in range(R)]
y, dx = torch.empty_like(xs[0]), torch.empty_like(xs[0])
gam = [torch.rand(c, device=dev) + 0.5 for _ in range(2)]
bet = [torch.randn(c, device=dev) for _ in range(2)]
gp = (ctypes.c_void_p * 2)(*[t.data_ptr() for t in gam])
bp = (ctypes.c_void_p * 2)(*[t.data_ptr() for t in bet])
stats = torch.empty(2, n * c, device=dev)
grads = torch.empty(2, 2, c, device=dev)
styles = (torch.arange(n, device=dev) % 2).to(torch.int64)
wsb = lib.micn_workspace_bytes(n, c, m, code, 2)
ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream().cuda_stream


def fwd(i):
    rc = lib.micn_fwd(xs[i].data_ptr(), y.data_ptr(), None, gp, bp, 2, styles.data_ptr(), stats[0].data_ptr(),
                      stats[1].data_ptr(), n, c, m, c * m, m, code, 0, 0.01, 1e-5, ws.data_ptr(), wsb, stream)
    assert rc == 0, rc


def bwd(i):
    rc = lib.micn_bwd(dys[i].data_ptr(), xs[i].data_ptr(), None, gp, bp, 2, styles.data_ptr(), stats[0].data_ptr(),
                      stats[1].data_ptr(), dx.data_ptr(), None, grads[0].data_ptr(), grads[1].data_ptr(), n, c, m, c * m,
                      m, code, 0, 0.01, ws.data_ptr(), wsb, stream)
    assert rc == 0, rc


names = ["entry", "setup", "first", "p1", "ctasum", "xchg", "p2"]
for which, fn in (("fwd", fwd), ("bwd", bwd)):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    grid = pkg._lib.get_option("last_grid")
    print(f"== {which}: path {pkg._lib.get_option('last_path')} CS {pkg._lib.get_option('last_cs')} grid {grid}")
    # single launches, event-timed
    ts = []
    for i in range(R):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(i); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    print("   single-launch event times us:", " ".join(f"{t:.1f}" for t in ts))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(60):
        fn(i % R)
    b.record(); torch.cuda.synchronize()
    print(f"   back-to-back avg us: {a.elapsed_time(b) * 1e3 / 60:.2f}")
    if pkg._lib.get_option("last_path") != 4:
        continue
    tr = torch.zeros(grid, 8, dtype=torch.int64, device=dev)
    pkg._lib.set_option("res_trace", tr.data_ptr())
    for rep in range(3):
        tr.zero_()
        fn(3 + rep)
        torch.cuda.synchronize()
        t = tr.cpu().double()
        t0 = t[:, 0].min()
        rel = (t[:, :7] - t0) / 1e3
        print(f"   rep {rep}: " + "  ".join(f"{nm} {rel[:, k].min():.2f}/{rel[:, k].median():.2f}/{rel[:, k].max():.2f}"
                                          for k, nm in enumerate(names)) + "   (us after the first CTA's entry: min/median/max over CTAs)")
    pkg._lib.set_option("res_trace", 0)
