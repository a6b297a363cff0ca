"""GPU-only time of every distinct instance_cond call of the C-Swin-UNETR list: fwd+bwd pairs replayed from a CUDA
graph (no host in the way), against the HBM roofline.  Shows which shapes are launch-bound and which path took them."""
import ctypes
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

pkg = importlib.import_module("mi-seg_b200")
lib = pkg._lib.lib()
dev = torch.device("cuda", 0)
tdt, code, es = torch.bfloat16, 1, 2
PEAK = 6542.1
REP = 10
for kv in os.environ.get("MICN_OPTS", "").split(","):  # library options, name=value
    if "=" in kv:
        pkg._lib.set_option(kv.split("=")[0], int(kv.split("=")[1]))
ONLY = os.environ.get("MICN_ONLY", "")  # e.g. "48x48,96x24": only these CxS shapes
styles = torch.tensor([1], device=dev)
extra = [("mid 96x32^3", 96, 32, False), ("mid 32x64^3", 32, 64, False), ("unet 32x128^3", 32, 128, False)]
for cs in filter(None, os.environ.get("MICN_EXTRA", "").split(",")):  # more CxS shapes
    extra.append((f"extra {cs}", int(cs.split("x")[0]), int(cs.split("x")[1]), False))
seen = set()
print(f"{'call':22s} {'C':>5s} {'S':>3s} {'chlast':>6s} {'MB':>7s} {'pair us':>8s} {'GB/s':>7s} {'frac':>5s} path")
total = 0.0
for name, c, sp, chlast in list(bench.SWIN_UNETR_CALLS) + extra:
    key = (c, sp, chlast)
    if key in seen or (ONLY and f"{c}x{sp}" not in ONLY.split(",")):
        continue
    seen.add(key)
    count = sum(1 for _, c2, s2, l2 in bench.SWIN_UNETR_CALLS if (c2, s2, l2) == key)
    m = sp ** 3
    shape = (1, c, sp, sp, sp)
    nset = max(2, int(400e6 // (c * m * es * 4)) + 1) if c * m * es < 100e6 else 3
    nset = min(nset, 24)
    sets = []
    for _ in range(nset):
        if chlast:
            x = (torch.randn(1, sp, sp, sp, c, device=dev) * 2 + 1).to(tdt)
        else:
            x = (torch.randn(*shape, device=dev) * 2 + 1).to(tdt)
        sets.append((x, torch.randn_like(x), torch.empty_like(x), torch.empty_like(x)))
    gam = [torch.rand(c, device=dev) + 0.5 for _ in range(2)]
    bet = [torch.randn(c, device=dev) for _ in range(2)]
    gp = (ctypes.c_void_p * 2)(*[t.data_ptr() for t in gam])
    bp = (ctypes.c_void_p * 2)(*[t.data_ptr() for t in bet])
    stats = torch.empty(2, c, device=dev)
    grads = torch.empty(2, 2, c, device=dev)
    wsb = lib.micn_workspace_bytes(1, c, m, code, 2)
    ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
    cws = torch.zeros(lib.micn_cl_workspace_bytes(1, c, m), dtype=torch.uint8, device=dev)

    def run(stream):
        for k in range(REP):
            x, dy, y, dx = sets[k % nset]
            if chlast:
                rc = lib.micn_fwd_cl(x.data_ptr(), y.data_ptr(), gp, bp, 2, styles.data_ptr(), stats[0].data_ptr(),
                                     stats[1].data_ptr(), 1, c, m, code, 1e-5, cws.data_ptr(), cws.numel(), stream)
                rc = rc or lib.micn_bwd_cl(dy.data_ptr(), x.data_ptr(), gp, bp, 2, styles.data_ptr(), stats[0].data_ptr(),
                                           stats[1].data_ptr(), dx.data_ptr(), grads[0].data_ptr(), grads[1].data_ptr(),
                                           1, c, m, code, cws.data_ptr(), cws.numel(), stream)
            else:
                rc = lib.micn_fwd(x.data_ptr(), y.data_ptr(), None, gp, bp, 2, styles.data_ptr(), stats[0].data_ptr(),
                                  stats[1].data_ptr(), 1, c, m, c * m, m, code, 0, 0.01, 1e-5, ws.data_ptr(), wsb, stream)
                rc = rc or lib.micn_bwd(dy.data_ptr(), x.data_ptr(), None, gp, bp, 2, styles.data_ptr(),
                                        stats[0].data_ptr(), stats[1].data_ptr(), dx.data_ptr(), None, grads[0].data_ptr(),
                                        grads[1].data_ptr(), 1, c, m, c * m, m, code, 0, 0.01, ws.data_ptr(), wsb, stream)
            if rc:
                raise RuntimeError(f"rc={rc}")

    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    side.wait_stream(torch.cuda.current_stream())  # the buffers and the zero-filled workspace were made on that stream
    with torch.cuda.stream(side):
        run(side.cuda_stream)
        with torch.cuda.graph(graph, stream=side):
            run(side.cuda_stream)
    torch.cuda.synchronize()
    for _ in range(3):
        graph.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(10):
        graph.replay()
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / (10 * REP)
    byts = 5 * c * m * es
    gbps = byts / us / 1e3
    total += us * count
    path = {0: "small", 1: "cluster", 2: "flat", 3: "chlast"}.get(pkg._lib.get_option("last_path"), "?")
    print(f"{name:22s} {c:5d} {sp:3d} {str(chlast):>6s} {c * m * es / 1e6:7.2f} {us:8.2f} {gbps:7.0f} {gbps / PEAK:5.2f} {path}"
          f"  x{count}", flush=True)
    del sets
print(f"sum over the 31 calls of one step: {total:.1f} us")
