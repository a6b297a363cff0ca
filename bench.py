#!/usr/bin/env python
"""bench.py - the instance_cond hot path on B200: `python bench.py --gpus N --steps K --warmup W`.

A step = one pass of the hot path over one batch: `micn_fwd` + `micn_bwd` (include/micn.h) on the
north-star activation, 1 x 48 x 96^3 bf16 per GPU (C-Swin-UNETR f=48 `encoder1.norm*`, BASELINE.json
configs[1]).  N > 1: one process per GPU (torchrun), every rank runs its own
patch (the norm is per-sample, no forward collective) and the per-style d(gamma)/d(beta) are
all-reduced over NCCL inside the step - weak scaling.

Prints ONE JSON line (rank 0):
  value     algorithmic GB/s, (2 + 3) * E * s bytes per step per GPU, inputs resident in HBM, CUDA events
  e2e       the same metric through `micn_fwd_bwd_host` (host pinned buffers in, host buffers out;
            H2D / D2H inside the timed region)
  roofline  the dominant kernel (backward: 3 * E * s bytes per launch, average duration over a region of
            back-to-back launches of that kernel, CUDA events on its stream) against MEASURED_PEAKS.json
  cpu_baseline  the reference's own ConditionalInstanceNorm3d (unmodified, from the git-ignored copy oracle/_ref/ that
            baseline/make_ref.py makes; kind "reference") on the box's host cores; the oracle's torch-CPU port of the same
            call sequence (kind "port") only when that copy is absent
`--impl reference` times that CPU module alone on the same config.

Timing of `value`: the clocks sampler starts, then an untimed pre-warm of 200 steps brings the GPU to its load
clocks, then `--regions` (default 5) timed regions of EXACTLY --steps steps each, every one bracketed by a
barrier + synchronize and CUDA events, max over ranks per region; the MEDIAN region is reported (all of them are listed
under "regions_ms").  Extra keys: model-level steps of the reference's own nets ("model_step": C-Swin-UNETR B=1/GPU,
DDP over NCCL for N > 1; C-UNETR B=4 and the C-UNet CPU config at N=1) and the sharded sliding-window volume
("sliding_window", BASELINE.json configs[4]) - see baseline/model_bench.py.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "instance_cond_fwd_bwd_hbm_GBps"
UNIT = "GB/s"
DT_BYTES = {"bf16": 2, "fp32": 4, "fp16": 2}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="1,48,96", help="N,C,S per GPU (S^3 voxels per slab)")
    ap.add_argument("--dtype", default="bf16", choices=list(DT_BYTES))
    ap.add_argument("--epilogue", default="none", choices=["none", "lrelu"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-torch-ref", action="store_true", help="skip the PyTorch/ATen GPU comparison leg (clean ncu launch lists)")
    ap.add_argument("--no-model-calls", action="store_true", help="skip the C-Swin-UNETR norm-call-list leg")
    ap.add_argument("--regions", type=int, default=5, help="timed regions of --steps steps each; the median is reported")
    ap.add_argument("--launch", default="stream", choices=["stream", "graph", "graphR"],
                    help="stream (default): the two C-ABI calls are issued from Python every step; graph: the micn_fwd + micn_bwd "
                         "pair of every buffer set is captured once into a CUDA graph and replayed (the calls keep no per-launch "
                         "state on the host); graphR: ONE graph holds the steps of all R buffer sets (--steps must be a multiple "
                         "of R; N > 1 needs the fused exchange) - programmatic dependent launch then spans the step boundaries")
    ap.add_argument("--collective", default="fused", choices=["fused", "nccl"],
                    help="N > 1: how d(gamma)/d(beta) are all-reduced every step.  fused (default): inside the backward kernel, "
                         "records stored straight into every peer's memory over NVLink (micn_bwd_allreduce); nccl: an "
                         "asynchronous ncclAllReduce per step overlapped with the next step (round 1's scheme)")
    ap.add_argument("--model-steps", default="auto", help="comma list of swin_unetr,unetr,unet_cpu,sliding_window; auto = "
                    "swin_unetr + sliding_window at every N, plus unetr and unet_cpu at N=1; none = skip")
    ap.add_argument("--model-step-iters", type=int, default=8)
    ap.add_argument("--sweep", action="store_true", help="also time the BASELINE.json microbench sweep (extra key)")
    ap.add_argument("--sweep-out", default=None, help="append every sweep point to this file as JSON lines")
    return ap.parse_args()


def workload(args):
    n, c, s = (int(v) for v in args.shape.split(","))
    return n, c, s, s * s * s


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def run_sweep(lib, pkg, dev, peak, out_path=None, budget_bytes=120e9):
    """BASELINE.json configs[3]: N in {1,4,8} x C in {24,48,96,192,384} x {48^3,96^3,128^3}, fp32 and bf16, forward and
    backward, device-resident inputs, CUDA events, rotating buffer sets > L2 (or one set when a tensor alone
    exceeds L2).  Points whose four live tensors exceed `budget_bytes` are skipped."""
    import torch

    rows = []
    S = 2
    stream = torch.cuda.current_stream().cuda_stream
    for dtype, tdt, code, es in (("bf16", torch.bfloat16, 1, 2), ("fp32", torch.float32, 0, 4)):
        for sp in (48, 96, 128):
            m = sp ** 3
            for n in (1, 4, 8):
                for c in (24, 48, 96, 192, 384):
                    E = n * c * m
                    R = max(1, min(3, int(400e6 // (4 * E * es)) + 1))
                    if 4 * E * es * R > budget_bytes:
                        R = 1
                    if 4 * E * es * R > budget_bytes:
                        rows.append({"N": n, "C": c, "S": sp, "dtype": dtype, "skipped": "exceeds the memory budget"})
                        continue
                    xs = [torch.empty(n, c, m, device=dev, dtype=tdt).normal_(1.0, 2.0) for _ in range(R)]
                    dys = [torch.empty(n, c, m, device=dev, dtype=tdt).normal_() for _ in range(R)]
                    ys = [torch.empty_like(xs[0]) for _ in range(R)]
                    dxs = [torch.empty_like(xs[0]) for _ in range(R)]
                    mean = torch.empty(n * c, device=dev)
                    rstd = torch.empty(n * c, device=dev)
                    gamma = 1 + 0.3 * torch.randn(S, c, device=dev)
                    beta = 0.3 * torch.randn(S, c, device=dev)
                    styles = (torch.arange(n, device=dev) % S).to(torch.int64)
                    grads = torch.empty(2, S, c, device=dev)
                    wsb = lib.micn_workspace_bytes(n, c, m, code, S)
                    ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
                    gp = (ctypes.c_void_p * S)(*[gamma[k].data_ptr() for k in range(S)])
                    bp = (ctypes.c_void_p * S)(*[beta[k].data_ptr() for k in range(S)])

                    def fwd(i):
                        rc = lib.micn_fwd(xs[i].data_ptr(), ys[i].data_ptr(), None, gp, bp, S, styles.data_ptr(),
                                          mean.data_ptr(), rstd.data_ptr(), n, c, m, c * m, m, code, 0, 0.01, 1e-5,
                                          ws.data_ptr(), wsb, stream)
                        if rc:
                            raise RuntimeError(f"micn_fwd rc={rc}")

                    def bwd(i):
                        rc = lib.micn_bwd(dys[i].data_ptr(), xs[i].data_ptr(), None, gp, bp, S, styles.data_ptr(),
                                          mean.data_ptr(), rstd.data_ptr(), dxs[i].data_ptr(), None,
                                          grads[0].data_ptr(), grads[1].data_ptr(), n, c, m, c * m, m, code, 0, 0.01,
                                          ws.data_ptr(), wsb, stream)
                        if rc:
                            raise RuntimeError(f"micn_bwd rc={rc}")

                    iters = max(10, min(30, int(3e9 // (5 * E * es)) + 1))
                    res = {}
                    for name, fn in (("fwd", fwd), ("bwd", bwd)):
                        for i in range(3):
                            fn(i % R)
                        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        a.record()
                        for i in range(iters):
                            fn(i % R)
                        b_.record()
                        torch.cuda.synchronize()
                        res[name] = a.elapsed_time(b_) / iters * 1e3  # us
                    fb, bb = 2 * E * es, 3 * E * es
                    row = {"N": n, "C": c, "S": sp, "dtype": dtype, "path": int(pkg._lib.get_option("last_path")),
                           "fwd_us": round(res["fwd"], 2), "bwd_us": round(res["bwd"], 2),
                           "fwd_gbps": round(fb / res["fwd"] * 1e-3, 1), "bwd_gbps": round(bb / res["bwd"] * 1e-3, 1),
                           "frac": round((fb + bb) / (res["fwd"] + res["bwd"]) * 1e-3 / peak, 3),
                           "voxels_per_s": round(n * m / ((res["fwd"] + res["bwd"]) * 1e-6)), "buffer_sets": R}
                    rows.append(row)
                    if out_path:
                        with open(out_path, "a") as f:
                            f.write(json.dumps(row) + "\n")
                    del xs, dys, ys, dxs, ws
                    torch.cuda.empty_cache()
    return rows


SWIN_UNETR_CALLS = (  # the 31 instance_cond calls of one C-Swin-UNETR (f=48, 96^3, B=1) forward
    [("encoder1.norm%d" % i, 48, 96, False) for i in (1, 2, 3)] +
    [("encoder2.norm%d" % i, 48, 48, False) for i in (1, 2)] + [("encoder3.norm%d" % i, 96, 24, False) for i in (1, 2)] +
    [("encoder4.norm%d" % i, 192, 12, False) for i in (1, 2)] + [("encoder10.norm%d" % i, 768, 3, False) for i in (1, 2)] +
    [("swin.l1.b%d.norm%d" % (b, i), 48, 48, False) for b in (0, 1) for i in (1, 2)] +
    [("swin.l2.b%d.norm%d" % (b, i), 96, 24, False) for b in (0, 1) for i in (1, 2)] +
    [("swin.l3.b%d.norm%d" % (b, i), 192, 12, False) for b in (0, 1) for i in (1, 2)] +
    [("swin.l4.b%d.norm%d" % (b, i), 384, 6, False) for b in (0, 1) for i in (1, 2)] +
    [("merge1.norm", 384, 24, True), ("merge2.norm", 768, 12, True), ("merge3.norm", 1536, 6, True),
     ("merge4.norm", 3072, 3, True)])


def run_model_calls(pkg, dev, tdt, reps=20):
    """The hot path at model scale: every instance_cond call of one C-Swin-UNETR training step (forward + backward),
    timed as one sequence four ways: (1) raw C-ABI launches back to back from Python (host-bound for the small calls),
    (1b) the same launches replayed from a CUDA graph (what the GPU needs), (2) through the
    drop-in nn.Module + autograd (what a training script calls; ~190 us of Python / autograd-engine time per call,
    which a real step hides behind its convolutions), (3) the reference's call sequence (per-sample F.instance_norm
    + torch.stack) through PyTorch/ATen on the same GPU.  PatchMerging inputs arrive channels-last (stride_C = 1)."""
    import torch
    import torch.nn.functional as F

    lib = pkg._lib.lib()
    torch.manual_seed(1)
    code = {torch.bfloat16: 1, torch.float32: 0, torch.float16: 2}[tdt]
    stream = torch.cuda.current_stream().cuda_stream
    calls, raw = [], []
    elems = 0
    for name, c, sp, chlast in SWIN_UNETR_CALLS:
        mod = pkg.FastConditionalInstanceNorm3d(num_styles=2, num_features=c).to(dev)
        shape = (1, c, sp, sp, sp)
        if chlast:
            x = (torch.randn(1, sp, sp, sp, c, device=dev) * 2 + 1).to(tdt).permute(0, 4, 1, 2, 3)
        else:
            x = (torch.randn(*shape, device=dev) * 2 + 1).to(tdt)
        x.requires_grad_(True)
        dy = torch.randn(*shape, device=dev).to(tdt)
        calls.append((mod, x, dy))
        m = sp ** 3
        wsb = lib.micn_workspace_bytes(1, c, m, code, 2)
        raw.append({"x": x.detach(), "dy": dy, "y": torch.empty(shape, device=dev, dtype=tdt),
                    "dx": torch.empty(shape, device=dev, dtype=tdt), "stats": torch.empty(2, c, device=dev),
                    "grads": torch.empty(2, 2, c, device=dev), "ws": torch.zeros(wsb, dtype=torch.uint8, device=dev),
                    "wsb": wsb, "c": c, "m": m, "chlast": chlast,
                    "cws": torch.zeros(lib.micn_cl_workspace_bytes(1, c, m), dtype=torch.uint8, device=dev),
                    "gp": (ctypes.c_void_p * 2)(*[n_.weight.data_ptr() for n_ in mod.norms]),
                    "bp": (ctypes.c_void_p * 2)(*[n_.bias.data_ptr() for n_ in mod.norms])})
        elems += c * m
    styles = torch.tensor([1], device=dev)

    def ours_module():
        for mod, x, dy in calls:
            mod(x, styles).backward(dy)

    def torch_module():
        for mod, x, dy in calls:
            w, b = mod.norms[1].weight, mod.norms[1].bias
            torch.stack([F.instance_norm(x[i].unsqueeze(0), None, None, w, b, True, 0.1, 1e-5).squeeze(0)
                         for i in range(x.shape[0])]).backward(dy)

    def ours_cabi(stream=stream):
        for r in raw:
            c, m = r["c"], r["m"]
            if r["chlast"]:  # token-major input: the channels-last kernels, no transposing copy
                xcl = r["x"].permute(0, 2, 3, 4, 1)
                rc = lib.micn_fwd_cl(xcl.data_ptr(), r["y"].data_ptr(), r["gp"], r["bp"], 2, styles.data_ptr(),
                                     r["stats"][0].data_ptr(), r["stats"][1].data_ptr(), 1, c, m, code, 1e-5,
                                     r["cws"].data_ptr(), r["cws"].numel(), stream)
                rc = rc or lib.micn_bwd_cl(r["dy"].data_ptr(), xcl.data_ptr(), r["gp"], r["bp"], 2, styles.data_ptr(),
                                           r["stats"][0].data_ptr(), r["stats"][1].data_ptr(), r["dx"].data_ptr(),
                                           r["grads"][0].data_ptr(), r["grads"][1].data_ptr(), 1, c, m, code,
                                           r["cws"].data_ptr(), r["cws"].numel(), stream)
            else:
                x = r["x"]
                rc = lib.micn_fwd(x.data_ptr(), r["y"].data_ptr(), None, r["gp"], r["bp"], 2, styles.data_ptr(),
                                  r["stats"][0].data_ptr(), r["stats"][1].data_ptr(), 1, c, m, c * m, m, code, 0, 0.01,
                                  1e-5, r["ws"].data_ptr(), r["wsb"], stream)
                rc = rc or lib.micn_bwd(r["dy"].data_ptr(), x.data_ptr(), None, r["gp"], r["bp"], 2, styles.data_ptr(),
                                        r["stats"][0].data_ptr(), r["stats"][1].data_ptr(), r["dx"].data_ptr(), None,
                                        r["grads"][0].data_ptr(), r["grads"][1].data_ptr(), 1, c, m, c * m, m, code, 0,
                                        0.01, r["ws"].data_ptr(), r["wsb"], stream)
            if rc:
                raise RuntimeError(f"model call list: rc={rc}")

    # the same 62 launches captured once into a CUDA graph and replayed (the calls keep no per-launch state on the
    # host, micn.h): what the GPU needs when the host is out of the way
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ours_cabi(side.cuda_stream)  # (first launches outside the capture: function attributes are set once)
        with torch.cuda.graph(graph, stream=side):
            ours_cabi(side.cuda_stream)
    torch.cuda.current_stream().wait_stream(side)

    out = {}
    for key, fn in (("ours_cabi_ms", ours_cabi), ("ours_cabi_graph_ms", graph.replay), ("ours_module_ms", ours_module),
                    ("torch_module_ms", torch_module)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record()
        for _ in range(reps):
            fn()
        b_.record()
        torch.cuda.synchronize()
        out[key] = a_.elapsed_time(b_) / reps
    es = torch.empty(0, dtype=tdt).element_size()
    out.update({"calls": len(calls), "elements_per_forward": elems, "algorithmic_bytes": 5 * elems * es,
                "ours_cabi_gbps": 5 * elems * es / (out["ours_cabi_ms"] * 1e-3) / 1e9,
                "ours_cabi_graph_gbps": 5 * elems * es / (out["ours_cabi_graph_ms"] * 1e-3) / 1e9,
                "patch_voxels_per_s_norm_only": 96 ** 3 / (out["ours_cabi_ms"] * 1e-3),
                "speedup_vs_torch_gpu_module_level": out["torch_module_ms"] / out["ours_module_ms"],
                "speedup_vs_torch_gpu_device_level": out["torch_module_ms"] / out["ours_cabi_ms"],
                "what": "all 31 instance_cond calls of one C-Swin-UNETR (f=48, 96^3, B=1) step, fwd+bwd; norm-only time, "
                        "convs / attention not run"})
    return out


def run_next_rows(pkg, dev, tdt, n, c, s, peak):
    """Measurement for the SURVEY.md 8(f) rows that are built: (1) the decoders' plain instance norm, (3) the PReLU
    epilogue of C-UNet's ADN, both on the headline activation through the nn.Module + autograd (CUDA events over
    back-to-back fwd+bwd, rotating inputs), and (4) the sliding-window driver's own overhead (windows/s with a
    pass-through predictor on a 192 x 192 x 160 volume: enumeration, gather, blend)."""
    import torch

    es = torch.empty(0, dtype=tdt).element_size()
    E = n * c * s ** 3
    R = 3
    xs = [(torch.randn(n, c, s, s, s, device=dev) * 2 + 1).to(tdt).requires_grad_(True) for _ in range(R)]
    dy = torch.randn(n, c, s, s, s, device=dev).to(tdt)
    styles = (torch.arange(n, device=dev) % 2).to(torch.int64)
    plain = pkg.FastInstanceNorm3d(c, affine=True).to(dev)
    cond = pkg.FastConditionalInstanceNorm3d(num_styles=2, num_features=c).to(dev)
    act = torch.nn.PReLU(init=0.25).to(dev)

    def timed(fn, reps=30):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(reps):
            fn(i)
        b_.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b_) / reps * 1e3

    t_plain = timed(lambda i: plain(xs[i % R]).backward(dy))
    t_prelu = timed(lambda i: cond.forward_fused(xs[i % R], styles, "lrelu", slope=act.weight).backward(dy))
    # dual-norm epilogue (SURVEY.md 8(d) bytes table, row 3): y = lrelu(norm2(a) + norm3(b)), forward 3*E*s + backward
    # 5*E*s, raw C-ABI launches back to back on rotating buffers, against the three-call composition it replaces
    dual = run_dual_leg(pkg, dev, tdt, n, c, s, peak)
    vol = torch.randn(1, 1, 192, 192, 160, device=dev)
    nwin = len(pkg.window_slices((192, 192, 160), (96, 96, 96), 0.5))
    t_sw = timed(lambda i: pkg.sliding_window_inference(vol, 96, 4, lambda w, modalities=None: w, overlap=0.5,
                                                        modalities=torch.tensor([1], device=dev)), reps=3)
    out = {"plain_instance_norm": {"us_per_fwd_bwd": t_plain, "gbps": 5 * E * es / t_plain * 1e-3,
                                   "frac": 5 * E * es / t_plain * 1e-3 / peak},
           "prelu_epilogue": {"us_per_fwd_bwd": t_prelu, "gbps": 5 * E * es / t_prelu * 1e-3,
                              "frac": 5 * E * es / t_prelu * 1e-3 / peak,
                              "note": "slope gradient accumulated by the backward kernel (one partial per CTA), summed by one small torch op"},
           "sliding_window_driver": {"windows": nwin, "ms_per_volume": t_sw * 1e-3, "windows_per_s": nwin / (t_sw * 1e-6),
                                     "note": "pass-through predictor: the driver's own gather / blend cost, 1 GPU"},
           "dual_norm_epilogue": dual,
           "what": f"module-level (nn.Module + autograd, host overhead included) on {n}x{c}x{s}^3"}
    return out


def run_dual_leg(pkg, dev, tdt, n, c, s, peak, reps=40):
    """micn_fwd_dual + micn_bwd_dual on the headline activation (UnetResBlock's downsample branch: C-Swin-UNETR encoder1),
    and the composition it replaces: micn_fwd(b) + micn_fwd(a, ADD_LRELU) + micn_bwd(ADD_LRELU) + micn_bwd(b)."""
    import torch

    lib = pkg._lib.lib()
    es = torch.empty(0, dtype=tdt).element_size()
    code = {torch.bfloat16: 1, torch.float32: 0, torch.float16: 2}[tdt]
    m = s ** 3
    E = n * c * m
    S = 2
    R = max(3, int(700e6 // (6 * E * es)) + 1)
    A = [(torch.randn(n, c, m, device=dev) * 2 + 1).to(tdt) for _ in range(R)]
    B = [(torch.randn(n, c, m, device=dev) * 0.5).to(tdt) for _ in range(R)]
    DY = [torch.randn(n, c, m, device=dev).to(tdt) for _ in range(R)]
    Y, DA, DB, RES = (torch.empty_like(A[0]) for _ in range(4))
    stats = torch.empty(4, n * c, device=dev)
    par = [1 + 0.3 * torch.randn(S, c, device=dev), 0.3 * torch.randn(S, c, device=dev),
           1 + 0.3 * torch.randn(S, c, device=dev), 0.3 * torch.randn(S, c, device=dev)]
    arr = [(ctypes.c_void_p * S)(*[t[k].data_ptr() for k in range(S)]) for t in par]
    grads = torch.empty(4, S, c, device=dev)
    styles = (torch.arange(n, device=dev) % S).to(torch.int64)
    wsb = lib.micn_workspace_bytes(n, c, m, code, S)
    ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    sp = [stats[k].data_ptr() for k in range(4)]
    gp = [grads[k].data_ptr() for k in range(4)]

    def fused(i):
        rc = lib.micn_fwd_dual(A[i].data_ptr(), B[i].data_ptr(), Y.data_ptr(), arr[0], arr[1], arr[2], arr[3], S,
                               styles.data_ptr(), sp[0], sp[1], sp[2], sp[3], n, c, m, code, 0.01, 1e-5, ws.data_ptr(), wsb,
                               stream)
        j = (i + 1) % R
        rc = rc or lib.micn_bwd_dual(DY[j].data_ptr(), A[j].data_ptr(), B[j].data_ptr(), arr[0], arr[1], arr[2], arr[3], S,
                                     styles.data_ptr(), sp[0], sp[1], sp[2], sp[3], DA.data_ptr(), DB.data_ptr(), gp[0],
                                     gp[1], gp[2], gp[3], n, c, m, code, 0.01, ws.data_ptr(), wsb, stream)
        if rc:
            raise RuntimeError(f"dual leg: rc={rc}")

    def composed(i):
        rc = lib.micn_fwd(B[i].data_ptr(), RES.data_ptr(), None, arr[2], arr[3], S, styles.data_ptr(), sp[2], sp[3], n, c, m,
                          c * m, m, code, 0, 0.01, 1e-5, ws.data_ptr(), wsb, stream)
        rc = rc or lib.micn_fwd(A[i].data_ptr(), Y.data_ptr(), RES.data_ptr(), arr[0], arr[1], S, styles.data_ptr(), sp[0],
                                sp[1], n, c, m, c * m, m, code, 2, 0.01, 1e-5, ws.data_ptr(), wsb, stream)
        j = (i + 1) % R
        rc = rc or lib.micn_bwd(DY[j].data_ptr(), A[j].data_ptr(), Y.data_ptr(), arr[0], arr[1], S, styles.data_ptr(), sp[0],
                                sp[1], DA.data_ptr(), RES.data_ptr(), gp[0], gp[1], n, c, m, c * m, m, code, 2, 0.01,
                                ws.data_ptr(), wsb, stream)
        rc = rc or lib.micn_bwd(RES.data_ptr(), B[j].data_ptr(), None, arr[2], arr[3], S, styles.data_ptr(), sp[2], sp[3],
                                DB.data_ptr(), None, gp[2], gp[3], n, c, m, c * m, m, code, 0, 0.01, ws.data_ptr(), wsb,
                                stream)
        if rc:
            raise RuntimeError(f"dual leg (composition): rc={rc}")

    out = {}
    if not lib.micn_dual_supported(n, c, m, code, 0) or not lib.micn_dual_supported(n, c, m, code, 1):
        return {"unsupported": True}
    for key, fn in (("fused", fused), ("composed", composed)):
        for i in range(3):
            fn(i % R)
        torch.cuda.synchronize()
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record()
        for i in range(reps):
            fn(i % R)
        b_.record()
        torch.cuda.synchronize()
        out[key + "_us"] = a_.elapsed_time(b_) / reps * 1e3
    alg = 8 * E * es
    out.update({"algorithmic_bytes": alg, "gbps": alg / out["fused_us"] * 1e-3, "frac": alg / out["fused_us"] * 1e-3 / peak,
                "composed_traffic_bytes": 12 * E * es, "speedup_vs_composed": out["composed_us"] / out["fused_us"],
                "path": int(pkg._lib.get_option("last_path")),
                "what": f"lrelu(norm2(a) + norm3(b)) fwd (3*E*s) + bwd (5*E*s) on {n}x{c}x{s}^3, one launch per direction; "
                        "composed = norm3 fwd, norm2+add+lrelu fwd, its bwd, norm3 bwd (12*E*s of traffic, 4 launches)"})
    return out


def ncu_traffic(args, n, c, s):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the backward kernel of this workload, from the committed
    `ncu --set full` captures (profiles/traffic.json, written from profiles/r02_*.txt); None when this workload has none."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            v = json.load(f).get(f"bwd_{args.dtype}_{n}x{c}x{s}")
        return float(v) if v is not None else None
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.thread = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.rows.append((time.perf_counter(), line.strip()))
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = [r for t, r in self.rows if (t0 is None or t >= t0 - 0.05) and (t1 is None or t <= t1 + 0.15)] or \
               [r for _, r in self.rows]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [v.strip() for v in r.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU reference (reference arm / cpu_baseline)
def workload_config(n, c, s, dtype_name, world, epilogue="none"):
    """The `config` object, identical in both arms (the driver compares them)."""
    return {"workload": f"instance_cond fwd+bwd, {n}x{c}x{s}^3 {dtype_name} per GPU (C-Swin-UNETR f=48 encoder1 norm; "
                        f"BASELINE.json configs[1] hot path), epilogue={epilogue}",
            "global_batch": n * world, "parallelism": f"dp{world}"}


def load_reference_module():
    """The reference's own hot-path module, unmodified: the git-ignored copy oracle/_ref/conditional_instance_norm.py
    (baseline/make_ref.py copies it from /root/reference/networks/norms/ in the build container; it needs only torch and
    travels to the GPU box with the snapshot).  None when the copy is absent."""
    import importlib.util

    path = os.path.join(ROOT, "oracle", "_ref", "conditional_instance_norm.py")
    if not os.path.isfile(path):
        return None
    spec = importlib.util.spec_from_file_location("micn_reference_conditional_instance_norm", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cpu_reference_time(n, c, s, dtype_name, budget_s, warmup, steps=None):
    """Forward + autograd backward of the reference's ConditionalInstanceNorm3d (conditional_instance_norm.py:59-68) on
    the host cores, all threads; falls back to the oracle's torch-CPU port of the same call sequence (per-sample
    F.instance_norm + torch.stack) when oracle/_ref is absent.  Returns (seconds per fwd+bwd, iterations, cores, kind)."""
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dt = {"bf16": torch.bfloat16, "fp32": torch.float32, "fp16": torch.float16}[dtype_name]
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(n, c, s, s, s, generator=g) * 2 + 1).to(dt)
    dy = torch.randn(n, c, s, s, s, generator=g).to(dt)
    w = [1 + 0.3 * torch.randn(c, generator=g) for _ in range(2)]
    b = [0.3 * torch.randn(c, generator=g) for _ in range(2)]
    styles = [i % 2 for i in range(n)]
    ref = load_reference_module()
    if ref is not None:
        kind = "reference"
        mod = ref.ConditionalInstanceNorm3d(num_styles=2, num_features=c)
        with torch.no_grad():
            for k in range(2):
                mod.norms[k].weight.copy_(w[k])
                mod.norms[k].bias.copy_(b[k])
        xr = x.clone().requires_grad_(True)

        def one():
            xr.grad = None
            mod.zero_grad(set_to_none=True)
            mod(xr, styles).backward(dy)
    else:
        kind = "port"
        from oracle import micn_oracle as O

        def one():
            O.port_fwd_bwd(x, dy, styles, w, b)
    for _ in range(max(1, warmup)):
        one()
    times = []
    t_start = time.perf_counter()
    while True:
        t0 = time.perf_counter()
        one()
        times.append(time.perf_counter() - t0)
        if steps is not None:
            if len(times) >= steps:
                break
        elif len(times) >= 3 and time.perf_counter() - t_start > budget_s:
            break
    return sum(times) / len(times), len(times), cores, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, c, s, m = workload(args)
    es = DT_BYTES[args.dtype]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    steps = min(args.steps, 100)  # bounded sample: each step is one full fwd+bwd of the workload on the CPU (~30 ms)
    sec, iters, cores, kind = cpu_reference_time(n, c, s, args.dtype, 0, min(args.warmup, 20), steps=steps)
    gbps = 5.0 * n * c * m * es / sec / 1e9
    what = ("the reference's unmodified ConditionalInstanceNorm3d (oracle/_ref copy) forward + autograd backward"
            if kind == "reference" else "the oracle's torch-CPU port of the reference call sequence")
    line = {
        "impl": "reference", "metric": METRIC, "value": gbps, "unit": UNIT, "n_gpus": args.gpus, "steps": iters,
        "warmup": min(args.warmup, 20), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": workload_config(n, c, s, args.dtype, world, args.epilogue),
        "cpu_baseline": {"value": gbps, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{iters} full fwd+bwd passes of the workload on the host CPU, {cores} threads: {what}"},
        "e2e": {"value": gbps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "voxels_per_s": n * m / sec,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ model-level legs
def run_model_legs(args, pkg, dev, world, rank):
    """BASELINE.json metric, second half ("C-SwinUNETR voxels/s @1-8 GPU") and configs[0], [2], [4]: whole steps of the
    reference's own nets, reference norms vs this repo's drop-in, on the same GPU(s).  Returns extra keys for the line."""
    which = args.model_steps
    if which == "none":
        return {}
    if which == "auto":
        legs = ["swin_unetr", "sliding_window"] + (["unetr", "unet_cpu"] if world == 1 else [])
    else:
        legs = [w for w in which.split(",") if w]
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    try:
        import model_bench as MB
    except Exception as e:  # noqa: BLE001
        return {"model_step": {"unavailable": repr(e)[:300]}}
    if not MB.reference_available():
        return {"model_step": {"unavailable": "no importable copy of the reference's networks/ package (baseline/_ref "
                                              "is made by baseline/make_ref.py where /root/reference exists)"}}
    out = {}
    K, W = args.model_step_iters, 3
    try:
        if "swin_unetr" in legs:
            out["model_step"] = MB.model_step_leg("swin_unetr", pkg, dev, world, rank, K, W, batch=1)
        if "unetr" in legs:
            out["model_step_unetr"] = MB.model_step_leg("unetr", pkg, dev, world, rank, K, W, batch=4)
        if "unet" in legs:
            out["model_step_unet"] = MB.model_step_leg("unet", pkg, dev, world, rank, K, W, batch=1)
        if "sliding_window" in legs:
            out["sliding_window"] = MB.sliding_window_leg(pkg, dev, world, rank)
        if "unet_cpu" in legs and rank == 0:
            out["model_step_unet_cpu"] = MB.unet_cpu_leg()
    except Exception as e:  # noqa: BLE001 - the headline line must survive a failing extra leg
        out["model_legs_error"] = repr(e)[:400]
    return out


# ------------------------------------------------------------------------------------------------ ours
def run_ours(args):
    import torch
    import torch.distributed as dist

    import mi_seg_b200 as pkg

    lib = pkg._lib.lib()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    sampler = ClockSampler(local)
    if rank == 0:  # started first: nvidia-smi needs up to seconds for its first answer on an 8-GPU box
        sampler.start()
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, c, s, m = workload(args)
    S = 2
    # One stream carries every kernel of this process, so the persistent flat kernels can be launched as programmatic
    # dependents (griddepcontrol: their launch latency and barrier set-up overlap the tail of the kernel before them) instead
    # of cooperatively; the library's default stays the cooperative launch, which is also safe when two streams launch
    # flat kernels at the same time (include/micn.h, "flat_pdl").
    pdl = not os.environ.get("MICN_BENCH_NO_PDL")
    if pdl:
        pkg._lib.set_option("flat_pdl", 1)
        pkg._lib.set_option("flat_coop", 0)
    for kv in os.environ.get("MICN_BENCH_OPTS", "").split(","):  # (debug knob: library options, name=value)
        if "=" in kv:
            pkg._lib.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    es = DT_BYTES[args.dtype]
    tdt = {"bf16": torch.bfloat16, "fp32": torch.float32, "fp16": torch.float16}[args.dtype]
    code = {"fp32": 0, "bf16": 1, "fp16": 2}[args.dtype]
    epi = {"none": 0, "lrelu": 1}[args.epilogue]
    E = n * c * m
    bytes_fwd, bytes_bwd = 2 * E * es, 3 * E * es

    # rotating buffer sets, footprint >> L2 (126 MB): every launch reads HBM-cold inputs
    R = max(4, int(600e6 // (4 * E * es)) + 1)
    torch.manual_seed(rank)
    xs = [(torch.randn(n, c, m, device=dev) * 2 + 1).to(tdt) for _ in range(R)]
    dys = [torch.randn(n, c, m, device=dev).to(tdt) for _ in range(R)]
    ys = [torch.empty_like(xs[0]) for _ in range(R)]
    dxs = [torch.empty_like(xs[0]) for _ in range(R)]
    means = [torch.empty(n * c, device=dev) for _ in range(R)]
    rstds = [torch.empty(n * c, device=dev) for _ in range(R)]
    gamma = 1 + 0.3 * torch.randn(S, c, device=dev)
    beta = 0.3 * torch.randn(S, c, device=dev)
    styles = (torch.arange(n, device=dev) % S).to(torch.int64)
    # dgamma, dbeta: one bucket per step for the all-reduce, double-buffered so that the collective of step i (a few
    # microseconds of NCCL on one SM) runs while step i+1 computes - the way DDP overlaps its buckets with the
    # rest of backward; a bucket is waited for before it is written again and before the clock stops
    # (one bucket per buffer set, so that a captured step always writes the same bucket)
    grads2 = [torch.empty(2, S, c, device=dev) for _ in range(R)]
    pending = [None] * R
    # N > 1, default: the exchange is fused into the backward kernel over NVLink peer memory (no NCCL call in the step)
    px, px_note = None, None
    if world > 1 and args.collective == "fused":
        try:
            px = pkg.PeerExchange(c, S, dev)
            px_note = f"peer buffers mapped with {px.how}"
        except Exception as e:  # noqa: BLE001 - no peer access on this box: NCCL it is
            px_note = "fused exchange unavailable (" + repr(e)[:200] + "): NCCL all-reduce instead"
    overlap = world > 1 and px is None and not os.environ.get("MICN_BENCH_SYNC_ALLREDUCE")
    if overlap:  # leave one SM to NCCL (the flat kernels are persistent, one CTA per SM)
        pkg._lib.set_option("flat_grid", torch.cuda.get_device_properties(dev).multi_processor_count - 1)
    wsb = lib.micn_workspace_bytes(n, c, m, code, S)
    ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
    gp = (ctypes.c_void_p * S)(*[gamma[k].data_ptr() for k in range(S)])
    bp = (ctypes.c_void_p * S)(*[beta[k].data_ptr() for k in range(S)])
    stream = torch.cuda.current_stream().cuda_stream

    def fwd(i):
        rc = lib.micn_fwd(xs[i].data_ptr(), ys[i].data_ptr(), None, gp, bp, S, styles.data_ptr(), means[i].data_ptr(),
                          rstds[i].data_ptr(), n, c, m, c * m, m, code, epi, 0.01, 1e-5, ws.data_ptr(), wsb, stream)
        if rc:
            raise RuntimeError(f"micn_fwd rc={rc}")

    def bwd(i, grads):
        rc = lib.micn_bwd(dys[i].data_ptr(), xs[i].data_ptr(), None, gp, bp, S, styles.data_ptr(), means[i].data_ptr(),
                          rstds[i].data_ptr(), dxs[i].data_ptr(), None, grads[0].data_ptr(), grads[1].data_ptr(),
                          n, c, m, c * m, m, code, epi, 0.01, ws.data_ptr(), wsb, stream)
        if rc:
            raise RuntimeError(f"micn_bwd rc={rc}")

    def bwd_fused(i, grads, mode=2):
        # mode 2 (MICN_FOLD_PREVIOUS): `grads` receives the all-reduced gradients of the PREVIOUS step - the bucket completes
        # one step behind, as with the asynchronous NCCL scheme; drain() folds the last step's with micn_allreduce_fold
        rc = lib.micn_bwd_allreduce(dys[i].data_ptr(), xs[i].data_ptr(), None, gp, bp, S, styles.data_ptr(), means[i].data_ptr(),
                                    rstds[i].data_ptr(), dxs[i].data_ptr(), None, grads[0].data_ptr(), grads[1].data_ptr(),
                                    n, c, m, c * m, m, code, epi, 0.01, ws.data_ptr(), wsb, px.ptrs, rank, world, mode, stream)
        if rc:
            raise RuntimeError(f"micn_bwd_allreduce rc={rc}")

    def step(i):
        # forward on set i, backward on the NEXT set (its statistics come from an earlier forward of the
        # same data): neither kernel finds its inputs in L2 from the launch before it
        b = i % R
        if pending[b] is not None:  # the bucket's previous all-reduce must be through before it is overwritten
            pending[b].wait()
            pending[b] = None
        launch_pair(b)
        if world > 1 and px is None and not os.environ.get("MICN_BENCH_NO_ALLREDUCE"):  # (debug knob: isolate the collective)
            if overlap:
                pending[b] = dist.all_reduce(grads2[b], async_op=True)
            else:
                dist.all_reduce(grads2[b])

    def launch_pair(b):  # replaced by a graph replay under --launch graph
        fwd(b)
        (bwd_fused if px is not None else bwd)((b + 1) % R, grads2[b])

    def drain():
        if px is not None and not os.environ.get("MICN_BENCH_SKIP_FOLD"):  # the last step's exchange (every earlier one was folded by the step after it)
            rc = lib.micn_allreduce_fold(px.ptrs, rank, world, c, S, grads2[0][0].data_ptr(), grads2[0][1].data_ptr(), stream)
            if rc:
                raise RuntimeError(f"micn_allreduce_fold rc={rc}")
        for b in range(R):
            if pending[b] is not None:
                pending[b].wait()
                pending[b] = None

    for i in range(R):  # statistics for every set
        fwd(i)
    torch.cuda.synchronize()

    # ---- N > 1: prove once that the collective delivers the right numbers (sum over ranks of the rank-local dgamma/dbeta)
    allreduce_check = None
    if world > 1 and not os.environ.get("MICN_BENCH_SKIP_CHECK"):  # (debug knob for timing-only experiments)
        bwd(0, grads2[0])
        local = grads2[0].clone()
        if px is not None:
            bwd_fused(0, grads2[0], 1)  # synchronous fold (every rank makes this call: the exchange is a collective)
        else:
            dist.all_reduce(grads2[0])
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local)
        expect = torch.stack(parts).double().sum(0)
        allreduce_check = float((grads2[0].double() - expect).abs().max() / expect.abs().max())
        if not allreduce_check < 5e-3:
            raise RuntimeError(f"all-reduced dgamma/dbeta differ from the sum of the rank-local gradients: {allreduce_check}")

    # ---- optional: one step per buffer set captured into a CUDA graph and replayed (what a captured training step does)
    graphs = None
    bwd(0, grads2[0])  # (first launch of every kernel outside any capture: function attributes are set once)
    if args.launch == "graph":
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        graphs = []
        with torch.cuda.stream(side):
            stream = side.cuda_stream
            for i in range(R):
                g_ = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_, stream=side):
                    fwd(i)
                    (bwd_fused if px is not None else bwd)((i + 1) % R, grads2[i])
                graphs.append(g_)
        torch.cuda.current_stream().wait_stream(side)
        stream = torch.cuda.current_stream().cuda_stream

        def launch_pair(b):  # noqa: F811 - the collective (N > 1) is still issued from the host after every replay
            graphs[b].replay()

    graph_all = None
    if args.launch == "graphR":  # (experiment, never the default: R consecutive steps in one graph)
        if (world > 1 and px is None) or args.steps % R:
            raise SystemExit(f"--launch graphR: --steps must be a multiple of R = {R}, and N > 1 needs --collective fused")
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            stream = side.cuda_stream
            graph_all = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph_all, stream=side):
                for i in range(R):
                    fwd(i)
                    (bwd_fused if px is not None else bwd)((i + 1) % R, grads2[i])
        torch.cuda.current_stream().wait_stream(side)
        stream = torch.cuda.current_stream().cuda_stream

    def run_steps(k):
        if graph_all is None:
            for i in range(k):
                step(i)
        else:
            for _ in range((k + R - 1) // R):
                graph_all.replay()

    # untimed pre-warm, independent of --warmup: the GPU drops to its idle clocks within a fraction of a second of
    # inactivity (the sampler start above is one), and a 20-step region lasts 1.7 ms - shorter than the clock ramp
    # (a fixed count: with the fused exchange every rank must make the same sequence of calls.  Deliberately short - 16 ms:
    # half a second of this load before the clock starts makes the step 4 % slower (84.5 instead of 80.8 us at 2 GPUs), the
    # same burst-vs-sustained gap MEASURED_PEAKS.json records for its own figures, and the roofline's denominator is the
    # burst copy rate)
    prewarm = max(200, args.warmup, 10 * args.steps)
    t_load0 = time.perf_counter()
    run_steps(prewarm)
    drain()
    torch.cuda.synchronize()
    launches_region = None
    region_ms = []
    t_wall0 = time.perf_counter()
    for r in range(max(1, args.regions)):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        launches0 = pkg._lib.get_option("launches")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_steps(args.steps)
        drain()  # every step's collective completes inside the timed region
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        region_ms.append(e0.elapsed_time(e1))
        if launches_region is None:
            launches_region = (pkg._lib.get_option("launches") - launches0) if graphs is None and graph_all is None else 2 * args.steps
    t_wall1 = time.perf_counter()
    launches = launches_region
    if px is not None and os.environ.get("MICN_BENCH_XCHG_DBG"):  # (bring-up: wait statistics of the folds, see micn_flat.cuh)
        torch.cuda.synchronize()
        w = px.local[16:40].view(torch.int32).tolist()
        print(f"[rank {rank}] fold waits: max {w[0]} ns, count {w[1]}, total {w[2]} ns, own-rank records {w[3]}, peer records {w[4]}",
              file=sys.stderr, flush=True)
    per_rank_us = None
    if world > 1:  # max over ranks, per region (the spread between the GPUs of the box is reported beside it)
        mine = torch.tensor([statistics.median(region_ms) / args.steps * 1e3], device=dev, dtype=torch.float64)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank_us = [round(float(v.item()), 2) for v in allr]
        t = torch.tensor(region_ms, device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        region_ms = [float(v) for v in t.tolist()]
    ms_total = statistics.median(region_ms)

    # per-kernel durations (roofline), same loop with an event after every launch
    K2 = min(args.steps, 200)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K2)]
    for i in range(K2):
        evs[i][0].record()
        fwd(i % R)
        evs[i][1].record()
        bwd((i + 1) % R, grads2[0])
        evs[i][2].record()
    torch.cuda.synchronize()
    fwd_us = [e[0].elapsed_time(e[1]) * 1e3 for e in evs]
    bwd_us = [e[1].elapsed_time(e[2]) * 1e3 for e in evs]
    fwd_pair_avg, bwd_pair_avg = sum(fwd_us) / K2, sum(bwd_us) / K2

    # An event between two launches exposes ~3 us of launch latency that back-to-back launches overlap (the two
    # event-pair figures add up to more than the measured step).  The roofline figure is therefore the kernel's
    # average duration over a region of K2 back-to-back launches of that kernel alone, one event pair around the
    # region, rotating buffer sets (each launch finds its inputs in HBM, not L2).
    def region_us(fn):
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record()
        for i in range(K2):
            fn(i)
        b_.record()
        torch.cuda.synchronize()
        return a_.elapsed_time(b_) * 1e3 / K2

    bwd_avg = region_us(lambda i: bwd(i % R, grads2[0]))
    fwd_avg = region_us(lambda i: fwd(i % R))
    # clocks: the timed regions last 8 ms in all, nvidia-smi answers every 50 ms at best - so the same steps run on for
    # ~0.35 s after the measurements and the sampler's answers between the start of the pre-warm and the end of this
    # observation loop are reported (SM clock under exactly this load, throttle reasons)
    for i in range(4000):
        step(i)
    drain()
    torch.cuda.synchronize()
    t_region_end = time.perf_counter()
    clocks = sampler.stop(t_load0, t_region_end) if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "pre-warm + timed regions + roofline regions + 4000 more steps of the same load (~0.4 s)"

    ms_step = ms_total / args.steps
    value = world * (bytes_fwd + bytes_bwd) / (ms_step * 1e-3) / 1e9

    # ---- e2e: host buffers through the C ABI (micn_fwd_bwd_host), blocking call, wall clock
    e2e = None
    if not args.no_e2e:
        scratch_b = lib.micn_host_scratch_bytes(n, c, m, code, S, 1)
        scratch = torch.empty(scratch_b, dtype=torch.uint8, device=dev)
        hx = xs[0].cpu().pin_memory()
        hdy = dys[0].cpu().pin_memory()
        hy = torch.empty_like(hx).pin_memory()
        hdx = torch.empty_like(hx).pin_memory()
        hg, hb = gamma.cpu().contiguous(), beta.cpu().contiguous()
        hst = styles.cpu()
        hdg, hdb = torch.empty(S, c), torch.empty(S, c)

        def host_step():
            rc = lib.micn_fwd_bwd_host(hx.data_ptr(), hdy.data_ptr(), hy.data_ptr(), hdx.data_ptr(), hg.data_ptr(),
                                       hb.data_ptr(), S, hst.data_ptr(), hdg.data_ptr(), hdb.data_ptr(), n, c, m, code,
                                       epi, 0.01, 1e-5, scratch.data_ptr(), scratch_b)
            if rc:
                raise RuntimeError(f"micn_fwd_bwd_host rc={rc}")

        ke = max(3, min(args.steps, 20))
        for _ in range(3):
            host_step()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            host_step()
        sec = (time.perf_counter() - t0) / ke
        if world > 1:
            t = torch.tensor([sec], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        # check the host path against the device path once (same data as set 0)
        torch.cuda.synchronize()
        def _rel(a, b):
            return float((a.float() - b.float()).abs().max() / b.float().abs().max())
        tol = 1e-5 if args.dtype == "fp32" else 2e-2
        ok = _rel(hy.to(dev), ys[0]) < tol and _rel(hdx.to(dev), dxs[0]) < tol
        small = 2 * S * c * 4 + n * 8
        e2e = {"value": world * (bytes_fwd + bytes_bwd) / sec / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": 2 * E * es + small, "d2h_bytes_per_step": 2 * E * es + 2 * S * c * 4,
               "ms_per_step": sec * 1e3, "steps": ke, "timer": "host wall clock around the blocking C-ABI call",
               "api": "micn_fwd_bwd_host (pinned host x, dy -> host y, dx, dgamma, dbeta)",
               "matches_device_path": ok}

    # ---- like-for-like GPU competitor: the reference's call sequence on the same B200 through PyTorch
    torch_gpu = None
    if rank == 0 and not args.no_torch_ref:
        import torch.nn.functional as F
        xg = xs[0].reshape(n, c, s, s, s).detach().requires_grad_(True)
        dyg = dys[0].reshape(n, c, s, s, s)
        wg = [gamma[k].clone().requires_grad_(True) for k in range(S)]
        bg = [beta[k].clone().requires_grad_(True) for k in range(S)]
        st_host = styles.tolist()

        def ref_step():
            y = torch.stack([F.instance_norm(xg[i].unsqueeze(0), None, None, wg[st_host[i]], bg[st_host[i]], True, 0.1,
                                             1e-5).squeeze(0) for i in range(n)])
            y.backward(dyg)
            xg.grad = None
        for _ in range(3):
            ref_step()
        torch.cuda.synchronize()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            ref_step()
        b_.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b_) / 10
        torch_gpu = {"ms_per_step": ms, "value": (bytes_fwd + bytes_bwd) / (ms * 1e-3) / 1e9, "unit": UNIT,
                     "what": "per-sample F.instance_norm + torch.stack + autograd on the same GPU (PyTorch/ATen)"}

    # ---- model-level legs (every rank takes part: DDP / sharded windows); see baseline/model_bench.py
    model_legs = run_model_legs(args, pkg, dev, world, rank)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    bwd_gbps = bytes_bwd / (bwd_avg * 1e-6) / 1e9
    fwd_gbps = bytes_fwd / (fwd_avg * 1e-6) / 1e9
    roofline = {"bound": "hbm", "kernel": "micn_bwd (backward, 3*E*s algorithmic bytes per launch)",
                "achieved": bwd_gbps, "peak": peak, "unit": "GB/s", "frac": bwd_gbps / peak,
                "traffic": ncu_traffic(args, n, c, s),
                "peak_source": peak_src, "avg_launch_us": bwd_avg, "bytes_per_launch": bytes_bwd,
                "timer": f"CUDA events around {K2} back-to-back launches of the kernel alone (rotating buffer sets)",
                "event_pair_per_launch_us": {"fwd": fwd_pair_avg, "bwd": bwd_pair_avg,
                                             "note": "an event between launches exposes ~3 us of launch latency each"},
                "fwd": {"achieved": fwd_gbps, "frac": fwd_gbps / peak, "avg_launch_us": fwd_avg,
                        "bytes_per_launch": bytes_fwd},
                "fwd_plus_bwd": {"achieved": (bytes_fwd + bytes_bwd) / ((fwd_avg + bwd_avg) * 1e-6) / 1e9,
                                 "frac": (bytes_fwd + bytes_bwd) / ((fwd_avg + bwd_avg) * 1e-6) / 1e9 / peak}}

    cpu_baseline = None
    if not args.no_cpu_baseline:
        sec, iters, cores, kind = cpu_reference_time(n, c, s, args.dtype, 10.0, 1)
        cpu_baseline = {"value": 5.0 * E * es / sec / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
                        "ms_per_step": sec * 1e3,
                        "sample": f"{iters} full fwd+bwd passes of the same {n}x{c}x{s}^3 {args.dtype} workload on the host "
                                  f"CPU, {cores} threads (" + ("the reference's unmodified ConditionalInstanceNorm3d, "
                                  "oracle/_ref copy" if kind == "reference" else "the oracle's torch-CPU port of the "
                                  "reference call sequence") + ")"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
        "data": "synthetic",
        "config": workload_config(n, c, s, args.dtype, world, args.epilogue),
        "notes": {"l2": f"{R} rotating buffer sets ({R * 4 * E * es / 1e6:.0f} MB) > 126 MB L2; backward reads a "
                        "different set than the forward before it",
                  "collective": ("d(gamma)/d(beta) all-reduced INSIDE the backward kernel: every rank stores its per-channel sums "
                                 "into every peer's memory over NVLink (micn_bwd_allreduce, MICN_FOLD_PREVIOUS; " + str(px_note) +
                                 "); the records of step i are folded in rank order at the start of step i+1's kernel "
                                 "(one-step lag, like an asynchronous bucket), the last step's by micn_allreduce_fold inside "
                                 "the timed region; no NCCL call in the step"
                                 if px is not None else
                                 "all_reduce(dgamma,dbeta) per step over NCCL, overlapped with the next step's kernels "
                                 "(double-buffered buckets, waited before reuse and before the clock stops; "
                                 "147 of 148 SMs run the norm kernels, one is left to NCCL)" if overlap else
                                 "all_reduce(dgamma,dbeta) per step over NCCL" if world > 1 else "none"),
                  "launch": args.launch + (", flat kernels launched as programmatic dependents (flat_pdl=1, flat_coop=0: one stream "
                                           "carries every kernel of this process)" if pdl else ", cooperative launches"),
                  "timing": f"{prewarm} untimed pre-warm steps, then {len(region_ms)} regions of {args.steps} steps "
                            "(CUDA events, barrier + synchronize on both sides, max over ranks per region); "
                            "ms_per_step = median region / steps"},
        "regions_ms": region_ms, "allreduce_check_rel_err": allreduce_check, "collective_note": px_note,
        "per_rank_us_per_step": per_rank_us,
        "voxels_per_s": world * n * m / (ms_step * 1e-3),
        "frac_of_peak": value / world / peak,
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clocks, "torch_gpu_reference": torch_gpu,
        "plan": {k: pkg._lib.get_option(k) for k in ("last_path", "last_cs", "last_slots", "last_grid")},
    }
    line.update(model_legs)
    if world == 1 and not args.no_model_calls:
        line["swin_unetr_norm_calls"] = run_model_calls(pkg, dev, tdt)
        line["next_rows"] = run_next_rows(pkg, dev, tdt, n, c, s, peak)
    if args.sweep and world == 1:
        del xs, dys, ys, dxs
        torch.cuda.empty_cache()
        line["sweep"] = run_sweep(lib, pkg, dev, peak, args.sweep_out)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
