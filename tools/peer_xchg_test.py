"""Multi-GPU check of micn_bwd_allreduce (fused NVLink exchange of d(gamma)/d(beta)) against micn_bwd + NCCL all-reduce.
torchrun --nproc-per-node N tools/peer_xchg_test.py      (also run by the gpu_scale script)"""
import ctypes
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mi_seg_b200 as pkg  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = pkg._lib.lib()
ok = True
for (n, c, sp, tdt, code) in ((1, 48, 96, torch.bfloat16, 1), (2, 6, 48, torch.float32, 0), (3, 10, 40, torch.bfloat16, 1)):
    S = 3
    m = sp ** 3
    torch.manual_seed(100 * rank + c)
    x = (torch.randn(n, c, m, device=dev) * 2 + 1).to(tdt)
    dy = torch.randn(n, c, m, device=dev).to(tdt)
    y, dx, dx2 = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    gam = 1 + 0.3 * torch.randn(S, c, device=dev)
    bet = 0.3 * torch.randn(S, c, device=dev)
    styles = ((torch.arange(n, device=dev) + rank) % S).to(torch.int64)
    stats = torch.empty(2, n * c, device=dev)
    g_ref, g_fused = torch.empty(2, S, c, device=dev), torch.empty(2, S, c, device=dev)
    wsb = lib.micn_workspace_bytes(n, c, m, code, S)
    ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
    ws2 = torch.zeros(wsb, dtype=torch.uint8, device=dev)
    gp = (ctypes.c_void_p * S)(*[gam[k].data_ptr() for k in range(S)])
    bp = (ctypes.c_void_p * S)(*[bet[k].data_ptr() for k in range(S)])
    px = pkg.PeerExchange(c, S, dev)
    if rank == 0:
        print("peer buffers via", px.how, getattr(px, "_symm_error", ""), flush=True)
    stream = torch.cuda.current_stream().cuda_stream
    rc = lib.micn_fwd(x.data_ptr(), y.data_ptr(), None, gp, bp, S, styles.data_ptr(), stats[0].data_ptr(), stats[1].data_ptr(),
                      n, c, m, c * m, m, code, 0, 0.01, 1e-5, ws.data_ptr(), wsb, stream)
    assert rc == 0

    def fused(mode=1):
        rc = lib.micn_bwd_allreduce(dy.data_ptr(), x.data_ptr(), None, gp, bp, S, styles.data_ptr(), stats[0].data_ptr(),
                                    stats[1].data_ptr(), dx2.data_ptr(), None, g_fused[0].data_ptr(), g_fused[1].data_ptr(),
                                    n, c, m, c * m, m, code, 0, 0.01, ws2.data_ptr(), wsb, px.ptrs, rank, world, mode, stream)
        assert rc == 0, rc

    rc = lib.micn_bwd(dy.data_ptr(), x.data_ptr(), None, gp, bp, S, styles.data_ptr(), stats[0].data_ptr(), stats[1].data_ptr(),
                      dx.data_ptr(), None, g_ref[0].data_ptr(), g_ref[1].data_ptr(), n, c, m, c * m, m, code, 0, 0.01,
                      ws.data_ptr(), wsb, stream)
    assert rc == 0
    dist.all_reduce(g_ref)
    for rep in range(5):  # both parities of the record buffers, several tags
        g_fused.fill_(float("nan"))
        fused()
        torch.cuda.synchronize()
        err = float((g_fused - g_ref).abs().max() / g_ref.abs().max())
        # (the plain call may take another path - e.g. the resident one - than the exchange's flat kernel: compare by value)
        same_dx = float((dx.float() - dx2.float()).abs().max() / dx.float().abs().max()) < (1e-5 if tdt == torch.float32 else 1e-2)
        # every rank must hold the same bits
        parts = [torch.empty_like(g_fused) for _ in range(world)]
        dist.all_gather(parts, g_fused)
        identical = all(torch.equal(parts[0], p_) for p_ in parts)
        good = err < 1e-5 and same_dx and identical
        ok = ok and good
        if rank == 0:
            print(f"shape {n}x{c}x{sp}^3 {tdt} rep {rep}: rel err vs NCCL {err:.2e}, dx matches {same_dx}, "
                  f"all ranks bit-identical {identical} -> {'ok' if good else 'FAIL'}", flush=True)
    # lagged mode: call k delivers the all-reduced gradients of call k-1; micn_allreduce_fold() the last call's.  The data
    # of consecutive calls differ (dy scaled by k), so a fold of the wrong call cannot pass.
    dy0 = dy.clone()
    for k in range(1, 5):
        dy.copy_((dy0.float() * k).to(tdt))
        g_fused.fill_(float("nan"))
        fused(2)
        torch.cuda.synchronize()
        if k > 1:
            err = float((g_fused - g_ref * (k - 1)).abs().max() / (g_ref * (k - 1)).abs().max())
            good = err < (1e-5 if tdt == torch.float32 else 2e-2)
            ok = ok and good
            if rank == 0:
                print(f"   lagged call {k}: delivers call {k - 1}, rel err {err:.2e} -> {'ok' if good else 'FAIL'}", flush=True)
    g_fused.fill_(float("nan"))
    assert lib.micn_allreduce_fold(px.ptrs, rank, world, c, S, g_fused[0].data_ptr(), g_fused[1].data_ptr(), stream) == 0
    torch.cuda.synchronize()
    err = float((g_fused - g_ref * 4).abs().max() / (g_ref * 4).abs().max())
    good = err < (1e-5 if tdt == torch.float32 else 2e-2)
    ok = ok and good
    if rank == 0:
        print(f"   micn_allreduce_fold: last call, rel err {err:.2e} -> {'ok' if good else 'FAIL'}", flush=True)
    dy.copy_(dy0)
    # CUDA graph replay (the tag advances on the device)
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        stream = side.cuda_stream
        with torch.cuda.graph(g, stream=side):
            fused()
    torch.cuda.current_stream().wait_stream(side)
    stream = torch.cuda.current_stream().cuda_stream
    for rep in range(3):
        g_fused.fill_(float("nan"))
        g.replay()
        torch.cuda.synchronize()
        err = float((g_fused - g_ref).abs().max() / g_ref.abs().max())
        ok = ok and err < 1e-5
        if rank == 0:
            print(f"   graph replay {rep}: rel err {err:.2e}", flush=True)
    dist.barrier()
t = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("PEER EXCHANGE", "OK" if t.item() == 1.0 else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if t.item() == 1.0 else 1)
