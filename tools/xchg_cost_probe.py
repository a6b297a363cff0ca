"""Single-GPU: what does the peer-exchange path add to the flat backward kernel, NVLink aside?  Both 'peer' buffers are
local; mode 0 (records only, nothing to wait for).  Headline shape."""
import ctypes
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("mi-seg_b200")
lib = pkg._lib.lib()
dev = torch.device("cuda", 0)
n, c, sp, S = 1, 48, 96, 2
m = sp ** 3
R = 4
xs = [(torch.randn(n, c, m, device=dev) * 2 + 1).bfloat16() for _ in range(R)]
dys = [torch.randn(n, c, m, device=dev).bfloat16() for _ in range(R)]
y, dx = torch.empty_like(xs[0]), torch.empty_like(xs[0])
gam, bet = torch.rand(S, c, device=dev) + 0.5, torch.randn(S, c, device=dev)
gp = (ctypes.c_void_p * S)(*[gam[k].data_ptr() for k in range(S)])
bp = (ctypes.c_void_p * S)(*[bet[k].data_ptr() for k in range(S)])
stats = torch.empty(2, n * c, device=dev)
grads = torch.empty(2, S, c, device=dev)
styles = torch.zeros(n, dtype=torch.int64, device=dev)
wsb = lib.micn_workspace_bytes(n, c, m, 1, S)
ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream().cuda_stream
lib.micn_fwd(xs[0].data_ptr(), y.data_ptr(), None, gp, bp, S, styles.data_ptr(), stats[0].data_ptr(), stats[1].data_ptr(), n, c, m,
             c * m, m, 1, 0, 0.01, 1e-5, ws.data_ptr(), wsb, stream)
for world in (2, 8):
    bufs = [torch.zeros(lib.micn_peer_buffer_bytes(c, S, world), dtype=torch.uint8, device=dev) for _ in range(world)]
    ptrs = (ctypes.c_void_p * world)(*[b.data_ptr() for b in bufs])

    def plain(i):
        assert lib.micn_bwd(dys[i].data_ptr(), xs[i].data_ptr(), None, gp, bp, S, styles.data_ptr(), stats[0].data_ptr(),
                            stats[1].data_ptr(), dx.data_ptr(), None, grads[0].data_ptr(), grads[1].data_ptr(), n, c, m, c * m, m,
                            1, 0, 0.01, ws.data_ptr(), wsb, stream) == 0

    def xchg(i):
        assert lib.micn_bwd_allreduce(dys[i].data_ptr(), xs[i].data_ptr(), None, gp, bp, S, styles.data_ptr(),
                                      stats[0].data_ptr(), stats[1].data_ptr(), dx.data_ptr(), None, grads[0].data_ptr(),
                                      grads[1].data_ptr(), n, c, m, c * m, m, 1, 0, 0.01, ws.data_ptr(), wsb, ptrs, 0, world, 0,
                                      stream) == 0

    for name, fn in (("plain micn_bwd", plain), (f"micn_bwd_allreduce world={world} mode 0 (local buffers)", xchg), ("plain again", plain)):
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            stream = side.cuda_stream
            fn(0)
            with torch.cuda.graph(graph, stream=side):
                for i in range(R):
                    fn(i)
        torch.cuda.current_stream().wait_stream(side)
        stream = torch.cuda.current_stream().cuda_stream
        for _ in range(20):
            graph.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            graph.replay()
        b.record()
        torch.cuda.synchronize()
        print(f"{name:62s} {a.elapsed_time(b) * 1e3 / (50 * R):7.2f} us per launch", flush=True)
