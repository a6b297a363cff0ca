"""bring-up: channels-last kernels vs the transposing-copy route on the token-major call shapes (GPU events, C ABI)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mi_seg_b200 as pkg

lib = pkg._lib.lib()
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream
print(f"{'shape':>22s} {'dtype':>5s} {'cl fwd us':>10s} {'cl bwd us':>10s} {'copy+flat/small us':>19s}  cl GB/s (5*E*s)")
for (n, c, m) in ((1, 384, 24 ** 3), (1, 768, 12 ** 3), (1, 1536, 6 ** 3), (1, 3072, 3 ** 3), (4, 768, 216), (8, 384, 24 ** 3)):
    for tdt, code in ((torch.bfloat16, 1), (torch.float32, 0)):
        x = (torch.randn(n, m, c, device=dev) * 2 + 1).to(tdt)
        dy = torch.randn(n, m, c, device=dev).to(tdt)
        y, dx = torch.empty_like(x), torch.empty_like(x)
        stats = torch.empty(2, n * c, device=dev)
        g = torch.ones(2, c, device=dev)
        b = torch.zeros(2, c, device=dev)
        gp = (ctypes.c_void_p * 2)(g[0].data_ptr(), g[1].data_ptr())
        bp = (ctypes.c_void_p * 2)(b[0].data_ptr(), b[1].data_ptr())
        st = (torch.arange(n, device=dev) % 2).to(torch.int64)
        gr = torch.empty(2, 2, c, device=dev)
        cws = torch.zeros(lib.micn_cl_workspace_bytes(n, c, m), dtype=torch.uint8, device=dev)
        wsb = lib.micn_workspace_bytes(n, c, m, code, 2)
        ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)

        def fwd_cl():
            assert lib.micn_fwd_cl(x.data_ptr(), y.data_ptr(), gp, bp, 2, st.data_ptr(), stats[0].data_ptr(),
                                   stats[1].data_ptr(), n, c, m, code, 1e-5, cws.data_ptr(), cws.numel(), stream) == 0

        def bwd_cl():
            assert lib.micn_bwd_cl(dy.data_ptr(), x.data_ptr(), gp, bp, 2, st.data_ptr(), stats[0].data_ptr(),
                                   stats[1].data_ptr(), dx.data_ptr(), gr[0].data_ptr(), gr[1].data_ptr(), n, c, m, code,
                                   cws.data_ptr(), cws.numel(), stream) == 0

        def copy_route():
            xc = x.permute(0, 2, 1).contiguous()
            dyc = dy.permute(0, 2, 1).contiguous()
            yc, dxc = torch.empty_like(xc), torch.empty_like(xc)
            assert lib.micn_fwd(xc.data_ptr(), yc.data_ptr(), None, gp, bp, 2, st.data_ptr(), stats[0].data_ptr(),
                                stats[1].data_ptr(), n, c, m, c * m, m, code, 0, 0.01, 1e-5, ws.data_ptr(), wsb, stream) == 0
            assert lib.micn_bwd(dyc.data_ptr(), xc.data_ptr(), None, gp, bp, 2, st.data_ptr(), stats[0].data_ptr(),
                                stats[1].data_ptr(), dxc.data_ptr(), None, gr[0].data_ptr(), gr[1].data_ptr(), n, c, m,
                                c * m, m, code, 0, 0.01, ws.data_ptr(), wsb, stream) == 0
            yc.permute(0, 2, 1).contiguous()
            dxc.permute(0, 2, 1).contiguous()

        res = []
        for fn in (fwd_cl, bwd_cl, copy_route):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(50):
                fn()
            e.record()
            torch.cuda.synchronize()
            res.append(a.elapsed_time(e) / 50 * 1e3)
        es = x.element_size()
        print(f"{str((n, c, m)):>22s} {str(tdt)[6:]:>5s} {res[0]:10.1f} {res[1]:10.1f} {res[2]:19.1f}  "
              f"{5 * n * c * m * es / (res[0] + res[1]) * 1e-3:8.1f}")
