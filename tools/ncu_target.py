"""Short command lines for `ncu` (B200_PROFILING.md: run plain first, then under ncu): a few forward + backward launches of
one case through the C ABI.   python tools/ncu_target.py <case>
  headline   1x48x96^3 bf16, micn_fwd + micn_bwd           (flat kernels)
  res48      1x48x48^3 bf16, micn_fwd + micn_bwd           (resident kernels)
  dual       1x48x96^3 bf16, micn_fwd_dual + micn_bwd_dual (flat kernels, two normalised inputs)
  fp32_128   4x96x128^3 fp32, micn_fwd + micn_bwd          (flat kernels; the sweep's fp32 backward dip)"""
import ctypes
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("mi-seg_b200")
lib = pkg._lib.lib()
dev = torch.device("cuda", 0)
case = sys.argv[1] if len(sys.argv) > 1 else "headline"
n, c, sp, tdt, code = {"headline": (1, 48, 96, torch.bfloat16, 1), "res48": (1, 48, 48, torch.bfloat16, 1),
                       "dual": (1, 48, 96, torch.bfloat16, 1), "fp32_128": (4, 96, 128, torch.float32, 0)}[case]
m, S = sp ** 3, 2
R = 3
xs = [(torch.randn(n, c, m, device=dev) * 2 + 1).to(tdt) for _ in range(R)]
bs = [(torch.randn(n, c, m, device=dev) * 0.5).to(tdt) for _ in range(R)] if case == "dual" else None
dys = [torch.randn(n, c, m, device=dev).to(tdt) for _ in range(R)]
y, dx, dx2 = (torch.empty_like(xs[0]) for _ in range(3))
par = [1 + 0.3 * torch.randn(S, c, device=dev), 0.3 * torch.randn(S, c, device=dev),
       1 + 0.3 * torch.randn(S, c, device=dev), 0.3 * torch.randn(S, c, device=dev)]
arr = [(ctypes.c_void_p * S)(*[t[k].data_ptr() for k in range(S)]) for t in par]
stats = torch.empty(4, n * c, device=dev)
grads = torch.empty(4, S, c, device=dev)
styles = (torch.arange(n, device=dev) % S).to(torch.int64)
wsb = lib.micn_workspace_bytes(n, c, m, code, S)
ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream().cuda_stream
sp_ = [stats[k].data_ptr() for k in range(4)]
gp_ = [grads[k].data_ptr() for k in range(4)]
for it in range(6):
    i = it % R
    if case == "dual":
        rc = lib.micn_fwd_dual(xs[i].data_ptr(), bs[i].data_ptr(), y.data_ptr(), arr[0], arr[1], arr[2], arr[3], S,
                               styles.data_ptr(), sp_[0], sp_[1], sp_[2], sp_[3], n, c, m, code, 0.01, 1e-5, ws.data_ptr(), wsb,
                               stream)
        rc = rc or lib.micn_bwd_dual(dys[i].data_ptr(), xs[i].data_ptr(), bs[i].data_ptr(), arr[0], arr[1], arr[2], arr[3], S,
                                     styles.data_ptr(), sp_[0], sp_[1], sp_[2], sp_[3], dx.data_ptr(), dx2.data_ptr(), gp_[0],
                                     gp_[1], gp_[2], gp_[3], n, c, m, code, 0.01, ws.data_ptr(), wsb, stream)
    else:
        rc = lib.micn_fwd(xs[i].data_ptr(), y.data_ptr(), None, arr[0], arr[1], S, styles.data_ptr(), sp_[0], sp_[1], n, c, m,
                          c * m, m, code, 0, 0.01, 1e-5, ws.data_ptr(), wsb, stream)
        rc = rc or lib.micn_bwd(dys[i].data_ptr(), xs[i].data_ptr(), None, arr[0], arr[1], S, styles.data_ptr(), sp_[0], sp_[1],
                                dx.data_ptr(), None, gp_[0], gp_[1], n, c, m, c * m, m, code, 0, 0.01, ws.data_ptr(), wsb, stream)
    assert rc == 0, rc
torch.cuda.synchronize()
print(case, "ok; path", pkg._lib.get_option("last_path"))
