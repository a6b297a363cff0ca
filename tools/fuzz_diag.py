"""Bring-up aid: for the seeded mid-size cases of tests/test_gpu_parity_r2.py, count the elements whose error exceeds the
tolerance and say whether they sit on the activation's kink (|pre| tiny: a legitimate fp32-vs-f64 mask flip) or not."""
import sys

import numpy as np
import torch

sys.path.insert(0, "tests")
sys.path.insert(0, ".")
import mi_seg_b200 as pkg  # noqa: E402
import test_gpu_parity_r2 as T  # noqa: E402
from oracle import micn_oracle as O  # noqa: E402


def run(shape, styles, num_styles, dtype, epilogue, path, k):
    pkg._lib.set_option("force_path", path)
    gen = torch.Generator().manual_seed(300 + k)
    n, c = shape[0], shape[1]
    gamma = (1 + 0.3 * torch.randn(num_styles, c, generator=gen)).numpy()
    beta = (0.3 * torch.randn(num_styles, c, generator=gen)).numpy()
    xq = (torch.randn(*shape, generator=gen) * 2.0 + 1.0).to(dtype)
    dyq = torch.randn(*shape, generator=gen).to(dtype)
    rq = (torch.randn(*shape, generator=gen) * 0.7).to(dtype)
    x = xq.cuda().requires_grad_(True)
    w = [torch.from_numpy(gamma[s]).cuda().requires_grad_(True) for s in range(num_styles)]
    b = [torch.from_numpy(beta[s]).cuda().requires_grad_(True) for s in range(num_styles)]
    st = torch.tensor(styles, dtype=torch.int64, device="cuda")
    res = rq.cuda().requires_grad_(True) if epilogue == "add_lrelu" else None
    y = pkg.instance_cond(x, st, w, b, epilogue=epilogue, residual=res)
    y.backward(dyq.cuda())
    torch.cuda.synchronize()
    xn, dyn, rn = xq.float().numpy(), dyq.float().numpy(), rq.float().numpy()
    yr, pre, m_, r_ = O.fwd_epilogue_f64(xn, styles, gamma, beta, residual=rn if epilogue == "add_lrelu" else None)
    dxr, drr, dgr, dbr, _ = O.bwd_epilogue_f64(dyn, pre, xn, styles, gamma, m_, r_, has_residual=epilogue == "add_lrelu")
    tol = T.TOL[dtype]
    dx = x.grad.float().cpu().numpy()
    err = np.abs(dx - dxr) / np.max(np.abs(dxr))
    bad = np.argwhere(err > tol)
    print(f"shape {shape} {dtype} {epilogue} path {path} last_path {pkg._lib.get_option('last_path')}: y "
          f"{T.rel_err(y.detach().float().cpu().numpy(), yr):.2e} dx {err.max():.2e} bad {len(bad)}", flush=True)
    for idx in bad[:8]:
        i = tuple(idx)
        print("   at", i, "pre", float(pre[i]), "y", float(y.detach()[i]), "dy", float(dyn[i]), "dx", float(dx[i]), "ref", float(dxr[i]))


if __name__ == "__main__":
    cases = T._mid_size_cases(32, 20261019)
    for k in (19, 31):
        shape, styles, S, dtype, epi, path, kk = cases[k]
        for pth in (-1, path):
            run(shape, styles, S, dtype, epi, pth, kk)
