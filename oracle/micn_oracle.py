"""CPU oracle for MI-Seg's modality-conditioned instance norm (`instance_cond`) hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under `mi-seg_b200/` may import this file; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs use it, and
only as the checker / reported CPU baseline - never as the product path.

Parity status: PINNED.  The reference ships no tests or golden vectors, so
this restatement is pinned against outputs of the reference module itself
(`/root/reference/networks/norms/conditional_instance_norm.py`) run in the build container:
`tests/golden/make_golden.py` imports the unmodified reference, runs forward + autograd
backward and commits the vectors under `tests/golden/*.npz`; `tests/test_oracle.py` checks
every function below against them.

Two restatements live here:

* ``*_f64``  - closed-form float64 numpy math (independent of torch).  This is the oracle.
* ``port_*`` - the reference's own call sequence on torch CPU ops (per-sample
  ``F.instance_norm`` + ``torch.stack``), used as the timed CPU baseline (`kind: "port"`),
  because the arithmetic of the reference lives in ATen, not in the reference repo.

Reference lines followed (relative to /root/reference):
  networks/norms/conditional_instance_norm.py:59-60   per-sample norm picked by styles[i] + stack
  networks/norms/conditional_instance_norm.py:52-57   un-batched input
  networks/norms/conditional_instance_norm.py:40-47   styles validation
  torch/nn/modules/instancenorm.py (third party)      F.instance_norm(use_input_stats=True), biased var, eps in sqrt
  networks/blocks/dynunet_block.py:100-126            UnetResBlock epilogue  lrelu(norm2(.) + residual)
  networks/blocks/dynunet_block.py:187-203            UnetBasicBlock epilogue lrelu(norm(.))
  networks/blocks/acti_norm.py:104-110                ADN "NDA": norm -> dropout(0) -> PReLU
"""
from __future__ import annotations

import numpy as np

EPS_DEFAULT = 1e-5
LRELU_SLOPE = 0.01  # dynunet_block.py:51


# ----------------------------------------------------------------------------- helpers
def normalize_styles(styles, batch: int, num_styles: int) -> np.ndarray:
    """Host-side restatement of how the reference indexes ``self.norms[styles[i]]``
    (conditional_instance_norm.py:60): python-list indexing of a ModuleList, so negative
    indices wrap and out-of-range raises IndexError."""
    arr = np.asarray(styles)
    if arr.ndim == 0:
        arr = arr.reshape(1)
    arr = arr.reshape(-1)
    if arr.shape[0] != batch:
        raise ValueError("Expected number of styles as batch size.")
    if not np.issubdtype(arr.dtype, np.integer):
        raise TypeError("styles must be integers")
    out = arr.astype(np.int64).copy()
    for i, s in enumerate(out):
        if s < -num_styles or s >= num_styles:
            raise IndexError(f"index {int(s)} is out of range")
        if s < 0:
            out[i] = s + num_styles
    return out


def _as_ncm(x: np.ndarray):
    n, c = x.shape[0], x.shape[1]
    return x.reshape(n, c, -1)


# ----------------------------------------------------------------------------- norm fwd/bwd
def fwd_f64(x, styles, gamma, beta, eps: float = EPS_DEFAULT):
    """y = (x - mean) * rstd * gamma[styles[n]] + beta[styles[n]]  per (n, c) over all spatial dims.

    x: [N, C, *spatial]; gamma/beta: [S, C]; returns (y, mean[N,C], rstd[N,C]) in float64.
    Biased variance, eps inside the sqrt (F.instance_norm -> batch_norm with batch stats)."""
    x64 = np.asarray(x, dtype=np.float64)
    g = np.asarray(gamma, dtype=np.float64)
    b = np.asarray(beta, dtype=np.float64)
    st = normalize_styles(styles, x64.shape[0], g.shape[0])
    xm = _as_ncm(x64)
    mean = xm.mean(axis=2)
    var = ((xm - mean[:, :, None]) ** 2).mean(axis=2)
    rstd = 1.0 / np.sqrt(var + eps)
    xhat = (xm - mean[:, :, None]) * rstd[:, :, None]
    y = xhat * g[st][:, :, None] + b[st][:, :, None]
    return y.reshape(x64.shape), mean, rstd


def bwd_f64(dy, x, styles, gamma, mean, rstd):
    """Closed-form backward of fwd_f64 (autograd of stack + native_batch_norm_backward + repeat):
         dbeta[s,c]  = sum_{n: styles[n]==s} sum_M dy
         dgamma[s,c] = sum_{n: styles[n]==s} sum_M dy * xhat
         dx = gamma[s,c] * rstd * (dy - mean_M(dy) - xhat * mean_M(dy * xhat))
    Returns (dx, dgamma[S,C], dbeta[S,C], present[S] bool).  Styles absent from the batch get zero
    rows here; the reference leaves their ``.grad`` as None - `present` says which."""
    dy64 = np.asarray(dy, dtype=np.float64)
    x64 = np.asarray(x, dtype=np.float64)
    g = np.asarray(gamma, dtype=np.float64)
    num_styles = g.shape[0]
    st = normalize_styles(styles, x64.shape[0], num_styles)
    xm, dym = _as_ncm(x64), _as_ncm(dy64)
    m = xm.shape[2]
    xhat = (xm - mean[:, :, None]) * rstd[:, :, None]
    s1 = dym.sum(axis=2)
    s2 = (dym * xhat).sum(axis=2)
    a = g[st] * rstd
    dx = a[:, :, None] * (dym - (s1 / m)[:, :, None] - xhat * (s2 / m)[:, :, None])
    dgamma = np.zeros_like(g)
    dbeta = np.zeros_like(g)
    present = np.zeros(num_styles, dtype=bool)
    for n, s in enumerate(st):
        dgamma[s] += s2[n]
        dbeta[s] += s1[n]
        present[s] = True
    return dx.reshape(x64.shape), dgamma, dbeta, present


# ----------------------------------------------------------------------------- fused epilogues
def lrelu(v, slope: float = LRELU_SLOPE):
    return np.where(v > 0, v, v * slope)


def lrelu_grad(v, slope: float = LRELU_SLOPE):
    # torch: grad = x > 0 ? 1 : slope (slope used at exactly 0)
    return np.where(v > 0, 1.0, slope)


def fwd_epilogue_f64(x, styles, gamma, beta, residual=None, slope: float = LRELU_SLOPE,
                     eps: float = EPS_DEFAULT):
    """out = lrelu(norm(x) [+ residual])  - dynunet_block.py:107-111 (no residual) and :113-125."""
    y, mean, rstd = fwd_f64(x, styles, gamma, beta, eps)
    pre = y if residual is None else y + np.asarray(residual, dtype=np.float64)
    return lrelu(pre, slope), pre, mean, rstd


def bwd_epilogue_f64(dout, pre, x, styles, gamma, mean, rstd, slope: float = LRELU_SLOPE,
                     has_residual: bool = False):
    """Backward of fwd_epilogue_f64: g = dout * lrelu'(pre); dresidual = g; then bwd_f64(g)."""
    g = np.asarray(dout, dtype=np.float64) * lrelu_grad(pre, slope)
    dx, dgamma, dbeta, present = bwd_f64(g, x, styles, gamma, mean, rstd)
    return dx, (g if has_residual else None), dgamma, dbeta, present


def fwd_dual_f64(a, b, styles, gamma_a, beta_a, gamma_b, beta_b, slope: float = LRELU_SLOPE, eps: float = EPS_DEFAULT):
    """out = lrelu(norm_a(a) + norm_b(b)): the downsample branch of UnetResBlock (dynunet_block.py:113-125 with the
    conv3 / norm3 residual of :82-98, :115-118).  Returns (out, pre, (mean_a, rstd_a), (mean_b, rstd_b))."""
    ya, ma, ra = fwd_f64(a, styles, gamma_a, beta_a, eps)
    yb, mb, rb = fwd_f64(b, styles, gamma_b, beta_b, eps)
    pre = ya + yb
    return lrelu(pre, slope), pre, (ma, ra), (mb, rb)


def bwd_dual_f64(dout, pre, a, b, styles, gamma_a, gamma_b, stats_a, stats_b, slope: float = LRELU_SLOPE):
    """Backward of fwd_dual_f64: g = dout * lrelu'(pre) flows into both norms.
    Returns (da, db, dgamma_a, dbeta_a, dgamma_b, dbeta_b, present)."""
    g = np.asarray(dout, dtype=np.float64) * lrelu_grad(pre, slope)
    da, dga, dba, present = bwd_f64(g, a, styles, gamma_a, stats_a[0], stats_a[1])
    db, dgb, dbb, _ = bwd_f64(g, b, styles, gamma_b, stats_b[0], stats_b[1])
    return da, db, dga, dba, dgb, dbb, present


def prelu(v, a):
    return np.where(v > 0, v, v * a)


def fwd_prelu_f64(x, styles, gamma, beta, a: float, eps: float = EPS_DEFAULT):
    """ADN 'NDA' with dropout p=0: prelu(norm(x)), single learnable slope (acti_norm.py:104-110)."""
    y, mean, rstd = fwd_f64(x, styles, gamma, beta, eps)
    return prelu(y, a), y, mean, rstd


def bwd_prelu_f64(dout, pre, x, styles, gamma, mean, rstd, a: float):
    d = np.asarray(dout, dtype=np.float64)
    g = d * np.where(pre > 0, 1.0, a)
    da = float((d * np.where(pre > 0, 0.0, pre)).sum())
    dx, dgamma, dbeta, present = bwd_f64(g, x, styles, gamma, mean, rstd)
    return dx, dgamma, dbeta, da, present


# ----------------------------------------------------------------------------- torch CPU port (timed baseline)
def port_forward(x, styles, weights, biases, eps: float = EPS_DEFAULT):
    """The reference's call sequence on torch ops: one F.instance_norm per sample chosen by
    styles[i], then torch.stack (conditional_instance_norm.py:59-60).  `weights`/`biases` are
    per-style lists of [C] tensors, i.e. norms[s].weight / norms[s].bias."""
    import torch
    import torch.nn.functional as F

    if isinstance(styles, torch.Tensor):
        idx = [int(v) for v in styles.reshape(-1).tolist()]
    elif isinstance(styles, int):
        idx = [styles]
    else:
        idx = [int(v) for v in styles]
    outs = []
    for i, s in enumerate(idx):
        xi = x[i].unsqueeze(0)
        outs.append(F.instance_norm(xi, None, None, weights[s], biases[s], True, 0.1, eps).squeeze(0))
    return torch.stack(outs)


def port_fwd_bwd(x, dy, styles, weights, biases, eps: float = EPS_DEFAULT):
    """Forward + autograd backward through the port; returns (y, dx, [dgamma_s], [dbeta_s])."""
    import torch

    xr = x.detach().clone().requires_grad_(True)
    ws = [w.detach().clone().requires_grad_(True) for w in weights]
    bs = [b.detach().clone().requires_grad_(True) for b in biases]
    y = port_forward(xr, styles, ws, bs, eps)
    y.backward(dy)
    return y.detach(), xr.grad, [w.grad for w in ws], [b.grad for b in bs]
