"""Out-of-bounds WRITE detection without a sanitizer (`pytest -m gpu`): every output of a C-ABI call - y, dx, d(residual),
the saved statistics, d(gamma)/d(beta) and the workspace - sits between two canary regions; after the call the canaries
must be untouched and the payload must equal what the same call produces into ordinary tensors.  One case per kernel
family (small: warp / 256 / 1024 threads per slab incl. the unaligned peel; cluster; flat, both CTA shapes; resident;
dual-norm; channels-last: fused, two-kernel, wide), with lengths that are NOT multiples of the 16-byte vector wherever the
path accepts them, so a vector store past the end of a slab would land in a canary."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GUARD = 512
CANARY = 0xA5
DT = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


@pytest.fixture(scope="module")
def pkg():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import mi_seg_b200
    mi_seg_b200._lib.lib()
    return mi_seg_b200


class Guarded:
    """`nbytes` payload bytes between two GUARD-byte canaries (the payload starts 512-byte aligned)."""

    def __init__(self, nbytes, zero=False):
        self.n = int(nbytes)
        self.buf = torch.full((self.n + 2 * GUARD,), CANARY, dtype=torch.uint8, device="cuda")
        if zero:
            self.buf[GUARD:GUARD + self.n] = 0

    @property
    def ptr(self):
        return self.buf.data_ptr() + GUARD

    def payload(self, dtype):
        return self.buf[GUARD:GUARD + self.n].view(dtype)

    def intact(self):
        return bool((self.buf[:GUARD] == CANARY).all()) and bool((self.buf[GUARD + self.n:] == CANARY).all())


def _ptr_array(t):
    return (ctypes.c_void_p * t.shape[0])(*[t[k].data_ptr() for k in range(t.shape[0])])


def _run_nc(pkg, shape, dtype, epi, guarded):
    """micn_fwd + micn_bwd on seeded inputs; outputs guarded or plain.  Returns the outputs as CPU tensors."""
    lib = pkg._lib.lib()
    n, c = shape[0], shape[1]
    m = int(np.prod(shape[2:]))
    S, es = 3, torch.empty((), dtype=dtype).element_size()
    gen = torch.Generator(device="cuda").manual_seed(5)
    x = (torch.randn(n, c, m, device="cuda", generator=gen) * 2 + 1).to(dtype)
    dy = torch.randn(n, c, m, device="cuda", generator=gen).to(dtype)
    res = torch.randn(n, c, m, device="cuda", generator=gen).to(dtype) if epi == 2 else None
    gam = 1 + 0.3 * torch.randn(S, c, device="cuda", generator=gen)
    bet = 0.3 * torch.randn(S, c, device="cuda", generator=gen)
    st = torch.tensor([(2 * i + 1) % S for i in range(n)], device="cuda")
    gp, bp = _ptr_array(gam), _ptr_array(bet)
    wsb = lib.micn_workspace_bytes(n, c, m, DT[dtype], S)
    sizes = {"y": n * c * m * es, "dx": n * c * m * es, "dres": n * c * m * es, "mean": n * c * 4, "rstd": n * c * 4,
             "dg": S * c * 4, "db": S * c * 4, "ws": wsb}
    if guarded:
        bufs = {k: Guarded(v, zero=(k == "ws")) for k, v in sizes.items()}
        ptr = {k: b.ptr for k, b in bufs.items()}
    else:
        bufs = {k: torch.zeros(max(v, 1), dtype=torch.uint8, device="cuda") for k, v in sizes.items()}
        ptr = {k: b.data_ptr() for k, b in bufs.items()}
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(2):  # the second round reuses the workspace (records, epoch word, counters) as the product does
        rc = lib.micn_fwd(x.data_ptr(), ptr["y"], res.data_ptr() if epi == 2 else None, gp, bp, S, st.data_ptr(), ptr["mean"],
                          ptr["rstd"], n, c, m, c * m, m, DT[dtype], epi, 0.01, 1e-5, ptr["ws"], wsb, stream)
        assert rc == 0, lib.micn_error_string(rc)
        rc = lib.micn_bwd(dy.data_ptr(), x.data_ptr(), ptr["y"] if epi == 2 else None, gp, bp, S, st.data_ptr(), ptr["mean"],
                          ptr["rstd"], ptr["dx"], ptr["dres"] if epi == 2 else None, ptr["dg"], ptr["db"], n, c, m, c * m, m,
                          DT[dtype], epi, 0.01, ptr["ws"], wsb, stream)
        assert rc == 0, lib.micn_error_string(rc)
    torch.cuda.synchronize()
    if guarded:
        for k, b in bufs.items():
            assert b.intact(), f"canary of {k} overwritten"
        return {k: b.payload(torch.uint8).cpu() for k, b in bufs.items() if k != "ws"}
    return {k: b[:sizes[k]].cpu() for k, b in bufs.items() if k != "ws"}


NC_CASES = [
    ((3, 20, 3, 3, 3), torch.float32, 2, 0, "small_warp_27"),
    ((2, 9, 5, 7, 11), torch.bfloat16, 1, 0, "small_256_odd_385"),
    ((2, 3, 17, 19, 23), torch.float16, 2, 0, "small_1024_odd_7429"),
    ((1, 5, 33, 31, 29), torch.float32, 0, -1, "unaligned_29667_automatic"),
    ((2, 6, 24, 24, 24), torch.bfloat16, 1, 1, "cluster"),
    ((2, 5, 32, 32, 32), torch.float32, 2, 2, "flat1_fp32"),
    ((3, 4, 20, 24, 28), torch.bfloat16, 0, 2, "flat2_bf16_ragged"),
    ((3, 4, 20, 24, 28), torch.float16, 2, 2, "flat2_fp16_residual"),
    ((2, 5, 32, 32, 32), torch.bfloat16, 1, 4, "resident"),
    ((3, 7, 20, 24, 28), torch.float32, 2, 4, "resident_fp32_ragged"),
    ((1, 150, 16, 16, 16), torch.float16, 0, 4, "resident_single_cta"),
]


@pytest.mark.parametrize("shape,dtype,epi,path,name", NC_CASES, ids=[c[4] for c in NC_CASES])
def test_nc_outputs_stay_inside_their_buffers(pkg, shape, dtype, epi, path, name):
    pkg._lib.set_option("force_path", path)
    try:
        plain = _run_nc(pkg, shape, dtype, epi, guarded=False)
        took = pkg._lib.get_option("last_path")
        guarded = _run_nc(pkg, shape, dtype, epi, guarded=True)
        assert pkg._lib.get_option("last_path") == took
    finally:
        pkg._lib.set_option("force_path", -1)
    if path >= 0:
        assert took == path
    for k in plain:
        if epi != 2 and k == "dres":
            continue
        assert torch.equal(plain[k], guarded[k]), k


CL_CASES = [
    ((2, 128, 50), torch.float16, "fused_short_columns"),
    ((3, 70, 105), torch.bfloat16, "fused_ragged_tile"),
    ((2, 10, 4001), torch.bfloat16, "two_kernel_odd_rows"),
    ((2, 96, 2003), torch.bfloat16, "wide_odd_rows"),
    ((1, 20, 1301), torch.float32, "wide_fp32_one_sample"),
    ((5, 262, 118), torch.float32, "five_samples_ragged_tile"),
]


@pytest.mark.parametrize("shape,dtype,name", CL_CASES, ids=[c[2] for c in CL_CASES])
def test_channels_last_outputs_stay_inside_their_buffers(pkg, shape, dtype, name):
    lib = pkg._lib.lib()
    n, c, m = shape
    S, es = 3, torch.empty((), dtype=dtype).element_size()
    gen = torch.Generator(device="cuda").manual_seed(6)
    x = (torch.randn(n, m, c, device="cuda", generator=gen) * 2 + 1).to(dtype)
    dy = torch.randn(n, m, c, device="cuda", generator=gen).to(dtype)
    gam = 1 + 0.3 * torch.randn(S, c, device="cuda", generator=gen)
    bet = 0.3 * torch.randn(S, c, device="cuda", generator=gen)
    st = torch.tensor([(2 * i + 1) % S for i in range(n)], device="cuda")
    gp, bp = _ptr_array(gam), _ptr_array(bet)
    wsb = lib.micn_cl_workspace_bytes(n, c, m)
    sizes = {"y": n * c * m * es, "dx": n * c * m * es, "mean": n * c * 4, "rstd": n * c * 4, "dg": S * c * 4, "db": S * c * 4,
             "ws": wsb}
    stream = torch.cuda.current_stream().cuda_stream
    outs = []
    for guarded in (False, True):
        bufs = {k: Guarded(v, zero=(k == "ws")) for k, v in sizes.items()}
        for _ in range(2):
            rc = lib.micn_fwd_cl(x.data_ptr(), bufs["y"].ptr, gp, bp, S, st.data_ptr(), bufs["mean"].ptr, bufs["rstd"].ptr,
                                 n, c, m, DT[dtype], 1e-5, bufs["ws"].ptr, wsb, stream)
            assert rc == 0, lib.micn_error_string(rc)
            rc = lib.micn_bwd_cl(dy.data_ptr(), x.data_ptr(), gp, bp, S, st.data_ptr(), bufs["mean"].ptr, bufs["rstd"].ptr,
                                 bufs["dx"].ptr, bufs["dg"].ptr, bufs["db"].ptr, n, c, m, DT[dtype], bufs["ws"].ptr, wsb, stream)
            assert rc == 0, lib.micn_error_string(rc)
        torch.cuda.synchronize()
        for k, b in bufs.items():
            assert b.intact(), f"canary of {k} overwritten"
        outs.append({k: b.payload(torch.uint8).cpu() for k, b in bufs.items() if k != "ws"})
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k  # and deterministic from one workspace to the next
    y = outs[0]["y"].view(dtype).view(n, m, c).float()
    assert bool(torch.isfinite(y).all()) and abs(float(y.mean())) < 1.0


@pytest.mark.parametrize("path", [4, 2], ids=["resident", "flat"])
def test_dual_norm_outputs_stay_inside_their_buffers(pkg, path):
    lib = pkg._lib.lib()
    n, c, m, S, dtype = 2, 6, 20 * 24 * 28, 3, torch.bfloat16
    gen = torch.Generator(device="cuda").manual_seed(7)
    a = (torch.randn(n, c, m, device="cuda", generator=gen) * 2 + 1).to(dtype)
    b = (torch.randn(n, c, m, device="cuda", generator=gen) - 0.5).to(dtype)
    dy = torch.randn(n, c, m, device="cuda", generator=gen).to(dtype)
    par = [1 + 0.3 * torch.randn(S, c, device="cuda", generator=gen) for _ in range(4)]
    pp = [_ptr_array(t) for t in par]
    st = torch.tensor([1, 0], device="cuda")
    wsb = lib.micn_workspace_bytes(n, c, m, DT[dtype], S)
    names = ["y", "da", "db_", "mean_a", "rstd_a", "mean_b", "rstd_b", "dga", "dba", "dgb", "dbb", "ws"]
    sizes = dict(zip(names, [n * c * m * 2] * 3 + [n * c * 4] * 4 + [S * c * 4] * 4 + [wsb]))
    bufs = {k: Guarded(v, zero=(k == "ws")) for k, v in sizes.items()}
    stream = torch.cuda.current_stream().cuda_stream
    pkg._lib.set_option("force_path", path)
    try:
        for _ in range(2):
            rc = lib.micn_fwd_dual(a.data_ptr(), b.data_ptr(), bufs["y"].ptr, pp[0], pp[1], pp[2], pp[3], S, st.data_ptr(),
                                   bufs["mean_a"].ptr, bufs["rstd_a"].ptr, bufs["mean_b"].ptr, bufs["rstd_b"].ptr, n, c, m,
                                   DT[dtype], 0.01, 1e-5, bufs["ws"].ptr, wsb, stream)
            assert rc == 0, lib.micn_error_string(rc)
            assert pkg._lib.get_option("last_path") == path
            rc = lib.micn_bwd_dual(dy.data_ptr(), a.data_ptr(), b.data_ptr(), pp[0], pp[1], pp[2], pp[3], S, st.data_ptr(),
                                   bufs["mean_a"].ptr, bufs["rstd_a"].ptr, bufs["mean_b"].ptr, bufs["rstd_b"].ptr,
                                   bufs["da"].ptr, bufs["db_"].ptr, bufs["dga"].ptr, bufs["dba"].ptr, bufs["dgb"].ptr,
                                   bufs["dbb"].ptr, n, c, m, DT[dtype], 0.01, bufs["ws"].ptr, wsb, stream)
            assert rc == 0, lib.micn_error_string(rc)
        torch.cuda.synchronize()
    finally:
        pkg._lib.set_option("force_path", -1)
    for k, buf in bufs.items():
        assert buf.intact(), f"canary of {k} overwritten"
    assert bool(torch.isfinite(bufs["da"].payload(dtype).float()).all())
