"""Out-of-bounds READ (and write) check without compute-sanitizer: every tensor a C-ABI call touches is placed so that it
ENDS exactly at the end of a mapped range of device memory, followed by reserved-but-unmapped address space (CUDA virtual
memory management: cuMemAddressReserve / cuMemCreate / cuMemMap).  A load or a TMA bulk copy that runs past the end of a
tensor - a full 16-byte vector over a ragged tail, a piece rounded up - faults with "illegal memory access" instead of
going unnoticed.  Cases: tests/test_gpu_guard_bands.py's list (one per kernel family).  The first fault ends the process
(the context is gone); the case being run is printed before its launches.

    python tools/oob_read_probe.py            # exit 0: no access past the end of any tensor
    python tools/oob_read_probe.py --control  # slabs declared 8 elements longer than the buffers: must fault
"""
import ctypes
import os
import sys

import numpy as np
import torch
from cuda.bindings import driver as cu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mi_seg_b200 as pkg  # noqa: E402
from test_gpu_guard_bands import CL_CASES, DT, NC_CASES  # noqa: E402


def ck(res):
    if res[0] != cu.CUresult.CUDA_SUCCESS:
        raise RuntimeError(f"CUDA driver error {res[0]}")
    return None if len(res) == 1 else (res[1] if len(res) == 2 else res[1:])


torch.zeros(1, device="cuda")  # primary context, current on this thread
DEV = torch.cuda.current_device()
PROP = cu.CUmemAllocationProp()
PROP.type = cu.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
PROP.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
PROP.location.id = DEV
GRAN = int(ck(cu.cuMemGetAllocationGranularity(PROP, cu.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_MINIMUM)))
_keep = []


class EndAligned:
    """`nbytes` bytes whose last byte is the last mapped byte; one granule of unmapped address space follows."""

    def __init__(self, nbytes, src=None, zero=False):
        self.n = int(nbytes)
        self.mapped = max(GRAN, (self.n + GRAN - 1) // GRAN * GRAN)
        self.va = ck(cu.cuMemAddressReserve(self.mapped + GRAN, GRAN, 0, 0))
        self.handle = ck(cu.cuMemCreate(self.mapped, PROP, 0))
        ck(cu.cuMemMap(self.va, self.mapped, 0, self.handle, 0))
        acc = cu.CUmemAccessDesc()
        acc.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
        acc.location.id = DEV
        acc.flags = cu.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
        ck(cu.cuMemSetAccess(self.va, self.mapped, [acc], 1))
        self.ptr = int(self.va) + self.mapped - self.n
        if zero and self.n:
            ck(cu.cuMemsetD8(self.ptr, 0, self.n))
        if src is not None:
            assert src.is_contiguous() and src.numel() * src.element_size() == self.n
            ck(cu.cuMemcpyDtoD(self.ptr, src.data_ptr(), self.n))
        _keep.append(self)

    def read(self, dtype):
        out = torch.empty(self.n, dtype=torch.uint8, device="cuda")
        ck(cu.cuMemcpyDtoD(out.data_ptr(), self.ptr, self.n))
        return out.view(dtype)


def ptr_array(bufs):
    return (ctypes.c_void_p * len(bufs))(*[b.ptr for b in bufs])


def sync(tag):
    try:
        torch.cuda.synchronize()
    except RuntimeError as e:  # illegal memory access: the context is gone
        print(f"FAULT in {tag}: {e}", flush=True)
        sys.exit(1)


def run_nc(shape, dtype, epi, path, name):
    lib = pkg._lib.lib()
    n, c = shape[0], shape[1]
    m = int(np.prod(shape[2:]))
    S, es = 3, torch.empty((), dtype=dtype).element_size()
    gen = torch.Generator(device="cuda").manual_seed(5)
    x_t = (torch.randn(n, c, m, device="cuda", generator=gen) * 2 + 1).to(dtype)
    dy_t = torch.randn(n, c, m, device="cuda", generator=gen).to(dtype)
    res_t = torch.randn(n, c, m, device="cuda", generator=gen).to(dtype)
    gam_t = 1 + 0.3 * torch.randn(S, c, device="cuda", generator=gen)
    bet_t = 0.3 * torch.randn(S, c, device="cuda", generator=gen)
    st_t = torch.tensor([(2 * i + 1) % S for i in range(n)], device="cuda")
    torch.cuda.synchronize()
    x, dy, res = EndAligned(n * c * m * es, x_t), EndAligned(n * c * m * es, dy_t), EndAligned(n * c * m * es, res_t)
    gam = [EndAligned(c * 4, gam_t[k].contiguous()) for k in range(S)]
    bet = [EndAligned(c * 4, bet_t[k].contiguous()) for k in range(S)]
    st = EndAligned(n * 8, st_t)
    y, dx, dres = (EndAligned(n * c * m * es) for _ in range(3))
    mean, rstd = EndAligned(n * c * 4), EndAligned(n * c * 4)
    dg, db = EndAligned(S * c * 4), EndAligned(S * c * 4)
    wsb = lib.micn_workspace_bytes(n, c, m, DT[dtype], S)
    ws = EndAligned(wsb, zero=True)
    gp, bp = ptr_array(gam), ptr_array(bet)
    stream = torch.cuda.current_stream().cuda_stream
    pkg._lib.set_option("force_path", path)
    print(f"nc {name} {shape} {dtype} epi {epi} ...", end=" ", flush=True)
    for _ in range(2):
        rc = lib.micn_fwd(x.ptr, y.ptr, res.ptr if epi == 2 else None, gp, bp, S, st.ptr, mean.ptr, rstd.ptr, n, c, m, c * m, m,
                          DT[dtype], epi, 0.01, 1e-5, ws.ptr, wsb, stream)
        assert rc == 0, lib.micn_error_string(rc)
        rc = lib.micn_bwd(dy.ptr, x.ptr, y.ptr if epi == 2 else None, gp, bp, S, st.ptr, mean.ptr, rstd.ptr, dx.ptr,
                          dres.ptr if epi == 2 else None, dg.ptr, db.ptr, n, c, m, c * m, m, DT[dtype], epi, 0.01, ws.ptr, wsb,
                          stream)
        assert rc == 0, lib.micn_error_string(rc)
    sync(name)
    took = pkg._lib.get_option("last_path")
    pkg._lib.set_option("force_path", -1)
    ok = bool(torch.isfinite(dx.read(dtype).float()).all()) and bool(torch.isfinite(y.read(dtype).float()).all())
    print(f"path {took} finite {ok}", flush=True)
    assert ok


def run_cl(shape, dtype, name):
    lib = pkg._lib.lib()
    n, c, m = shape
    S, es = 3, torch.empty((), dtype=dtype).element_size()
    gen = torch.Generator(device="cuda").manual_seed(6)
    x_t = (torch.randn(n, m, c, device="cuda", generator=gen) * 2 + 1).to(dtype)
    dy_t = torch.randn(n, m, c, device="cuda", generator=gen).to(dtype)
    gam_t = 1 + 0.3 * torch.randn(S, c, device="cuda", generator=gen)
    bet_t = 0.3 * torch.randn(S, c, device="cuda", generator=gen)
    st_t = torch.tensor([(2 * i + 1) % S for i in range(n)], device="cuda")
    torch.cuda.synchronize()
    x, dy = EndAligned(n * c * m * es, x_t), EndAligned(n * c * m * es, dy_t)
    gam = [EndAligned(c * 4, gam_t[k].contiguous()) for k in range(S)]
    bet = [EndAligned(c * 4, bet_t[k].contiguous()) for k in range(S)]
    st = EndAligned(n * 8, st_t)
    y, dx = EndAligned(n * c * m * es), EndAligned(n * c * m * es)
    mean, rstd = EndAligned(n * c * 4), EndAligned(n * c * 4)
    dg, db = EndAligned(S * c * 4), EndAligned(S * c * 4)
    wsb = lib.micn_cl_workspace_bytes(n, c, m)
    ws = EndAligned(wsb, zero=True)
    gp, bp = ptr_array(gam), ptr_array(bet)
    stream = torch.cuda.current_stream().cuda_stream
    print(f"cl {name} {shape} {dtype} ...", end=" ", flush=True)
    for _ in range(2):
        rc = lib.micn_fwd_cl(x.ptr, y.ptr, gp, bp, S, st.ptr, mean.ptr, rstd.ptr, n, c, m, DT[dtype], 1e-5, ws.ptr, wsb, stream)
        assert rc == 0, lib.micn_error_string(rc)
        rc = lib.micn_bwd_cl(dy.ptr, x.ptr, gp, bp, S, st.ptr, mean.ptr, rstd.ptr, dx.ptr, dg.ptr, db.ptr, n, c, m, DT[dtype],
                             ws.ptr, wsb, stream)
        assert rc == 0, lib.micn_error_string(rc)
    sync(name)
    ok = bool(torch.isfinite(dx.read(dtype).float()).all())
    print(f"finite {ok}", flush=True)
    assert ok


def run_dual(path):
    lib = pkg._lib.lib()
    n, c, m, S, dtype = 2, 6, 20 * 24 * 28, 3, torch.bfloat16
    gen = torch.Generator(device="cuda").manual_seed(7)
    src = [(torch.randn(n, c, m, device="cuda", generator=gen)).to(dtype) for _ in range(3)]
    par_t = [1 + 0.3 * torch.randn(S, c, device="cuda", generator=gen) for _ in range(4)]
    st_t = torch.tensor([1, 0], device="cuda")
    torch.cuda.synchronize()
    a, b, dy = (EndAligned(n * c * m * 2, t) for t in src)
    par = [[EndAligned(c * 4, t[k].contiguous()) for k in range(S)] for t in par_t]
    pp = [ptr_array(p) for p in par]
    st = EndAligned(n * 8, st_t)
    y, da, db_ = (EndAligned(n * c * m * 2) for _ in range(3))
    stats = [EndAligned(n * c * 4) for _ in range(4)]
    grads = [EndAligned(S * c * 4) for _ in range(4)]
    wsb = lib.micn_workspace_bytes(n, c, m, DT[dtype], S)
    ws = EndAligned(wsb, zero=True)
    stream = torch.cuda.current_stream().cuda_stream
    pkg._lib.set_option("force_path", path)
    print(f"dual path {path} ...", end=" ", flush=True)
    for _ in range(2):
        rc = lib.micn_fwd_dual(a.ptr, b.ptr, y.ptr, pp[0], pp[1], pp[2], pp[3], S, st.ptr, stats[0].ptr, stats[1].ptr,
                               stats[2].ptr, stats[3].ptr, n, c, m, DT[dtype], 0.01, 1e-5, ws.ptr, wsb, stream)
        assert rc == 0, lib.micn_error_string(rc)
        rc = lib.micn_bwd_dual(dy.ptr, a.ptr, b.ptr, pp[0], pp[1], pp[2], pp[3], S, st.ptr, stats[0].ptr, stats[1].ptr,
                               stats[2].ptr, stats[3].ptr, da.ptr, db_.ptr, grads[0].ptr, grads[1].ptr, grads[2].ptr,
                               grads[3].ptr, n, c, m, DT[dtype], 0.01, ws.ptr, wsb, stream)
        assert rc == 0, lib.micn_error_string(rc)
    sync(f"dual{path}")
    pkg._lib.set_option("force_path", -1)
    ok = bool(torch.isfinite(da.read(dtype).float()).all())
    print(f"finite {ok}", flush=True)
    assert ok


def control():
    """Negative control: the same placement, but the library is told the slabs are 8 elements longer than the buffers
    are - the small kernel must run off the end of x and FAULT (proves the unmapped range behind a tensor is live)."""
    lib = pkg._lib.lib()
    n, c, m, S, dtype = 1, 4, 1000, 1, torch.float32
    x = EndAligned(n * c * m * 4, torch.randn(n, c, m, device="cuda"))
    y = EndAligned(n * c * (m + 8) * 4)
    gam, bet = EndAligned(c * 4, torch.ones(c, device="cuda")), EndAligned(c * 4, torch.zeros(c, device="cuda"))
    st = EndAligned(8, torch.zeros(1, dtype=torch.int64, device="cuda"))
    mean, rstd = EndAligned(n * c * 4), EndAligned(n * c * 4)
    wsb = lib.micn_workspace_bytes(n, c, m + 8, 0, S)
    ws = EndAligned(wsb, zero=True)
    torch.cuda.synchronize()
    rc = lib.micn_fwd(x.ptr, y.ptr, None, ptr_array([gam]), ptr_array([bet]), S, st.ptr, mean.ptr, rstd.ptr, n, c, m + 8,
                      c * (m + 8), m + 8, 0, 0, 0.01, 1e-5, ws.ptr, wsb, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    try:
        torch.cuda.synchronize()
    except RuntimeError as e:
        print("CONTROL_FAULTED_AS_EXPECTED:", str(e).splitlines()[0], flush=True)
        os._exit(0)
    print("CONTROL DID NOT FAULT: the probe proves nothing", flush=True)
    os._exit(2)


if __name__ == "__main__":
    print(f"granularity {GRAN} bytes", flush=True)
    if "--control" in sys.argv:
        control()
    for shape, dtype, epi, path, name in NC_CASES:
        run_nc(shape, dtype, epi, path, name)
    extra = [((1, 3, 96, 96, 96), torch.bfloat16, 0, 2, "flat2_north_star_slab"), ((2, 3, 48, 48, 50), torch.float32, 1, 2, "flat1_ragged_pieces"),
             ((1, 48, 48, 48, 48), torch.bfloat16, 1, 4, "resident_model_shape"),
             ((1, 48, 48, 48, 48), torch.bfloat16, 2, -1, "model_shape_residual_automatic")]
    for shape, dtype, epi, path, name in extra:
        run_nc(shape, dtype, epi, path, name)
    for shape, dtype, name in CL_CASES:
        run_cl(shape, dtype, name)
    for path in (4, 2):
        run_dual(path)
    print("OOB_PROBE_DONE: no access past the end of any tensor", flush=True)
