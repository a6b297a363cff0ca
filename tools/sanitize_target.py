"""Target for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): one small call of every kernel family
through the package's public API - small (warp / 256 / 1024 threads per slab, unaligned peel), cluster, flat (both CTA
shapes, every epilogue, the cross-sample fold), resident, dual-norm, channels-last (fused, two-kernel, wide), PReLU slope
gradient, plain instance norm, the host-buffer entry point.  Sizes are the smallest each path accepts: the tools slow a
kernel down 10-100x.  Prints the path every call took; exits non-zero on a non-finite result.

    compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_target.py
    SANITIZE_SET=short compute-sanitizer --tool racecheck --kernel-name kns=micn python tools/sanitize_target.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mi_seg_b200 as pkg  # noqa: E402

SHORT = os.environ.get("SANITIZE_SET", "full") == "short"
lib = pkg._lib
torch.manual_seed(1)
ran = []


def module(c, S=3):
    mod = pkg.FastConditionalInstanceNorm3d(S, c).cuda()
    with torch.no_grad():
        for k in range(S):
            mod.norms[k].weight.normal_(1, 0.3)
            mod.norms[k].bias.normal_(0, 0.3)
    return mod


def call(tag, shape, dtype, epi="none", path=-1, **opts):
    lib.set_option("force_path", path)
    for k, v in opts.items():
        lib.set_option(k, v)
    n, c = shape[0], shape[1]
    mod = module(c)
    x = (torch.randn(*shape, device="cuda") * 2 + 1).to(dtype).requires_grad_(True)
    res = (torch.randn(*shape, device="cuda")).to(dtype).requires_grad_(True) if epi == "add_lrelu" else None
    st = [(2 * i + 1) % 3 for i in range(n)]
    y = mod(x, st) if epi == "none" else mod.forward_fused(x, st, epi, residual=res)
    fwd_path = lib.get_option("last_path")
    y.backward(torch.randn_like(y))
    torch.cuda.synchronize()
    ok = bool(torch.isfinite(y.float()).all()) and bool(torch.isfinite(x.grad.float()).all())
    ran.append((tag, tuple(shape), str(dtype).split(".")[-1], epi, fwd_path, lib.get_option("last_path"), ok))
    print(ran[-1], flush=True)
    for k in opts:
        lib.set_option(k, -1)
    lib.set_option("force_path", -1)
    assert ok, tag


bf, hf, f32 = torch.bfloat16, torch.float16, torch.float32
call("small_warp", (3, 20, 3, 3, 3), f32, "add_lrelu")
call("small_256", (2, 8, 12, 12, 12), bf, "lrelu")
call("small_1024_unaligned", (2, 3, 17, 19, 23), hf, "add_lrelu")
call("cluster", (2, 6, 24, 24, 24), bf, "lrelu", path=1)
call("flat1_fp32", (2, 5, 32, 32, 32), f32, "add_lrelu", path=2)
call("flat2_bf16", (2, 5, 32, 32, 32), bf, "none", path=2)
call("resident", (2, 5, 32, 32, 32), bf, "lrelu", path=4)
if not SHORT:
    call("flat2_fp16_lrelu", (3, 4, 24, 24, 32), hf, "lrelu", path=2)
    call("flat_pdl_plain_launch", (2, 5, 32, 32, 32), bf, "add_lrelu", path=2, flat_pdl=1, flat_coop=0)
    call("resident_fp32_res", (3, 7, 20, 24, 28), f32, "add_lrelu", path=4)
    call("small_reg", (2, 16, 16, 16, 16), bf, "none", path=0)

# dual-norm epilogue: resident and flat
for pth in ((4, 2) if not SHORT else (4,)):
    lib.set_option("force_path", pth)
    na, nb = module(6), module(6)
    a = torch.randn(2, 6, 24, 24, 24, device="cuda").bfloat16().requires_grad_(True)
    b = torch.randn(2, 6, 24, 24, 24, device="cuda").bfloat16().requires_grad_(True)
    y = pkg.norms.forward_fused_dual(na, a, nb, b, [1, 0])
    y.backward(torch.randn_like(y))
    torch.cuda.synchronize()
    print(("dual", pth, lib.get_option("last_path"), bool(torch.isfinite(a.grad.float()).all())), flush=True)
    lib.set_option("force_path", -1)

# channels-last: fused short columns, two-kernel route, wide loads
for tag, shape, dtype in (("cl_fused", (2, 128, 50), hf), ("cl_two_kernel", (2, 10, 4000), bf), ("cl_wide", (2, 96, 2000), bf),
                          ("cl_wide_fp32", (1, 20, 1300), f32)):
    if SHORT and tag == "cl_wide_fp32":
        continue
    n, c, m = shape
    mod = pkg.FastConditionalInstanceNorm1d(3, c).cuda()
    x = (torch.randn(n, m, c, device="cuda") * 2 + 1).to(dtype).permute(0, 2, 1).requires_grad_(True)
    y = mod(x, [(2 * i + 1) % 3 for i in range(n)])
    assert lib.get_option("last_path") == 3
    y.backward(torch.randn(n, m, c, device="cuda").to(dtype).permute(0, 2, 1))
    torch.cuda.synchronize()
    print((tag, shape, bool(torch.isfinite(x.grad.float()).all())), flush=True)

# PReLU slope gradient, plain instance norm
act = torch.nn.PReLU(init=0.25).cuda()
mod = module(6)
x = torch.randn(2, 6, 16, 16, 16, device="cuda", requires_grad=True)
mod.forward_fused(x, [1, 0], "lrelu", slope=act.weight).sum().backward()
plain = pkg.FastInstanceNorm3d(6, affine=True).cuda()
plain(torch.randn(2, 6, 16, 16, 16, device="cuda", requires_grad=True)).sum().backward()
torch.cuda.synchronize()
print(("prelu_and_plain", float(act.weight.grad)), flush=True)

# host-buffer entry point
if not SHORT:
    L = lib.lib()
    n, c, m, S = 3, 5, 693, 3
    xh = (torch.randn(n, c, m) * 2 + 1).pin_memory()
    dyh = torch.randn(n, c, m).pin_memory()
    yh, dxh = torch.empty_like(xh).pin_memory(), torch.empty_like(xh).pin_memory()
    gam, bet = (1 + 0.3 * torch.randn(S, c)).contiguous(), (0.3 * torch.randn(S, c)).contiguous()
    st = torch.tensor([1, 0, 2], dtype=torch.int64)
    dg, db = torch.zeros(S, c), torch.zeros(S, c)
    nbytes = L.micn_host_scratch_bytes(n, c, m, 0, S, 1)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    rc = L.micn_fwd_bwd_host(xh.data_ptr(), dyh.data_ptr(), yh.data_ptr(), dxh.data_ptr(), gam.data_ptr(), bet.data_ptr(), S,
                             st.data_ptr(), dg.data_ptr(), db.data_ptr(), n, c, m, 0, 1, 0.01, 1e-5, scratch.data_ptr(), nbytes)
    assert rc == 0
    print(("host_buffers", bool(torch.isfinite(dxh).all())), flush=True)
print("SANITIZE_TARGET_DONE", flush=True)
