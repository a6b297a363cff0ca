"""GPU parity tests proper (run on the B200 box: `pytest -m gpu`).  Everything goes through the C ABI
of libmicn.so (ctypes) - either via the drop-in nn.Module / autograd function or by raw pointers.

Tolerances are BASELINE.json's: 1e-5 relative for fp32, 1e-2 for bf16/fp16, measured as
max|a-b| / max|b| per tensor; parameter gradients (long fp32 sums) get 5x."""
import ctypes
import os

import numpy as np
import pytest
import torch

from conftest import golden_files, load_golden, rel_err
from oracle import micn_oracle as O

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-5, torch.bfloat16: 1e-2, torch.float16: 1e-2}
DT = {"float32": torch.float32, "bfloat16": torch.bfloat16, "float16": torch.float16}


@pytest.fixture(scope="module")
def pkg():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import mi_seg_b200
    mi_seg_b200._lib.lib()  # fail loudly if the extension is missing
    if os.environ.get("MICN_TEST_FLAT_PDL"):  # the whole suite with the flat kernels launched as programmatic dependents
        mi_seg_b200._lib.set_option("flat_pdl", 1)
        mi_seg_b200._lib.set_option("flat_coop", 0)
    return mi_seg_b200


def _module(pkg, dim, num_styles, gamma, beta, device="cuda"):
    cls = {1: pkg.FastConditionalInstanceNorm1d, 2: pkg.FastConditionalInstanceNorm2d,
           3: pkg.FastConditionalInstanceNorm3d}[dim]
    mod = cls(num_styles=num_styles, num_features=gamma.shape[1]).to(device)
    with torch.no_grad():
        for s in range(num_styles):
            mod.norms[s].weight.copy_(torch.from_numpy(gamma[s]))
            mod.norms[s].bias.copy_(torch.from_numpy(beta[s]))
    return mod


def _styles_arg(g, device):
    st, how = g["styles"], str(g["styles_as"])
    if how == "list":
        return [int(v) for v in st]
    if how == "int":
        return int(st[0])
    t = torch.tensor(st, dtype=torch.int64, device=device)
    return t.reshape(-1, 1) if how == "tensor_b1" else t


def _grads(mod):
    c = mod.norms[0].num_features
    z = np.zeros(c, np.float32)
    dg = np.stack([n.weight.grad.cpu().numpy() if n.weight.grad is not None else z for n in mod.norms])
    db = np.stack([n.bias.grad.cpu().numpy() if n.bias.grad is not None else z for n in mod.norms])
    present = [n.weight.grad is not None for n in mod.norms]
    return dg, db, present


# ------------------------------------------------------------------------------------------------ golden vectors
@pytest.mark.parametrize("path", golden_files("norm"), ids=os.path.basename)
@pytest.mark.parametrize("styles_on", ["as_recorded", "cuda_tensor"])
def test_module_matches_reference_golden(pkg, path, styles_on):
    g = load_golden(path)
    dtype = DT[str(g["dtype"])]
    tol = TOL[dtype]
    if "bigmean" in path:
        tol = 1e-4  # |mean|/std = 500 conditioning of the fp32 golden itself (see tests/test_oracle.py)
    mod = _module(pkg, int(g["dim"]), int(g["num_styles"]), g["gamma"], g["beta"])
    x = torch.from_numpy(g["x"]).to("cuda", dtype).requires_grad_(True)
    dy = torch.from_numpy(g["dy"]).to("cuda", dtype)
    styles = _styles_arg(g, "cuda")
    if styles_on == "cuda_tensor":
        if str(g["styles_as"]) == "int":
            pytest.skip("unbatched int style has no tensor form to vary")
        styles = torch.tensor(g["styles"], dtype=torch.int64, device="cuda")
    y = mod(x, styles)
    assert y.is_contiguous() and y.dtype == dtype and y.shape == x.shape
    y.backward(dy)
    assert rel_err(y.detach().float().cpu().numpy(), g["y"]) < tol
    assert rel_err(x.grad.float().cpu().numpy(), g["dx"]) < tol
    dg, db, present = _grads(mod)
    assert rel_err(dg, g["dgamma"]) < 5 * tol
    assert rel_err(db, g["dbeta"]) < 5 * tol
    # absent styles keep .grad None, as in the reference - also for CUDA style tensors (read back once per tensor)
    assert present == list(g["present"])


@pytest.mark.parametrize("path", golden_files("block"), ids=os.path.basename)
def test_fused_epilogues_match_reference_block_golden(pkg, path):
    """lrelu(norm1(conv1_out)) and lrelu(norm2(conv2_out) + residual) against tensors recorded inside
    the real UnetResBlock / UnetBasicBlock (dynunet_block.py:100-126, 187-203)."""
    g = load_golden(path)
    tol = 1e-5
    st = torch.tensor(g["styles"], dtype=torch.int64, device="cuda")
    S = int(g["num_styles"])

    def params(nm):
        w = [torch.from_numpy(g[f"{nm}_gamma"][s]).cuda().requires_grad_(True) for s in range(S)]
        b = [torch.from_numpy(g[f"{nm}_beta"][s]).cuda().requires_grad_(True) for s in range(S)]
        return w, b

    a2 = torch.from_numpy(g["conv2_out"]).cuda().requires_grad_(True)
    dout = torch.from_numpy(g["dout"]).cuda()
    w2, b2 = params("norm2")
    if str(g["kind"]) == "basic":
        out = pkg.instance_cond(a2, st, w2, b2, epilogue="lrelu")
        out.backward(dout)
    else:
        if "conv3_out" in g:
            a3 = torch.from_numpy(g["conv3_out"]).cuda().requires_grad_(True)
            w3, b3 = params("norm3")
            res = pkg.instance_cond(a3, st, w3, b3)
            assert rel_err(res.detach().cpu().numpy(), g["norm3_out"]) < tol
        else:
            res = torch.from_numpy(g["x"]).cuda().requires_grad_(True)
        out = pkg.instance_cond(a2, st, w2, b2, epilogue="add_lrelu", residual=res)
        out.backward(dout)
        if "conv3_out" in g:
            assert rel_err(a3.grad.cpu().numpy(), g["conv3_out_grad"]) < tol
            assert rel_err(np.stack([w.grad.cpu().numpy() for w in w3]), g["norm3_dgamma"]) < 5 * tol
            assert rel_err(np.stack([b.grad.cpu().numpy() for b in b3]), g["norm3_dbeta"]) < 5 * tol
    assert rel_err(out.detach().cpu().numpy(), g["out"]) < tol
    assert rel_err(a2.grad.cpu().numpy(), g["conv2_out_grad"]) < tol
    assert rel_err(np.stack([w.grad.cpu().numpy() for w in w2]), g["norm2_dgamma"]) < 5 * tol
    assert rel_err(np.stack([b.grad.cpu().numpy() for b in b2]), g["norm2_dbeta"]) < 5 * tol
    # norm1 -> lrelu forward
    w1, b1 = params("norm1")
    o1 = pkg.instance_cond(torch.from_numpy(g["conv1_out"]).cuda(), st, w1, b1, epilogue="lrelu")
    ref1, _, _, _ = O.fwd_epilogue_f64(g["conv1_out"], g["styles"], g["norm1_gamma"], g["norm1_beta"])
    assert rel_err(o1.detach().cpu().numpy(), ref1) < tol


# ------------------------------------------------------------------------------------------------ seeded cases vs oracle
KINK = 1e-5


def follow_kernel_on_the_kink(pre, y, band=KINK):
    """LeakyReLU has no derivative at 0: where the float64 pre-activation is within `band` of it, fp32 arithmetic may land
    on either side (one element in ~10^7 does), and dx / d(gamma) / d(beta) legitimately follow that choice.  The oracle's
    backward takes the side the kernel's own forward output shows for those elements - and only for those."""
    y = np.asarray(y, dtype=np.float64)
    near = np.abs(pre) < band
    if not near.any():
        return pre
    return np.where(near, np.where(y > 0, band, -band), pre)


def _case(pkg, shape, styles, num_styles, dtype, epilogue="none", seed=0, stride_pad=0, mean=1.0, std=2.0):
    gen = torch.Generator().manual_seed(seed)
    n, c = shape[0], shape[1]
    gamma = (1 + 0.3 * torch.randn(num_styles, c, generator=gen)).numpy()
    beta = (0.3 * torch.randn(num_styles, c, generator=gen)).numpy()
    x32 = torch.randn(*shape, generator=gen) * std + mean
    dy32 = torch.randn(*shape, generator=gen)
    r32 = torch.randn(*shape, generator=gen) * 0.7
    xq, dyq, rq = x32.to(dtype), dy32.to(dtype), r32.to(dtype)
    if stride_pad:  # x as a channel-padded view: slabs dense, stride_c != M
        big = torch.zeros((n, c + stride_pad) + tuple(shape[2:]), dtype=dtype, device="cuda")
        big[:, :c] = xq.cuda()
        x = big[:, :c].detach().requires_grad_(True)
        assert not x.is_contiguous()
    else:
        x = xq.cuda().requires_grad_(True)
    w = [torch.from_numpy(gamma[s]).cuda().requires_grad_(True) for s in range(num_styles)]
    b = [torch.from_numpy(beta[s]).cuda().requires_grad_(True) for s in range(num_styles)]
    st = torch.tensor(styles, dtype=torch.int64, device="cuda")
    res = rq.cuda().requires_grad_(True) if epilogue == "add_lrelu" else None
    y = pkg.instance_cond(x, st, w, b, epilogue=epilogue, residual=res)
    y.backward(dyq.cuda())
    torch.cuda.synchronize()
    # oracle on the SAME (quantised) inputs
    xn, dyn, rn = xq.float().numpy(), dyq.float().numpy(), rq.float().numpy()
    if epilogue == "none":
        yr, m_, r_ = O.fwd_f64(xn, styles, gamma, beta)
        dxr, dgr, dbr, _ = O.bwd_f64(dyn, xn, styles, gamma, m_, r_)
        drr = None
    else:
        yr, pre, m_, r_ = O.fwd_epilogue_f64(xn, styles, gamma, beta, residual=rn if epilogue == "add_lrelu" else None)
        pre = follow_kernel_on_the_kink(pre, y.detach().float().cpu().numpy())
        dxr, drr, dgr, dbr, _ = O.bwd_epilogue_f64(dyn, pre, xn, styles, gamma, m_, r_,
                                                   has_residual=epilogue == "add_lrelu")
    tol = TOL[dtype] * max(1.0, abs(mean) / std / 0.5)
    assert rel_err(y.detach().float().cpu().numpy(), yr) < tol
    assert rel_err(x.grad.float().cpu().numpy(), dxr) < tol
    if drr is not None:
        assert rel_err(res.grad.float().cpu().numpy(), drr) < tol
    ptol = (5e-5 if dtype == torch.float32 else 5e-3) * max(1.0, abs(mean) / std / 0.5)
    assert rel_err(np.stack([t.grad.cpu().numpy() for t in w]), dgr) < ptol
    assert rel_err(np.stack([t.grad.cpu().numpy() for t in b]), dbr) < ptol


SHAPES = [
    ((3, 20, 3, 3, 3), "tiny_27"),            # encoder10-like 3^3 slabs (unaligned, warp per slab)
    ((2, 12, 6, 6, 6), "6^3"),
    ((2, 8, 12, 12, 12), "12^3"),
    ((2, 6, 24, 24, 24), "24^3"),
    ((2, 5, 48, 48, 48), "48^3"),             # one slab = 221 KB bf16: multi-piece path
    ((1, 3, 96, 96, 96), "96^3"),             # north-star slab size
    ((2, 3, 17, 19, 23), "odd_7429"),         # M not a multiple of the vector width
    ((4, 7, 768), "1d_tokens"),               # ConditionalInstanceNorm1d-like [B, C, L]
]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16], ids=["fp32", "bf16", "fp16"])
@pytest.mark.parametrize("shape,name", SHAPES, ids=[s[1] for s in SHAPES])
def test_norm_vs_oracle(pkg, shape, name, dtype):
    styles = [(i * 7 + 1) % 2 for i in range(shape[0])]
    _case(pkg, shape, styles, 2, dtype, seed=sum(map(ord, name)) % 1000)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("epilogue", ["lrelu", "add_lrelu"])
@pytest.mark.parametrize("shape", [(2, 6, 6, 6, 6), (2, 4, 24, 24, 24), (2, 3, 48, 48, 48)], ids=["6^3", "24^3", "48^3"])
def test_epilogues_vs_oracle(pkg, shape, epilogue, dtype):
    _case(pkg, shape, [1, 0], 2, dtype, epilogue=epilogue, seed=11)


def test_one_workspace_serves_calls_of_different_shapes(pkg):
    """micn.h: a zero-filled workspace is reusable by calls of any shape.  The module path shares one workspace per
    stream, so interleave shapes whose shape-dependent regions overlap (small path with N > 1: per-channel arrival
    counters; flat path: exchange records) and check every call."""
    for rep in range(2):
        _case(pkg, (3, 5, 6, 6, 6), [0, 1, 0], 2, torch.float32, seed=rep)
        _case(pkg, (2, 40, 4, 4, 4), [1, 1], 2, torch.float32, seed=10 + rep)
        _case(pkg, (2, 3, 40, 40, 40), [1, 0], 2, torch.bfloat16, seed=20 + rep)
        _case(pkg, (5, 7, 300), [0, 1, 1, 0, 1], 2, torch.float16, seed=30 + rep)


def _random_cases(count, seed):
    rng = np.random.RandomState(seed)
    out = []
    for k in range(count):
        n, c = int(rng.randint(1, 6)), int(rng.randint(1, 41))
        nd = int(rng.randint(1, 4))
        target = int(np.exp(rng.uniform(np.log(2), np.log(60000))))  # slab lengths from 2 to ~60 k elements, log-uniform
        dims = []
        for d in range(nd - 1):
            f = int(rng.randint(1, max(2, int(round(target ** (1.0 / nd))) * 2)))
            dims.append(f)
            target = max(1, target // f)
        dims.append(max(2 if nd == 1 else 1, target))
        if int(np.prod(dims)) < 2:
            dims[-1] = 2
        num_styles = int(rng.randint(1, 5))
        styles = [int(rng.randint(-num_styles, num_styles)) for _ in range(n)]
        styles = [s + num_styles if s < 0 else s for s in styles]
        dtype = [torch.float32, torch.bfloat16, torch.float16][int(rng.randint(0, 3))]
        epilogue = ["none", "lrelu", "add_lrelu"][int(rng.randint(0, 3))]
        pad = int(rng.randint(0, 3)) if (epilogue == "none" and n > 1) else 0  # (a one-sample slice is contiguous)
        out.append(((n, c) + tuple(dims), styles, num_styles, dtype, epilogue, pad, k))
    return out


@pytest.mark.parametrize("shape,styles,num_styles,dtype,epilogue,pad,k", _random_cases(36, 20261018),
                         ids=lambda v: None if not isinstance(v, int) else None)
def test_seeded_random_shapes_vs_oracle(pkg, shape, styles, num_styles, dtype, epilogue, pad, k):
    """Ragged shapes nobody picked by hand: 1-3 spatial dims, slab lengths 2 ... 60 k elements (every path and every
    head / tail peeling case), 1-4 styles, all dtypes and epilogues, channel-padded views."""
    _case(pkg, shape, styles, num_styles, dtype, epilogue=epilogue, seed=100 + k, stride_pad=pad)


def test_strided_channel_view(pkg):
    _case(pkg, (2, 5, 16, 16, 16), [0, 1], 2, torch.float32, stride_pad=3, seed=5)
    _case(pkg, (2, 5, 48, 48, 48), [1, 1], 2, torch.bfloat16, stride_pad=2, seed=6)


def test_three_styles_and_absent(pkg):
    _case(pkg, (5, 4, 8, 8, 8), [2, 0, 2, 2, 0], 3, torch.float32, seed=7)


def test_large_mean_small_std(pkg):
    _case(pkg, (2, 3, 32, 32, 32), [0, 1], 2, torch.float32, mean=50.0, std=0.1, seed=8)


def test_channels_last_input_native_and_reference_layout(pkg):
    """PatchMerging / ViT feed channels-last strided views (patch_merging.py:136-141).  By default the columns are
    reduced in place and the output keeps the input's layout; with the native route off the
    result comes back as a dense NC* tensor like the reference's torch.stack.  Same values either way."""
    mod = pkg.FastConditionalInstanceNorm3d(2, 16).cuda()
    xcl = torch.randn(2, 6, 6, 6, 16, device="cuda")
    xv = xcl.permute(0, 4, 1, 2, 3)
    y_ref = mod(xv.contiguous(), [0, 1])
    y1 = mod(xv, [0, 1])
    assert pkg._lib.get_option("last_path") == 3
    assert y1.shape == xv.shape and y1.stride() == xv.stride()
    assert float((y1 - y_ref).detach().abs().max()) < 1e-5
    pkg.set_channels_last_native(False)
    try:
        y2 = mod(xv, [0, 1])
        assert y2.is_contiguous() and torch.equal(y2, y_ref)
    finally:
        pkg.set_channels_last_native(True)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16], ids=["fp32", "bf16", "fp16"])
@pytest.mark.parametrize("shape,name", [((2, 16, 6, 6, 6), "16x6^3"), ((1, 384, 8, 8, 8), "merge_384"),
                                        ((3, 70, 5, 7, 3), "C70_ragged"), ((4, 768, 216), "vit_tokens"),
                                        ((2, 10, 4000), "long_columns"),
                                        ((1, 100, 3, 3, 3), "one_sample_27_rows_ragged_tile"),
                                        ((1, 48, 1100), "one_sample_two_kernel_route"),
                                        ((2, 96, 2000), "wide_loads_two_samples"),
                                        ((3, 136, 1500), "wide_loads_ragged_tile"),
                                        ((2, 20, 1300), "wide_fp32_only_c20")], ids=lambda v: v if isinstance(v, str) else None)
def test_channels_last_route_vs_oracle(pkg, shape, name, dtype):
    n = shape[0]
    _cl_case(pkg, shape, [(i * 5 + 1) % 3 for i in range(n)], 3, dtype, seed=len(name))


def _cl_case(pkg, shape, styles, num_styles, dtype, seed=0, offset=0):
    """Logical NC* tensor in channels-last memory (stride_C = 1), optionally `offset` elements into its allocation (an
    offset that breaks the 16-byte alignment sends the call to the pair kernels instead of the wide ones)."""
    gen = torch.Generator().manual_seed(seed)
    n, c = shape[0], shape[1]
    gamma = (1 + 0.3 * torch.randn(num_styles, c, generator=gen)).numpy()
    beta = (0.3 * torch.randn(num_styles, c, generator=gen)).numpy()
    x32 = torch.randn(*shape, generator=gen) * 2 + 1
    dy32 = torch.randn(*shape, generator=gen)
    xq, dyq = x32.to(dtype), dy32.to(dtype)
    perm = [0] + list(range(2, len(shape))) + [1]
    inv = [0, len(shape) - 1] + list(range(1, len(shape) - 1))
    xcl = xq.permute(perm).contiguous()
    if offset:
        buf = torch.zeros(xcl.numel() + offset, dtype=dtype, device="cuda")
        buf[offset:] = xcl.cuda().reshape(-1)
        x = buf[offset:].view(xcl.shape).permute(inv).requires_grad_(True)
    else:
        x = xcl.cuda().permute(inv).requires_grad_(True)  # logical NC*, channels-last memory
    assert x.stride(1) == 1
    w = [torch.from_numpy(gamma[s]).cuda().requires_grad_(True) for s in range(num_styles)]
    b = [torch.from_numpy(beta[s]).cuda().requires_grad_(True) for s in range(num_styles)]
    st = torch.tensor(styles, dtype=torch.int64, device="cuda")
    y = pkg.instance_cond(x, st, w, b)
    assert pkg._lib.get_option("last_path") == 3 and y.stride() == x.stride()
    y.backward(dyq.cuda())
    torch.cuda.synchronize()
    xn, dyn = xq.float().numpy(), dyq.float().numpy()
    yr, m_, r_ = O.fwd_f64(xn, styles, gamma, beta)
    dxr, dgr, dbr, present = O.bwd_f64(dyn, xn, styles, gamma, m_, r_)
    tol = TOL[dtype]
    assert rel_err(y.detach().float().cpu().numpy(), yr) < tol
    assert rel_err(x.grad.float().cpu().numpy(), dxr) < tol
    ptol = 5e-5 if dtype == torch.float32 else 5e-3
    zero = np.zeros(c, dtype=np.float32)
    assert [t.grad is not None for t in w] == list(present) or all(t.grad is not None for t in w)
    assert rel_err(np.stack([zero if t.grad is None else t.grad.cpu().numpy() for t in w]), dgr) < ptol
    assert rel_err(np.stack([zero if t.grad is None else t.grad.cpu().numpy() for t in b]), dbr) < ptol


def _cl_random_cases(count, seed):
    rng = np.random.RandomState(seed)
    out = []
    for k in range(count):
        c = 2 * int(np.exp(rng.uniform(np.log(1), np.log(400))))          # even channel counts 2 ... 800
        rows = int(np.exp(rng.uniform(np.log(2), np.log(30000))))
        n = int(rng.randint(1, 6))
        while n * c * rows > 3_000_000:
            rows = max(2, rows // 2)
        num_styles = int(rng.randint(1, 5))
        styles = [int(rng.randint(0, num_styles)) for _ in range(n)]
        dtype = [torch.float32, torch.bfloat16, torch.float16][int(rng.randint(0, 3))]
        offset = [0, 0, 2, 4, 8][int(rng.randint(0, 5))]                 # elements: 4 ... 32 bytes into the allocation
        out.append(((n, c, rows), styles, num_styles, dtype, offset, k))
    return out


@pytest.mark.parametrize("shape,styles,num_styles,dtype,offset,k", _cl_random_cases(30, 20261021),
                         ids=lambda v: None if not isinstance(v, int) else None)
def test_channels_last_seeded_random_shapes_vs_oracle(pkg, shape, styles, num_styles, dtype, offset, k):
    """Token-major shapes nobody picked by hand: 2 ... 800 channels (ragged 64-channel tiles), 2 ... 30 k rows (the fused
    short-column kernels, the two-kernel route, its wide variant and the ragged row splits), 1-5 samples with 1-4 styles
    (the in-kernel parameter-gradient fold), tensors that start 4 ... 32 bytes into their allocation."""
    _cl_case(pkg, shape, styles, num_styles, dtype, seed=500 + k, offset=offset)


def test_channels_last_odd_channel_count_takes_the_copy_route(pkg):
    mod = pkg.FastConditionalInstanceNorm1d(2, 7).cuda()
    x = torch.randn(3, 50, 7, device="cuda").permute(0, 2, 1)
    y = mod(x, [0, 1, 1])
    assert pkg._lib.get_option("last_path") != 3 and y.is_contiguous()
    assert float((y - mod(x.contiguous(), [0, 1, 1])).abs().max()) == 0.0


# ------------------------------------------------------------------------------------------------ behaviours
def test_error_messages_match_reference(pkg):
    mod = pkg.FastConditionalInstanceNorm3d(2, 4).cuda()
    x = torch.randn(2, 4, 4, 4, 4, device="cuda")
    with pytest.raises(ValueError, match="Expected number of styles as batch size."):
        mod(x, [0])
    with pytest.raises(ValueError, match="Expected number of styles as batch size."):
        mod(x, 1)
    with pytest.raises(ValueError, match="Expected one style when input is not a batch."):
        mod(x[0], [0, 1])
    with pytest.raises(ValueError, match="expected 4D or 5D input"):
        mod(x[0, 0], [0])
    with pytest.raises(ValueError, match="to match num_features"):
        mod(torch.randn(2, 3, 4, 4, 4, device="cuda"), [0, 1])
    with pytest.raises(IndexError):
        mod(x, [0, 2])
    with pytest.raises(TypeError):
        mod(x, torch.tensor([0.0, 1.0]))
    with pytest.raises(ValueError, match="Expected more than 1 spatial element"):
        mod(torch.randn(2, 4, 1, 1, 1, device="cuda"), [0, 1])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.FastConditionalInstanceNorm3d(2, 4)(torch.randn(2, 4, 4, 4, 4), [0, 1])


def test_unbatched_and_negative_styles(pkg):
    mod = pkg.FastConditionalInstanceNorm3d(2, 4).cuda()
    with torch.no_grad():
        mod.norms[1].weight.fill_(2.0)
        mod.norms[1].bias.fill_(0.5)
    x = torch.randn(4, 5, 5, 5, device="cuda")
    for st in (1, [1], torch.tensor([1]), -1, torch.tensor(-1, device="cuda")):
        y = mod(x, st)
        ref, _, _ = O.fwd_f64(x[None].cpu().numpy(), [1], np.full((2, 4), 2.0), np.full((2, 4), 0.5))
        assert y.shape == x.shape
        assert rel_err(y.detach().cpu().numpy(), ref[0]) < 1e-5


def test_autocast_keeps_input_dtype(pkg):
    """instance_norm is in neither autocast list: bf16 in -> bf16 out, fp32 in -> fp32 out."""
    mod = pkg.FastConditionalInstanceNorm3d(2, 4).cuda()
    x = torch.randn(2, 4, 8, 8, 8, device="cuda")
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert mod(x, [0, 1]).dtype == torch.float32
        assert mod(x.bfloat16(), [0, 1]).dtype == torch.bfloat16


def test_repeated_calls_are_bitwise_deterministic(pkg):
    mod = pkg.FastConditionalInstanceNorm3d(2, 6).cuda()
    x = torch.randn(2, 6, 48, 48, 48, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    dy = torch.randn_like(x)
    outs = []
    for _ in range(3):
        x.grad = None
        mod.zero_grad(set_to_none=True)
        y = mod(x, [0, 1])
        y.backward(dy)
        outs.append((y.detach().clone(), x.grad.clone(), mod.norms[0].weight.grad.clone()))
    for o in outs[1:]:
        assert all(torch.equal(a, b) for a, b in zip(o, outs[0]))


def test_raw_c_abi_call(pkg):
    """micn_fwd / micn_bwd by raw pointers, as a non-Python host would call them (include/micn.h)."""
    lib = pkg._lib.lib()
    n, c, m, S = 2, 3, 4096, 2
    x = torch.randn(n, c, m, device="cuda") * 2 + 1
    dy = torch.randn(n, c, m, device="cuda")
    gam = torch.rand(S, c, device="cuda") + 0.5
    bet = torch.randn(S, c, device="cuda")
    st = torch.tensor([1, 0], device="cuda")
    y, dx = torch.empty_like(x), torch.empty_like(x)
    mean, rstd = torch.empty(n * c, device="cuda"), torch.empty(n * c, device="cuda")
    dg, db = torch.empty(S, c, device="cuda"), torch.empty(S, c, device="cuda")
    wsb = lib.micn_workspace_bytes(n, c, m, 0, S)
    ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda")
    gp = (ctypes.c_void_p * S)(*[gam[s].data_ptr() for s in range(S)])
    bp = (ctypes.c_void_p * S)(*[bet[s].data_ptr() for s in range(S)])
    stream = torch.cuda.current_stream().cuda_stream
    rc = lib.micn_fwd(x.data_ptr(), y.data_ptr(), None, gp, bp, S, st.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                      n, c, m, c * m, m, 0, 0, 0.01, 1e-5, ws.data_ptr(), wsb, stream)
    assert rc == 0
    rc = lib.micn_bwd(dy.data_ptr(), x.data_ptr(), None, gp, bp, S, st.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                      dx.data_ptr(), None, dg.data_ptr(), db.data_ptr(), n, c, m, c * m, m, 0, 0, 0.01,
                      ws.data_ptr(), wsb, stream)
    assert rc == 0
    torch.cuda.synchronize()
    yr, m_, r_ = O.fwd_f64(x.cpu().numpy(), [1, 0], gam.cpu().numpy(), bet.cpu().numpy())
    dxr, dgr, dbr, _ = O.bwd_f64(dy.cpu().numpy(), x.cpu().numpy(), [1, 0], gam.cpu().numpy(), m_, r_)
    assert rel_err(y.cpu().numpy(), yr) < 1e-5 and rel_err(dx.cpu().numpy(), dxr) < 1e-5
    assert rel_err(mean.cpu().numpy(), m_.reshape(-1)) < 1e-5 and rel_err(rstd.cpu().numpy(), r_.reshape(-1)) < 1e-5
    assert rel_err(dg.cpu().numpy(), dgr) < 5e-5 and rel_err(db.cpu().numpy(), dbr) < 5e-5
    # argument errors are reported, not raised
    assert lib.micn_fwd(x.data_ptr(), y.data_ptr(), None, gp, bp, 17, st.data_ptr(), None, None, n, c, m, c * m, m,
                        0, 0, 0.01, 1e-5, None, 0, stream) == -3
    assert lib.micn_fwd(x.data_ptr(), y.data_ptr(), None, gp, bp, S, st.data_ptr(), None, None, n, c, m, c * m, m,
                        9, 0, 0.01, 1e-5, None, 0, stream) == -2


def test_out_of_range_device_style_is_flagged_not_fatal(pkg):
    """Sync-free mode (set_sync_free_styles(True), also what a CUDA-graph capture uses): a CUDA style tensor cannot be
    validated without a device sync, so the kernel clamps the id and sets the workspace status bit instead of faulting.
    (Default mode reads the tensor back once and raises IndexError: tests/test_gpu_parity_r2.py.)"""
    lib = pkg._lib.lib()
    mod = pkg.FastConditionalInstanceNorm3d(2, 4).cuda()
    x = torch.randn(2, 4, 8, 8, 8, device="cuda")
    pkg.set_sync_free_styles(True)
    try:
        y = mod(x, torch.tensor([0, 5], device="cuda"))
    finally:
        pkg.set_sync_free_styles(False)
    torch.cuda.synchronize()
    assert torch.isfinite(y).all()
    from importlib import import_module
    f = import_module("mi-seg_b200.functional")
    ws = f._workspaces[(x.device.index, torch.cuda.current_stream().cuda_stream)]
    status = ctypes.c_int(0)
    assert lib.micn_read_status(ws.data_ptr(), torch.cuda.current_stream().cuda_stream, ctypes.byref(status)) == 0
    assert status.value & 1


# ------------------------------------------------------------------------------------------------ full size (properties)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "fp32"])
def test_north_star_shape_properties_and_torch_reference(pkg, dtype):
    """1x48x96^3 (BASELINE.json headline).  Size-independent properties of instance norm:
      * per slab, y has mean beta and variance gamma^2 * var/(var+eps);
      * per slab, sum(dx) = 0 and sum(dx * xhat) = 0 (dx is orthogonal to 1 and xhat);
      * backward is linear in dy.
    plus a torch fp32 reference (F.instance_norm on the same GPU) within BASELINE tolerance."""
    torch.manual_seed(0)
    n, c, s = 1, 48, 96
    mod = pkg.FastConditionalInstanceNorm3d(2, c).cuda()
    with torch.no_grad():
        for k in range(2):
            mod.norms[k].weight.copy_(1 + 0.3 * torch.randn(c))
            mod.norms[k].bias.copy_(0.3 * torch.randn(c))
    x = (torch.randn(n, c, s, s, s, device="cuda") * 2 + 1).to(dtype).requires_grad_(True)
    dy = torch.randn(n, c, s, s, s, device="cuda").to(dtype)
    y = mod(x, [1])
    y.backward(dy)
    gamma, beta = mod.norms[1].weight.detach(), mod.norms[1].bias.detach()
    yf = y.detach().float().reshape(c, -1)
    tol = 2e-3 if dtype == torch.bfloat16 else 2e-5
    assert (yf.mean(1) - beta).abs().max().item() < tol * 5
    assert (yf.var(1, unbiased=False).sqrt() - gamma.abs()).abs().max().item() < tol * 5
    xf = x.detach().float().reshape(c, -1)
    xhat = (xf - xf.mean(1, keepdim=True)) / (xf.var(1, unbiased=False, keepdim=True) + 1e-5).sqrt()
    dxf = x.grad.float().reshape(c, -1)
    scale = dxf.abs().sum(1)
    assert (dxf.sum(1).abs() / scale).max().item() < (2e-3 if dtype == torch.bfloat16 else 1e-5)
    assert ((dxf * xhat).sum(1).abs() / scale).max().item() < (2e-3 if dtype == torch.bfloat16 else 1e-5)
    # torch fp32 reference on the GPU
    x32 = x.detach().float().requires_grad_(True)
    g32, b32 = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = torch.nn.functional.instance_norm(x32, None, None, g32, b32, True, 0.1, 1e-5)
    yr.backward(dy.float())
    t = TOL[dtype]
    assert ((y.detach().float() - yr).abs().max() / yr.abs().max()).item() < t
    assert ((x.grad.float() - x32.grad).abs().max() / x32.grad.abs().max()).item() < t
    pt = 5e-5 if dtype == torch.float32 else 5e-3
    assert ((mod.norms[1].weight.grad - g32.grad).abs().max() / g32.grad.abs().max()).item() < pt
    assert ((mod.norms[1].bias.grad - b32.grad).abs().max() / b32.grad.abs().max()).item() < pt
    assert mod.norms[0].weight.grad is None
    # linearity of backward in dy
    x2 = x.detach().clone().requires_grad_(True)
    y2 = mod(x2, [1])
    y2.backward((2 * dy.float()).to(dtype))
    lin = (x2.grad.float() - 2 * x.grad.float()).abs().max() / x.grad.float().abs().max()
    assert lin.item() < (1e-2 if dtype == torch.bfloat16 else 1e-6)


# ------------------------------------------------------------------------------------------------ whole blocks (fused)
class _Conv(torch.nn.Module):  # MONAI Convolution(conv_only) keeps the Conv3d under `.conv`
    def __init__(self, cin, cout, k, stride):
        super().__init__()
        self.conv = torch.nn.Conv3d(cin, cout, k, stride, padding=(k - stride + 1) // 2, bias=False)

    def forward(self, x):
        return self.conv(x)


def _make_block(pkg, kind, cin, cout, stride, num_styles):
    class UnetResBlock(torch.nn.Module):  # same attribute names as dynunet_block.py:44-98
        def __init__(self):
            super().__init__()
            self.conv1, self.conv2 = _Conv(cin, cout, 3, stride), _Conv(cout, cout, 3, 1)
            self.lrelu = torch.nn.LeakyReLU(0.01, inplace=True)
            self.norm1 = pkg.FastConditionalInstanceNorm3d(num_styles, cout)
            self.norm2 = pkg.FastConditionalInstanceNorm3d(num_styles, cout)
            if kind == "res" and (cin != cout or stride != 1):
                self.conv3 = _Conv(cin, cout, 1, stride)
                self.norm3 = pkg.FastConditionalInstanceNorm3d(num_styles, cout)

    class UnetBasicBlock(UnetResBlock):
        pass

    return (UnetResBlock if kind == "res" else UnetBasicBlock)()


@pytest.mark.parametrize("path", golden_files("block"), ids=os.path.basename)
def test_fused_block_matches_reference_block_end_to_end(pkg, path):
    """A block with MI-Seg's attribute layout, the recorded conv weights and fast norms, fused by
    fuse_blocks(): output, input gradient and EVERY parameter gradient against the real block."""
    g = load_golden(path)
    kind, S = str(g["kind"]), int(g["num_styles"])
    cin, cout, stride = g["x"].shape[1], g["out"].shape[1], int(g["stride"])
    blk = _make_block(pkg, kind, cin, cout, stride, S).cuda()
    with torch.no_grad():
        for nm in ("conv1", "conv2", "conv3"):
            if hasattr(blk, nm):
                getattr(blk, nm).conv.weight.copy_(torch.from_numpy(g[f"{nm}_weight"]))
        for nm in ("norm1", "norm2", "norm3"):
            if hasattr(blk, nm):
                for s in range(S):
                    getattr(blk, nm).norms[s].weight.copy_(torch.from_numpy(g[f"{nm}_gamma"][s]))
                    getattr(blk, nm).norms[s].bias.copy_(torch.from_numpy(g[f"{nm}_beta"][s]))
    assert pkg.fuse_blocks(blk) == 1
    x = torch.from_numpy(g["x"]).cuda().requires_grad_(True)
    st = torch.tensor(g["styles"], dtype=torch.int64)  # CPU tensor: host-visible, absent styles -> None grads
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        out = blk(x, st)
        out.backward(torch.from_numpy(g["dout"]).cuda())
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    tol = 2e-5  # conv3d on cuDNN vs the CPU reference adds its own fp32 rounding
    assert rel_err(out.detach().cpu().numpy(), g["out"]) < tol
    assert rel_err(x.grad.cpu().numpy(), g["dx"]) < tol
    for nm in ("conv1", "conv2", "conv3"):
        if hasattr(blk, nm):
            assert rel_err(getattr(blk, nm).conv.weight.grad.cpu().numpy(), g[f"{nm}_weight_grad"]) < 5 * tol, nm
    for nm in ("norm1", "norm2", "norm3"):
        if hasattr(blk, nm):
            dg, db, present = _grads(getattr(blk, nm))
            assert rel_err(dg, g[f"{nm}_dgamma"]) < 5 * tol, nm
            assert rel_err(db, g[f"{nm}_dbeta"]) < 5 * tol, nm
            assert present == list(g[f"{nm}_present"])
    with pytest.raises(ValueError, match="Modalities must be passed"):
        blk(x)


# ------------------------------------------------------------------------------------------------
# the decoders' plain instance norm through the same kernels
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("affine", [True, False], ids=["affine", "plain"])
@pytest.mark.parametrize("shape", [(2, 6, 12, 12, 12), (1, 3, 48, 48, 48), (3, 5, 40)], ids=["12^3", "48^3", "1d"])
def test_plain_instance_norm_vs_oracle(pkg, shape, affine, dtype):
    gen = torch.Generator().manual_seed(11)
    n, c = shape[0], shape[1]
    cls = pkg.FastInstanceNorm1d if len(shape) == 3 else pkg.FastInstanceNorm3d
    mod = cls(c, affine=affine).cuda()
    gamma, beta = np.ones((1, c), np.float32), np.zeros((1, c), np.float32)
    if affine:
        gamma = (1 + 0.3 * torch.randn(1, c, generator=gen)).numpy()
        beta = (0.3 * torch.randn(1, c, generator=gen)).numpy()
        with torch.no_grad():
            mod.weight.copy_(torch.from_numpy(gamma[0]))
            mod.bias.copy_(torch.from_numpy(beta[0]))
    xq = (torch.randn(*shape, generator=gen) * 2 + 1).to(dtype)
    dyq = torch.randn(*shape, generator=gen).to(dtype)
    x = xq.cuda().requires_grad_(True)
    y = mod(x)
    y.backward(dyq.cuda())
    torch.cuda.synchronize()
    st = [0] * n
    yr, m_, r_ = O.fwd_f64(xq.float().numpy(), st, gamma, beta)
    dxr, dgr, dbr, _ = O.bwd_f64(dyq.float().numpy(), xq.float().numpy(), st, gamma, m_, r_)
    assert y.dtype == dtype and y.shape == x.shape
    assert rel_err(y.detach().float().cpu().numpy(), yr) < TOL[dtype]
    assert rel_err(x.grad.float().cpu().numpy(), dxr) < TOL[dtype]
    if affine:
        ptol = 5e-5 if dtype == torch.float32 else 5e-3
        assert rel_err(mod.weight.grad.cpu().numpy()[None], dgr) < ptol
        assert rel_err(mod.bias.grad.cpu().numpy()[None], dbr) < ptol


def test_plain_instance_norm_matches_torch_module_and_functional(pkg):
    """Same numbers as torch's own nn.InstanceNorm3d / F.instance_norm on the same GPU (fp32, 1e-5)."""
    torch.manual_seed(5)
    ref = torch.nn.InstanceNorm3d(7, affine=True).cuda()
    with torch.no_grad():
        ref.weight.normal_(1, 0.3)
        ref.bias.normal_(0, 0.3)
    fast = pkg.FastInstanceNorm3d(7, affine=True).cuda()
    fast.load_state_dict(ref.state_dict())
    assert isinstance(fast, torch.nn.InstanceNorm3d)
    x = torch.randn(2, 7, 20, 20, 20, device="cuda") * 3 - 1
    a, b = fast(x), ref(x)
    assert float(((a - b).abs().max() / b.abs().max()).detach()) < 1e-5
    f = pkg.fast_instance_norm(x)
    g = torch.nn.functional.instance_norm(x)
    assert float((f - g).abs().max() / g.abs().max()) < 1e-5
    # convert_plain re-classes an existing model in place
    model = torch.nn.Sequential(torch.nn.Conv3d(7, 7, 1), torch.nn.InstanceNorm3d(7, affine=True)).cuda()
    want = model(x)
    assert pkg.convert_plain(model) == 1 and type(model[1]) is pkg.FastInstanceNorm3d
    got = model(x)
    assert float((got - want).abs().max() / want.abs().max()) < 1e-5


def test_cuda_graph_replay_is_safe(pkg, request):
    """A captured graph replays the same kernel parameters; the flat path's record tags come from the workspace
    (device-side epoch), so records of the previous replay can never be mistaken for current ones."""
    torch.manual_seed(3)
    n, c, s = 2, 6, 48  # 221 KB bf16 slabs: the flat (cross-CTA exchange) path
    mod = pkg.FastConditionalInstanceNorm3d(num_styles=2, num_features=c).cuda()
    with torch.no_grad():
        for k in range(2):
            mod.norms[k].weight.normal_(1, 0.3)
            mod.norms[k].bias.normal_(0, 0.3)
    styles = torch.tensor([1, 0], device="cuda")
    x_static = torch.randn(n, c, s, s, s, device="cuda").bfloat16()
    pkg._lib.set_option("force_path", 2)  # (this shape would take the resident path, which keeps no cross-launch state)
    request.addfinalizer(lambda: pkg._lib.set_option("force_path", -1))
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side), torch.no_grad():
        for _ in range(2):
            mod(x_static, styles)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph), torch.no_grad():
        y_static = mod(x_static, styles)
    assert pkg._lib.get_option("last_path") == 2
    gamma = np.stack([m.weight.detach().cpu().numpy() for m in mod.norms])
    beta = np.stack([m.bias.detach().cpu().numpy() for m in mod.norms])
    for rep in range(4):
        xn = (torch.randn(n, c, s, s, s) * (1 + rep) + rep).bfloat16()
        x_static.copy_(xn.cuda())
        graph.replay()
        torch.cuda.synchronize()
        yr, _, _ = O.fwd_f64(xn.float().numpy(), [1, 0], gamma, beta)
        assert rel_err(y_static.float().cpu().numpy(), yr) < TOL[torch.bfloat16], rep


# ------------------------------------------------------------------------------------------------
# C-UNet's ADN "NDA" = prelu(norm(x)), slope read on the device
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("shape", [(2, 6, 12, 12, 12), (2, 3, 48, 48, 48)], ids=["12^3", "48^3"])
def test_prelu_epilogue_vs_oracle(pkg, shape, dtype):
    gen = torch.Generator().manual_seed(23)
    n, c = shape[0], shape[1]
    styles = [1, 0]
    gamma = (1 + 0.3 * torch.randn(2, c, generator=gen)).numpy()
    beta = (0.3 * torch.randn(2, c, generator=gen)).numpy()
    mod = _module(pkg, 3, 2, gamma, beta)
    act = torch.nn.PReLU(init=0.25).cuda()
    xq = (torch.randn(*shape, generator=gen) * 2 + 1).to(dtype)
    dyq = torch.randn(*shape, generator=gen).to(dtype)
    x = xq.cuda().requires_grad_(True)
    y = mod.forward_fused(x, styles, "lrelu", slope=act.weight)
    y.backward(dyq.cuda())
    torch.cuda.synchronize()
    xn, dyn = xq.float().numpy(), dyq.float().numpy()
    yr, pre, m_, r_ = O.fwd_prelu_f64(xn, styles, gamma, beta, 0.25)
    dxr, dgr, dbr, dar, _ = O.bwd_prelu_f64(dyn, pre, xn, styles, gamma, m_, r_, 0.25)
    tol = TOL[dtype]
    assert rel_err(y.detach().float().cpu().numpy(), yr) < tol
    assert rel_err(x.grad.float().cpu().numpy(), dxr) < tol
    dg, db, _ = _grads(mod)
    ptol = 5e-5 if dtype == torch.float32 else 5e-3
    assert rel_err(dg, dgr) < ptol and rel_err(db, dbr) < ptol
    # the slope gradient is read off the (rounded) output: the tolerance of the I/O dtype applies to a long sum
    da = float(act.weight.grad.item())
    assert abs(da - dar) <= (1e-4 if dtype == torch.float32 else 2e-2) * max(1.0, abs(dar)), (da, dar)
    # same numbers as the unfused composition through PyTorch's own PReLU
    x2 = xq.cuda().requires_grad_(True)
    act2 = torch.nn.PReLU(init=0.25).cuda().to(dtype)  # torch's own PReLU wants the weight in the I/O dtype
    y2 = act2(mod(x2, styles))
    assert float((y2.float() - y.float()).abs().max()) <= (1e-5 if dtype == torch.float32 else 4e-2)


# ------------------------------------------------------------------------------------------------ host-buffer entry point
HOST_CASES = [
    ((1, 48, 16 * 16 * 16), "bf16", 1, "one_sample_tapered_groups"),   # the bench layout in small: 10 channel groups
    ((3, 5, 7 * 9 * 11), "fp32", 0, "odd_lengths_fewer_channels_than_groups"),
    ((9, 4, 1024), "fp16", 1, "many_rows_pitched_copies"),              # N > 8: the 2-D copy branch
    ((2, 13, 4096), "fp32", 0, "forward_only"),
]


@pytest.mark.parametrize("shape,dt,epi,name", HOST_CASES, ids=[c[3] for c in HOST_CASES])
def test_host_buffer_call_vs_oracle(pkg, shape, dt, epi, name):
    """micn_fwd_bwd_host: host tensors in, host tensors out (what bench.py's `e2e` times), against the oracle."""
    lib = pkg._lib.lib()
    n, c, m = shape
    S = 3
    tdt = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}[dt]
    code = {"fp32": 0, "bf16": 1, "fp16": 2}[dt]
    gen = torch.Generator().manual_seed(7)
    x = (torch.randn(n, c, m, generator=gen) * 2 + 1).to(tdt).pin_memory()
    dy = torch.randn(n, c, m, generator=gen).to(tdt).pin_memory()
    y, dx = torch.empty_like(x).pin_memory(), torch.empty_like(x).pin_memory()
    gam = (1 + 0.3 * torch.randn(S, c, generator=gen)).contiguous()
    bet = (0.3 * torch.randn(S, c, generator=gen)).contiguous()
    styles = [(2 * i + 1) % S for i in range(n)]
    st = torch.tensor(styles, dtype=torch.int64)
    dg, db = torch.full((S, c), 9.0), torch.full((S, c), 9.0)
    bwd = name != "forward_only"
    nb = lib.micn_host_scratch_bytes(n, c, m, code, S, 1 if bwd else 0)
    assert nb > 0
    scratch = torch.empty(nb, dtype=torch.uint8, device="cuda")
    for _ in range(2):  # the second call reuses the scratch and the cached streams / events
        rc = lib.micn_fwd_bwd_host(x.data_ptr(), dy.data_ptr() if bwd else None, y.data_ptr(), dx.data_ptr() if bwd else None,
                                   gam.data_ptr(), bet.data_ptr(), S, st.data_ptr(), dg.data_ptr() if bwd else None,
                                   db.data_ptr() if bwd else None, n, c, m, code, epi, 0.01, 1e-5, scratch.data_ptr(), nb)
        assert rc == 0, pkg._lib.lib().micn_error_string(rc)
    xn, dyn = x.float().numpy(), dy.float().numpy()
    tol = TOL[tdt]
    if epi == 0:
        yr, m_, r_ = O.fwd_f64(xn, styles, gam.numpy(), bet.numpy())
        assert rel_err(y.float().numpy(), yr) < tol
        if bwd:
            dxr, dgr, dbr, _ = O.bwd_f64(dyn, xn, styles, gam.numpy(), m_, r_)
    else:
        yr, pre, m_, r_ = O.fwd_epilogue_f64(xn, styles, gam.numpy(), bet.numpy())
        assert rel_err(y.float().numpy(), yr) < tol
        dxr, _, dgr, dbr, _ = O.bwd_epilogue_f64(dyn, pre, xn, styles, gam.numpy(), m_, r_, has_residual=False)
    if bwd:
        ptol = 5e-5 if tdt == torch.float32 else 5e-3
        assert rel_err(dx.float().numpy(), dxr) < tol
        assert rel_err(dg.numpy(), dgr) < ptol and rel_err(db.numpy(), dbr) < ptol
    # a scratch that is too small is refused, not overrun
    assert lib.micn_fwd_bwd_host(x.data_ptr(), None, y.data_ptr(), None, gam.data_ptr(), bet.data_ptr(), S, st.data_ptr(),
                                 None, None, n, c, m, code, epi, 0.01, 1e-5, scratch.data_ptr(), 1024) == -4
