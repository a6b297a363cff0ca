"""CPU-side tests: the host mirror of the reference interface, the C-ABI surface, and the data-parallel
plumbing on gloo (world_size 2).  No kernel runs here (there is no GPU in the build container)."""
import ctypes
import importlib
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import REFERENCE_ROOT, ROOT, rel_err
from oracle import micn_oracle as O

import mi_seg_b200 as pkg

norms_mod = importlib.import_module("mi-seg_b200.norms")
parallel = importlib.import_module("mi-seg_b200.parallel")


# ------------------------------------------------------------------------------------------------ C ABI surface
def _header_functions():
    src = open(os.path.join(ROOT, "include", "micn.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(micn_[a-z_0-9]+)\s*\(", src)))


def test_library_loads_and_exports_every_header_symbol():
    lib = pkg._lib.lib()
    names = _header_functions()
    assert "micn_fwd" in names and "micn_bwd" in names
    for n in names:
        assert hasattr(lib, n), f"libmicn.so does not export {n}"
        assert n in pkg._lib.SYMBOLS, f"_lib.SYMBOLS lacks {n}"
    assert sorted(pkg._lib.SYMBOLS) == names
    assert lib.micn_version() >= 100
    assert b"dtype" in lib.micn_error_string(-2)
    assert lib.micn_workspace_bytes(2, 48, 96 ** 3, 1, 2) >= 64 + 2 * 48 * 8
    assert lib.micn_set_option(b"no_such_option", 1) != 0


def test_round2_entry_points_validate_their_arguments_without_a_gpu():
    """Host-side checks of the round-2 entry points (no compute call is made): sizes, fold modes, refusals."""
    lib = pkg._lib.lib()
    assert lib.micn_peer_buffer_bytes(48, 2, 8) == 64 + 4 * 8 * 2 * 48 * 16
    assert lib.micn_peer_buffer_bytes(0, 2, 8) == 0 and lib.micn_peer_buffer_bytes(48, 2, 0) == 0
    assert (pkg._lib.FOLD_NONE, pkg._lib.FOLD_THIS, pkg._lib.FOLD_PREVIOUS) == (0, 1, 2)
    # no device here: the supported-query answers "no" instead of failing
    assert lib.micn_dual_supported(1, 48, 96 ** 3, 1, 0) in (0, 1)
    assert lib.micn_dual_supported(0, 48, 96 ** 3, 1, 0) == 0
    # argument errors are reported before anything touches a device
    assert lib.micn_allreduce_fold(None, 0, 2, 48, 2, None, None, None) == -1
    assert lib.micn_fwd_dual(None, None, None, None, None, None, None, 2, None, None, None, None, None, 1, 4, 64, 9, 0.01, 1e-5,
                             None, 0, None) == -2
    for knob in ("pdl", "flat_pdl", "res_cs", "res_copies", "res_off", "res_min_bytes", "flat_refuse", "xchg_dbg"):
        before = lib.micn_get_option(knob.encode())
        assert lib.micn_set_option(knob.encode(), 1) == 0, knob
        assert lib.micn_get_option(knob.encode()) == 1
        assert lib.micn_set_option(knob.encode(), before) == 0


def test_cpp_binding_loads_and_matches_the_library():
    """mi-seg_b200/_micn_torch.so (csrc/micn_torch.cpp: C++ autograd functions over the same C ABI) is built by
    __graft_entry__.build(); it must load here, report the library's version and refuse CPU tensors like the ctypes path."""
    import torch

    path = os.path.join(ROOT, "mi-seg_b200", "_micn_torch.so")
    assert os.path.exists(path), "run `python -c 'import __graft_entry__ as g; g.build()'`"
    f = importlib.import_module("mi-seg_b200.functional")
    ext = f._ext()
    assert ext is not None and pkg.binding_in_use() == "cpp"
    assert ext.micn_version() == pkg._lib.lib().micn_version()
    mod = pkg.FastConditionalInstanceNorm3d(2, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mod(torch.randn(2, 4, 4, 4, 4), [0, 1])
    pkg.set_binding("ctypes")
    try:
        assert pkg.binding_in_use() == "ctypes"
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            mod(torch.randn(2, 4, 4, 4, 4), [0, 1])
    finally:
        pkg.set_binding("auto")


def test_workspace_sizes_and_knobs_without_a_gpu():
    """Pure host arithmetic of the C ABI: workspace sizes grow with the problem, carry the fixed prefix (header +
    per-channel arrival counters) that lets one zero-filled workspace serve calls of any shape, and every knob the
    tools and the docs name is accepted (read back, then restored)."""
    lib = pkg._lib.lib()
    prefix = 64 + 16384 * 4
    small = lib.micn_workspace_bytes(1, 4, 27, 0, 2)
    big = lib.micn_workspace_bytes(8, 384, 96 ** 3, 1, 2)
    assert prefix <= small < big
    assert lib.micn_workspace_bytes(0, 4, 27, 0, 2) >= 0 and lib.micn_workspace_bytes(1, 4, 27, 9, 2) == 0
    assert lib.micn_cl_workspace_bytes(2, 768, 216) >= prefix + 2 * 768 * 8
    assert lib.micn_host_scratch_bytes(1, 48, 96 ** 3, 1, 2, 1) > 4 * 48 * 96 ** 3 * 2
    assert lib.micn_host_scratch_bytes(1, 48, 96 ** 3, 9, 2, 1) == 0
    for knob in ("force_path", "flat_slots", "flat_slots_b", "flat_lag", "flat_l2_mb", "flat_piece_vecs", "flat_grid",
                 "flat_min_bytes", "flat_shape_fwd", "flat_shape_bwd", "small_tps", "small_reg", "cl_wide",
                 "host_groups", "host_taper", "host_trace", "cluster_size"):
        before = lib.micn_get_option(knob.encode())
        assert lib.micn_set_option(knob.encode(), 3) == 0, knob
        assert lib.micn_get_option(knob.encode()) == 3
        assert lib.micn_set_option(knob.encode(), before) == 0
    assert lib.micn_get_option(b"launches") >= 0


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    L = importlib.import_module("mi-seg_b200._lib")
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", str(tmp_path / "libmicn.so"))
    with pytest.raises(L.MicnError, match="no CPU / PyTorch fallback"):
        L.lib()


def test_product_package_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mi-seg_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "micn_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


# ------------------------------------------------------------------------------------------------ module mirror
@pytest.mark.parametrize("dim", [1, 2, 3])
def test_state_dict_layout_and_init_match_reference_contract(dim):
    cls = {1: pkg.FastConditionalInstanceNorm1d, 2: pkg.FastConditionalInstanceNorm2d,
           3: pkg.FastConditionalInstanceNorm3d}[dim]
    m = cls(num_styles=3, num_features=5)
    sd = m.state_dict()
    assert list(sd) == [f"norms.{s}.{k}" for s in range(3) for k in ("weight", "bias")]
    for s in range(3):
        assert torch.equal(sd[f"norms.{s}.weight"], torch.ones(5)) and torch.equal(sd[f"norms.{s}.bias"], torch.zeros(5))
        assert sd[f"norms.{s}.weight"].dtype == torch.float32
    assert len(list(m.parameters())) == 6 and not list(m.buffers())
    import inspect
    assert list(inspect.signature(cls).parameters)[:6] == ["num_styles", "num_features", "eps", "momentum", "affine",
                                                           "track_running_stats"]
    with pytest.warns(UserWarning, match="Ignored affine=False"):
        cls(2, 4, affine=False)
    with pytest.raises(NotImplementedError):
        cls(2, 4, track_running_stats=True)


def test_validation_order_and_messages_on_cpu():
    m = pkg.FastConditionalInstanceNorm3d(2, 4)
    x = torch.randn(2, 4, 3, 3, 3)
    with pytest.raises(ValueError, match="expected 4D or 5D input \\(got 3D input\\)"):
        m(x[0, 0], [0])
    with pytest.raises(ValueError, match="Expected number of styles as batch size."):
        m(x, [0, 1, 0])
    with pytest.raises(ValueError, match="Expected number of styles as batch size."):
        m(x, 0)
    with pytest.raises(ValueError, match="Expected one style when input is not a batch."):
        m(x[0], [0, 1])
    with pytest.raises(IndexError):
        m(x, [0, -3])
    with pytest.raises(TypeError):
        m(x, torch.tensor([0.5, 1.0]))
    with pytest.raises(ValueError, match="to match num_features \\(4\\), but got: 3"):
        m(torch.randn(2, 3, 3, 3, 3), [0, 1])
    with pytest.raises(ValueError, match="Expected more than 1 spatial element"):
        m(torch.randn(2, 4, 1, 1, 1), [0, 1])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(x, [0, 1])
    m1 = pkg.FastConditionalInstanceNorm1d(2, 4)
    with pytest.raises(ValueError, match="expected 2D or 3D input \\(got 4D input\\)"):
        m1(torch.randn(2, 4, 3, 3), [0, 1])


def test_host_styles_normalisation():
    f = norms_mod._host_styles
    assert f([1, -1, 0, -2], 2) == [1, 1, 0, 0]
    assert f(torch.tensor([[1], [0]]), 2) == [1, 0]
    assert f(1, 2) == [1]
    assert f([torch.tensor(1), np.int64(0)], 2) == [1, 0]
    with pytest.raises(IndexError):
        f([2], 2)
    with pytest.raises(TypeError):
        f([0.0], 2)


@pytest.mark.reference
def test_install_registers_behind_the_reference_factory(have_reference):
    if not have_reference:
        pytest.skip("/root/reference not present")
    code = r"""
import sys, importlib
sys.path.insert(0, %r); sys.path.insert(0, %r); sys.path.insert(0, %r)
import _monai_stub; _monai_stub.install()
pkg = importlib.import_module("mi-seg_b200")
classes = pkg.install()
from networks.layers.factories import Norm
from networks.layers.utils import get_norm_layer
from networks.norms.utils import parse_normalization
from networks.norms.conditional_instance_norm import _ConditionalInstanceNorm
import torch
name = parse_normalization("instance_cond", True, None, 2)
for dim in (1, 2, 3):
    m = get_norm_layer(name=name, spatial_dims=dim, channels=7)
    assert type(m) is classes[dim - 1], type(m)
    assert isinstance(m, _ConditionalInstanceNorm)
    assert not isinstance(m, torch.nn.LayerNorm)
    assert list(m.state_dict()) == ["norms.0.weight", "norms.0.bias", "norms.1.weight", "norms.1.bias"]
# a real block built through the factory picks the fast class up without any edit in networks/
from networks.blocks.dynunet_block import UnetResBlock
blk = UnetResBlock(3, 2, 4, kernel_size=3, stride=1, norm_name=name)
assert type(blk.norm1) is classes[2] and type(blk.norm3) is classes[2]
ref_keys = None
pkg.uninstall()
blk2 = UnetResBlock(3, 2, 4, kernel_size=3, stride=1, norm_name=name)
assert type(blk2.norm1).__name__ == "ConditionalInstanceNorm3d"
assert list(blk.state_dict()) == list(blk2.state_dict())      # checkpoints load with strict=True
blk.load_state_dict(blk2.state_dict(), strict=True)
conv = pkg.convert_module(blk2)
assert type(conv.norm1).__name__ == "FastConditionalInstanceNorm3d" and list(conv.state_dict()) == list(blk.state_dict())
print("OK")
""" % (ROOT, os.path.join(ROOT, "tests"), REFERENCE_ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr


@pytest.mark.reference
def test_install_plain_routes_the_decoder_norm_through_the_factory(have_reference):
    """`("instance", {"affine": True})` builds the fast subclass of torch's InstanceNorm."""
    if not have_reference:
        pytest.skip("/root/reference not present")
    code = r"""
import sys, importlib
sys.path.insert(0, %r); sys.path.insert(0, %r); sys.path.insert(0, %r)
import _monai_stub; _monai_stub.install()
import torch
pkg = importlib.import_module("mi-seg_b200")
classes = pkg.install_plain()
from networks.layers.utils import get_norm_layer
from networks.blocks.dynunet_block import UnetResBlock
name = ("instance", {"affine": True})
for dim in (1, 2, 3):
    m = get_norm_layer(name=name, spatial_dims=dim, channels=5)
    assert type(m) is classes[dim - 1] and isinstance(m, getattr(torch.nn, "InstanceNorm%%dd" %% dim))
    assert list(m.state_dict()) == ["weight", "bias"]
blk = UnetResBlock(3, 2, 4, kernel_size=3, stride=1, norm_name=name)
assert type(blk.norm1) is classes[2]
assert pkg.fuse_blocks(blk) == 1            # plain fast norms fuse like the conditional ones
pkg.uninstall()
blk2 = UnetResBlock(3, 2, 4, kernel_size=3, stride=1, norm_name=name)
assert type(blk2.norm1) is torch.nn.InstanceNorm3d
assert list(blk.state_dict()) == list(blk2.state_dict())
assert pkg.convert_plain(blk2) == 3 and type(blk2.norm2) is classes[2]
print("OK")
""" % (ROOT, os.path.join(ROOT, "tests"), REFERENCE_ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr


# ------------------------------------------------------------------------------------------------ multi-process (gloo)
def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 8, 600):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


_GLOO_WORKER = r"""
import os, sys, importlib
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
from oracle import micn_oracle as O
pkg = importlib.import_module("mi-seg_b200")
parallel = importlib.import_module("mi-seg_b200.parallel")
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
rng = np.random.default_rng(0)
N, C, S = 6, 5, 3
x = rng.normal(1.0, 2.0, (N, C, 4, 4, 4)); dy = rng.normal(size=x.shape)
styles = [0, 1, 1, 0, 1, 1]                      # style 2 absent everywhere; style 0 absent on rank 1's shard? no: see below
gamma = 1 + 0.3 * rng.normal(size=(S, C)); beta = 0.3 * rng.normal(size=(S, C))
lo, hi = parallel.shard_range(N, rank, world)
mod = pkg.FastConditionalInstanceNorm3d(S, C)
_, mean, rstd = O.fwd_f64(x[lo:hi], styles[lo:hi], gamma, beta)
_, dg, db, present = O.bwd_f64(dy[lo:hi], x[lo:hi], styles[lo:hi], gamma, mean, rstd)
for s in range(S):                                # what the kernels + autograd leave on this rank
    mod.norms[s].weight.grad = torch.tensor(dg[s], dtype=torch.float32) if present[s] else None
    mod.norms[s].bias.grad = torch.tensor(db[s], dtype=torch.float32) if present[s] else None
params = parallel.style_parameters(mod)
assert len(params) == 2 * S
parallel.allreduce_style_grads(params, average=False)
_, m_all, r_all = O.fwd_f64(x, styles, gamma, beta)
_, dg_all, db_all, present_all = O.bwd_f64(dy, x, styles, gamma, m_all, r_all)
for s in range(S):
    if present_all[s]:
        assert np.allclose(mod.norms[s].weight.grad.numpy(), dg_all[s], rtol=1e-5, atol=1e-5)
        assert np.allclose(mod.norms[s].bias.grad.numpy(), db_all[s], rtol=1e-5, atol=1e-5)
    else:
        assert mod.norms[s].weight.grad is None and mod.norms[s].bias.grad is None
dist.barrier(); dist.destroy_process_group()
print("RANK_OK", rank)
"""


def test_style_gradient_allreduce_world2_gloo(tmp_path):
    """Per-style gradients summed over ranks equal the full-batch gradients (oracle), absent styles stay
    None; 2 processes, gloo, 127.0.0.1."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER % {"root": ROOT})
    port = 29500 + (os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"RANK_OK {r}" in o, o


def test_adn_fusion_selection_logic():
    """blocks._adn_fusable picks exactly the 'N [D(p=0)] A' blocks whose norm is a fast module and whose activation is
    a single-parameter PReLU or a LeakyReLU (acti_norm.py:104-110 ordering 'NDA')."""
    blocks = importlib.import_module("mi-seg_b200.blocks")

    class ADN(torch.nn.Sequential):
        pass

    def make(norm, act, drop=None):
        m = ADN()
        m.add_module("N", norm)
        if drop is not None:
            m.add_module("D", drop)
        m.add_module("A", act)
        return m

    fast = pkg.FastConditionalInstanceNorm3d(2, 4)
    plain = pkg.FastInstanceNorm3d(4, affine=True)
    assert blocks._adn_fusable(make(fast, torch.nn.PReLU()))
    assert blocks._adn_fusable(make(plain, torch.nn.PReLU(), torch.nn.Dropout(0.0)))
    assert blocks._adn_fusable(make(fast, torch.nn.LeakyReLU(0.1)))
    assert not blocks._adn_fusable(make(fast, torch.nn.PReLU(num_parameters=4)))      # per-channel slopes: not fused
    assert not blocks._adn_fusable(make(fast, torch.nn.PReLU(), torch.nn.Dropout(0.1)))
    assert not blocks._adn_fusable(make(torch.nn.InstanceNorm3d(4), torch.nn.PReLU()))
    assert not blocks._adn_fusable(make(fast, torch.nn.ReLU()))
    model = torch.nn.Sequential(make(fast, torch.nn.PReLU()), make(fast, torch.nn.ReLU()))
    assert pkg.fuse_blocks(model) == 1
    assert model[0].forward.__func__ is blocks.adn_forward
