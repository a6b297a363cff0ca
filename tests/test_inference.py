"""Sharded sliding-window inference driver (BASELINE.json configs[4]): host logic on CPU.

The checker is a literal numpy restatement of MONAI 1.1.0's `sliding_window_inference` for mode="constant"
(monai/inferers/utils.py: pad up to the roi, `dense_patch_slices` window order, sum / count), written with plain loops;
the reference calls it at predict_whs.py:72-100, lightning_monai.py:86-93, test.py:153-159."""
import importlib
import itertools
import math
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
inference = importlib.import_module("mi-seg_b200.inference")


def _predictor(windows, modalities=None):
    """Deterministic stand-in for a conditional net: 2 output channels that depend on the voxel values, on the
    position inside the window and on the window's modality."""
    b = windows.shape[0]
    ramp = torch.linspace(0, 1, windows.shape[-1]).view(1, 1, *([1] * (windows.dim() - 3)), -1)
    m = torch.zeros(b) if modalities is None else modalities.reshape(-1).float()
    assert m.numel() == b, "one modality per window"
    mm = m.view(b, *([1] * (windows.dim() - 1)))
    return torch.cat([windows.sum(1, keepdim=True) * (1 + mm) + ramp, windows[:, :1] * 0.5 - mm], dim=1)


def _oracle(x, roi, overlap, mods):
    """MONAI's algorithm with loops (numpy, float64)."""
    x = x.double().numpy()
    B, nd = x.shape[0], x.ndim - 2
    orig = x.shape[2:]
    roi = tuple(roi)
    pads = [(0, 0), (0, 0)]
    for k in range(nd):
        diff = max(roi[k] - orig[k], 0)
        pads.append((diff // 2, diff - diff // 2))
    xp = np.pad(x, pads)
    size = xp.shape[2:]
    starts = []
    for s_, r_ in zip(size, roi):
        iv = r_ if r_ == s_ else max(int(r_ * (1 - overlap)), 1)
        num = int(math.ceil((s_ - r_) / iv)) + 1
        starts.append([min(i * iv, s_ - r_) for i in range(num)])
    out = np.zeros((B, 2) + size)
    cnt = np.zeros((B, 1) + size)
    nwin = 0
    for b in range(B):
        for st in itertools.product(*starts):
            sl = tuple(slice(a, a + r_) for a, r_ in zip(st, roi))
            w = torch.from_numpy(xp[(slice(b, b + 1), slice(None)) + sl]).float()
            p = _predictor(w, None if mods is None else torch.tensor([mods[b]])).double().numpy()[0]
            out[(b, slice(None)) + sl] += p
            cnt[(b, slice(None)) + sl] += 1
            nwin += 1
    out = out / cnt
    crop = (slice(None), slice(None)) + tuple(slice(p[0], p[0] + o) for p, o in zip(pads[2:], orig))
    return out[crop], nwin


def test_window_enumeration_matches_the_north_star_volume():
    sl = inference.window_slices((512, 512, 300), (96, 96, 96), 0.5)
    assert len(sl) == 10 * 10 * 6  # 600 windows
    assert sl[0] == (slice(0, 96),) * 3 and sl[-1] == (slice(416, 512), slice(416, 512), slice(204, 300))
    assert sl[1] == (slice(0, 96), slice(0, 96), slice(48, 144))  # last dimension fastest
    # a window never leaves the volume: starts are clamped, not padded
    assert all(s.stop <= lim for w in sl for s, lim in zip(w, (512, 512, 300)))


@pytest.mark.parametrize("shape,roi,overlap,swb", [
    ((2, 1, 20, 17, 13), (8, 8, 8), 0.5, 1), ((2, 1, 20, 17, 13), (8, 8, 8), 0.25, 3),
    ((1, 2, 5, 30, 9), (8, 8, 8), 0.5, 4),        # first dimension smaller than the roi: padded, then cropped
    ((3, 1, 16, 16), (8, 8), 0.5, 5),             # 2-D, batch of three images with different modalities
])
def test_unsharded_result_matches_the_monai_restatement(shape, roi, overlap, swb):
    torch.manual_seed(0)
    x = torch.randn(*shape)
    mods = [(i * 2 + 1) % 3 for i in range(shape[0])]
    want, nwin = _oracle(x, roi, overlap, mods)
    calls = []

    def pred(w, modalities=None):
        calls.append(w.shape[0])
        return _predictor(w, modalities)

    got = inference.sliding_window_inference(x, roi, swb, pred, overlap=overlap, modalities=torch.tensor(mods))
    assert got.shape == want.shape and got.dtype == x.dtype
    assert np.allclose(got.double().numpy(), want, rtol=1e-5, atol=1e-5)
    assert sum(calls) == nwin and max(calls) <= swb
    # no modalities: the predictor is called without the keyword, as MONAI would
    got0 = inference.sliding_window_inference(x, roi, swb, lambda w: _predictor(w), overlap=overlap)
    want0, _ = _oracle(x, roi, overlap, None)
    assert np.allclose(got0.double().numpy(), want0, rtol=1e-5, atol=1e-5)


def test_modalities_are_validated_like_the_norm():
    x = torch.randn(2, 1, 16, 16, 16)
    with pytest.raises(ValueError, match="Expected number of styles as batch size."):
        inference.sliding_window_inference(x, 8, 2, _predictor, overlap=0.5, modalities=torch.tensor([1]))


_WORKER = r"""
import os, sys, importlib
import torch, torch.distributed as dist
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import test_inference as T
inference = importlib.import_module("mi-seg_b200.inference")
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
torch.manual_seed(0)
x = torch.randn(2, 1, 20, 17, 13)
mods = torch.tensor([1, 0])
seen = []
def pred(w, modalities=None):
    seen.append(w.shape[0])
    return T._predictor(w, modalities)
full = inference.sliding_window_inference(x, (8, 8, 8), 3, T._predictor, overlap=0.5, modalities=mods, shard=False)
part = inference.sliding_window_inference(x, (8, 8, 8), 3, pred, overlap=0.5, modalities=mods)
total = 2 * len(inference.window_slices((20, 17, 13), (8, 8, 8), 0.5))
assert sum(seen) == len(range(rank, total, world)), (sum(seen), total)   # each rank ran only its own windows
assert torch.allclose(part, full, rtol=1e-5, atol=1e-5)                   # ... and every rank holds the full volume
dist.barrier(); dist.destroy_process_group()
print("RANK_OK", rank)
"""


def test_sharded_windows_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT})
    port = 31500 + (os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"RANK_OK {r}" in o, o


def test_window_slices_properties_hold_for_random_geometries():
    """Size-independent properties of the window enumeration (MONAI `dense_patch_slices`): every window has the roi's
    size and lies inside the volume, every voxel is covered, the count per dimension is MONAI's
    ceil((size - roi) / interval) + 1, and the order is first-dimension-slowest."""
    rng = np.random.RandomState(5)
    for _ in range(200):
        nd = int(rng.randint(1, 4))
        roi = tuple(int(rng.randint(1, 9)) for _ in range(nd))
        size = tuple(int(r + rng.randint(0, 14)) for r in roi)
        overlap = float(rng.choice([0.0, 0.25, 0.5, 0.75, 0.9]))
        sl = inference.window_slices(size, roi, overlap)
        expect = 1
        for s_, r_ in zip(size, roi):
            iv = r_ if r_ == s_ else max(int(r_ * (1 - overlap)), 1)
            expect *= int(math.ceil((s_ - r_) / iv)) + 1
        assert len(sl) == expect
        cover = np.zeros(size, dtype=np.int32)
        for w in sl:
            assert all(0 <= a.start and a.stop <= s_ and a.stop - a.start == r_ for a, s_, r_ in zip(w, size, roi))
            cover[w] += 1
        assert cover.min() >= 1
        starts = [tuple(a.start for a in w) for w in sl]
        assert starts == sorted(starts)  # lexicographic = first dimension slowest


def test_full_size_window_count_of_the_inference_config():
    """BASELINE.json configs[4] / SURVEY.md 8(d): 512 x 512 x 300 volume, roi 96^3, overlap 0.5 -> 10 * 10 * 6 windows."""
    assert len(inference.window_slices((512, 512, 300), (96, 96, 96), 0.5)) == 600
