"""Model-level legs of bench.py (BASELINE.json metric, second half: "C-SwinUNETR voxels/s @1-8 GPU"; configs[0..2], [4]).

BENCH INFRASTRUCTURE, not product code: it builds the reference's OWN nets (the unmodified `networks/` package, from
/root/reference in the build container or from the git-ignored copy `baseline/_ref/networks` on the GPU box, through
`baseline/monai_stub.py` because MONAI is absent) and times whole training / inference steps two ways on the same GPU:

  "reference"  the nets as the reference builds them (its Python-loop `instance_cond`, torch's InstanceNorm3d / LeakyReLU)
  "ours"       the same nets after `install()` + `install_plain()` + `fuse_blocks()` - the drop-in this repo is

Convolutions, window attention, MLPs, loss and optimizer are stock PyTorch/cuDNN in both (north_star).  The step follows
utils/trainer.py:33-49 (autocast forward, loss, backward, optimizer step, zero_grad(set_to_none=True)) with
tune.py:103-109's DDP wrapping (find_unused_parameters=True for instance_cond); inputs follow SURVEY.md 8(d): image
randn(B,1,96,96,96), label randint(0,6), modality = arange(B) % 2 with B=1 alternating 0/1 per step.
"""
from __future__ import annotations

import os
import sys
import time

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)
import monai_stub  # noqa: E402

OUT_CHANNELS = 6
ROI = 96


def reference_available() -> bool:
    return monai_stub.reference_root() is not None


def _nets():
    root = monai_stub.reference_root()
    if root is None:
        raise RuntimeError("no importable copy of the reference's networks/ package (run baseline/make_ref.py where "
                           "/root/reference exists)")
    monai_stub.install(root)
    from networks.nets.swin_unetr import SwinUNETR
    from networks.nets.unet import UNet
    from networks.nets.unetr import UNETR
    from networks.norms.utils import parse_normalization
    return SwinUNETR, UNETR, UNet, parse_normalization


def build_model(kind: str, variant: str, pkg=None, roi: int = ROI):
    """`kind`: swin_unetr (configs[1]: f=48, heads 3) | unetr (configs[2]: defaults) | unet (configs[0]: f=16, 4 layers,
    strides 2-2-2, 2 res units, PReLU, "NDA"); every encoder / ViT norm is instance_cond with 2 styles, the decoders keep
    the default affine instance norm (parser.py:27-36).  `variant`: "reference" | "ours"."""
    SwinUNETR, UNETR, UNet, parse_normalization = _nets()
    if variant == "ours":
        pkg.install()
        pkg.install_plain()
    elif pkg is not None:
        pkg.uninstall()
    ic = parse_normalization("instance_cond", True, None, 2)
    inst = parse_normalization("instance", True)
    torch.manual_seed(1234)
    if kind == "swin_unetr":
        net = SwinUNETR(img_size=(roi,) * 3, in_channels=1, out_channels=OUT_CHANNELS, depths=(2, 2, 2, 2),
                        num_heads=(3, 6, 12, 24), feature_size=48, vit_norm_name=ic, encoder_norm_name=ic,
                        decoder_norm_name=inst)
    elif kind == "unetr":
        net = UNETR(in_channels=1, out_channels=OUT_CHANNELS, img_size=(roi,) * 3, feature_size=16, hidden_size=768,
                    mlp_dim=3072, num_heads=12, pos_embed="perceptron", vit_norm_name=ic, encoder_norm_name=ic,
                    decoder_norm_name=inst)
    elif kind == "unet":
        net = UNet(spatial_dims=3, in_channels=1, out_channels=OUT_CHANNELS, channels=[32, 64, 128, 256],
                   strides=[2, 2, 2], kernel_size=3, up_kernel_size=3, num_res_units=2, act="prelu", norm_down=ic,
                   norm_up=inst, adn_ordering="NDA")
    else:
        raise ValueError(kind)
    fused = 0
    if variant == "ours":
        fused = pkg.fuse_blocks(net)
        pkg.uninstall()  # the factory is global state: leave it as the reference set it
    return net, fused


def dice_ce_loss(logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Soft Dice + cross-entropy on integer labels [B,1,D,H,W]: the role MONAI's DiceCELoss plays in the reference
    (lightning_monai.py:8; MONAI is absent).  Identical in both variants."""
    t = target.squeeze(1).long()
    ce = F.cross_entropy(logits.float(), t)
    p = torch.softmax(logits.float(), dim=1)
    oh = F.one_hot(t, logits.shape[1]).movedim(-1, 1).to(p.dtype)
    dims = tuple(range(2, logits.dim()))
    inter = (p * oh).sum(dims)
    den = p.sum(dims) + oh.sum(dims)
    return ce + (1.0 - (2.0 * inter + 1e-5) / (den + 1e-5)).mean()


def _norm_kernel_share(step_fn, n=2):
    """Share of the GPU kernel time of `n` steps spent in normalisation / activation-epilogue kernels (ours: micn_*;
    reference: ATen batch_norm / instance-norm kernels, the torch.stack copies, LeakyReLU and residual adds are NOT
    counted - only the norm kernels proper).  None if the profiler is unavailable."""
    try:
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for i in range(n):
                step_fn(i)
            torch.cuda.synchronize()
        tot = norm = 0.0
        for ev in prof.key_averages():
            t = float(getattr(ev, "self_device_time_total", 0.0) or getattr(ev, "self_cuda_time_total", 0.0))
            tot += t
            k = ev.key.lower()
            if "micn_" in k or "batch_norm" in k or "instance_norm" in k or "instancenorm" in k:
                norm += t
        return {"norm_kernel_us_per_step": norm / n, "all_kernel_us_per_step": tot / n,
                "share": (norm / tot) if tot > 0 else None}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)[:200]}


def train_step_bench(kind: str, variant: str, pkg, dev, world: int, rank: int, steps: int, warmup: int,
                     batch: int, dtype=torch.bfloat16, profile_share: bool = True) -> dict:
    """Time `steps` training steps (CUDA events, barrier + synchronize on both sides, max over ranks)."""
    import torch.distributed as dist

    net, fused = build_model(kind, variant, pkg)
    net = net.to(dev).train()
    model = net
    if world > 1:
        from torch.nn.parallel import DistributedDataParallel as DDP
        kw = dict(find_unused_parameters=True)  # tune.py:103-109 for instance_cond norms
        for opt in os.environ.get("MICN_DDP_OPTS", "").split(","):  # experiments: bucket_view, bucket_cap=MB, no_find_unused
            if opt == "bucket_view":
                kw["gradient_as_bucket_view"] = True
            elif opt.startswith("bucket_cap="):
                kw["bucket_cap_mb"] = int(opt.split("=")[1])
            elif opt == "no_find_unused":
                kw["find_unused_parameters"] = False
                if pkg is not None:
                    pkg.set_sync_free_styles(True)  # zero gradients instead of None for absent styles: every parameter is used
        model = DDP(net, device_ids=[dev.index], output_device=dev.index, **kw)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5)
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    R = 2  # two synthetic patches per rank (the CT / MR "patch pair"), alternated
    datas = [torch.randn(batch, 1, ROI, ROI, ROI, generator=g).to(dev) for _ in range(R)]
    targets = [torch.randint(0, OUT_CHANNELS, (batch, 1, ROI, ROI, ROI), generator=g).to(dev) for _ in range(R)]
    # modality = arange(B) % 2; with B = 1 the modality alternates per step (and per rank)
    mods = [((torch.arange(batch) + r + rank) % 2).to(dev) for r in range(R)]
    losses = []

    def step(i):
        with torch.autocast("cuda", dtype=dtype):
            out = model(datas[i % R], mods[i % R])
            loss = dice_ce_loss(out, targets[i % R])
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    for i in range(max(warmup, 2)):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        losses.append(step(i))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    ms_step = ms / steps
    out = {"ms_per_step": ms_step, "voxels_per_s": batch * ROI ** 3 * world / (ms_step * 1e-3),
           "loss_first": float(losses[0]), "loss_last": float(losses[-1]), "blocks_fused": fused,
           "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9}
    if profile_share and (world == 1 or os.environ.get("MICN_PROFILE_DDP")):
        share = _norm_kernel_share(step)  # (every rank runs the profiled steps: they contain collectives)
        if rank == 0:
            out["norm_share"] = share
    if pkg is not None:
        pkg.set_sync_free_styles(False)
    del model, net, opt
    torch.cuda.empty_cache()
    return out


def graphed_step_bench(kind: str, pkg, dev, steps: int, warmup: int, batch: int, dtype=torch.bfloat16) -> dict:
    """The "ours" training step captured ONCE into a CUDA graph and replayed (single GPU): forward, loss, backward and
    AdamW(capturable) in one graph, the modality fed through a static device tensor.  Possible because the drop-in reads
    the style ids ON THE DEVICE (sync-free mode, selected automatically during capture); the reference's norm indexes its
    ModuleList with `styles[i]` - one host sync per sample per norm call - and cannot be captured.  In this mode styles
    absent from the batch receive zero gradients instead of None (set_sync_free_styles docstring)."""
    net, fused = build_model(kind, "ours", pkg)
    net = net.to(dev).train()
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4, weight_decay=1e-5, capturable=True)
    g = torch.Generator(device="cpu").manual_seed(100)
    R = 2
    datas = [torch.randn(batch, 1, ROI, ROI, ROI, generator=g).to(dev) for _ in range(R)]
    targets = [torch.randint(0, OUT_CHANNELS, (batch, 1, ROI, ROI, ROI), generator=g).to(dev) for _ in range(R)]
    mods = [((torch.arange(batch) + r) % 2).to(dev) for r in range(R)]
    s_data, s_target, s_mod = datas[0].clone(), targets[0].clone(), mods[0].clone()
    s_loss = torch.zeros((), device=dev)
    pkg.set_sync_free_styles(True)
    try:
        def body():
            with torch.autocast("cuda", dtype=dtype):
                loss = dice_ce_loss(net(s_data, s_mod), s_target)
            loss.backward()
            opt.step()
            opt.zero_grad(set_to_none=False)
            s_loss.copy_(loss.detach())

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            body()

        def step(i):
            s_data.copy_(datas[i % R], non_blocking=True)
            s_target.copy_(targets[i % R], non_blocking=True)
            s_mod.copy_(mods[i % R], non_blocking=True)
            graph.replay()

        for i in range(max(warmup, 2)):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        first = None
        e0.record()
        for i in range(steps):
            step(i)
            if first is None:
                first = s_loss.clone()
        e1.record()
        torch.cuda.synchronize()
        ms_step = e0.elapsed_time(e1) / steps
        return {"ms_per_step": ms_step, "voxels_per_s": batch * ROI ** 3 / (ms_step * 1e-3), "loss_first": float(first),
                "loss_last": float(s_loss), "blocks_fused": fused,
                "note": "whole step (forward, loss, backward, AdamW capturable) replayed from ONE CUDA graph; style ids read on "
                        "the device; absent styles get zero gradients in this mode"}
    finally:
        pkg.set_sync_free_styles(False)
        del net, opt
        torch.cuda.empty_cache()


def model_step_leg(kind: str, pkg, dev, world: int, rank: int, steps: int, warmup: int, batch: int) -> dict:
    """Both variants back to back on the same GPU(s); voxels/s = B * 96^3 * world / step time."""
    res = {"what": f"{kind} training step (reference nets from its own networks/ package; bf16 autocast, AdamW, "
                   f"Dice+CE on synthetic 96^3 patches, B={batch}/GPU, modality alternating 0/1; DDP over NCCL when "
                   f"n_gpus > 1, find_unused_parameters=True as tune.py:103-109)",
           "n_gpus": world, "batch_per_gpu": batch, "steps": steps, "warmup": max(warmup, 2), "dtype": "bf16 autocast"}
    for variant in ("reference", "ours"):
        try:
            res[variant] = train_step_bench(kind, variant, pkg, dev, world, rank, steps, warmup, batch)
        except Exception as e:  # noqa: BLE001 - a failing leg must not take the headline line down
            res[variant] = {"error": repr(e)[:400]}
    if world == 1 and not os.environ.get("MICN_NO_GRAPHED_STEP"):
        try:
            res["ours_cuda_graph"] = graphed_step_bench(kind, pkg, dev, steps, warmup, batch)
        except Exception as e:  # noqa: BLE001
            res["ours_cuda_graph"] = {"error": repr(e)[:400]}
            try:
                torch.cuda.synchronize()
            except Exception:  # noqa: BLE001
                pass
    if "ms_per_step" in res.get("reference", {}) and "ms_per_step" in res.get("ours", {}):
        res["speedup"] = res["reference"]["ms_per_step"] / res["ours"]["ms_per_step"]
        res["voxels_per_s"] = res["ours"]["voxels_per_s"]
    return res


def unet_cpu_leg(steps: int = 1) -> dict:
    """BASELINE.json configs[0] / BASELINE.md section 4: the reference C-UNet forward + backward on one synthetic 96^3
    CT patch, then one MR patch, batch 1, fp32, on the host CPU (the reference's own CPU-runnable case; the product has
    no CPU path)."""
    net, _ = build_model("unet", "reference")
    net.train()
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 1, ROI, ROI, ROI, generator=g)
    y = torch.randint(0, OUT_CHANNELS, (1, 1, ROI, ROI, ROI), generator=g)
    times = []
    for it in range(steps + 1):  # first pass = warm-up
        t0 = time.perf_counter()
        for m in (0, 1):  # CT then MR
            net.zero_grad(set_to_none=True)
            loss = dice_ce_loss(net(x, torch.tensor([m])), y)
            loss.backward()
        if it:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return {"what": "reference C-UNet (f=16, 4 layers, instance_cond) fwd+bwd on a 96^3 CT patch then an MR patch, B=1, "
                    "fp32, host CPU", "s_per_patch_pair": sec, "voxels_per_s": 2 * ROI ** 3 / sec,
            "cores": os.cpu_count() or 1, "pairs_timed": len(times)}


def sliding_window_leg(pkg, dev, world: int, rank: int, volume=(512, 512, 300), sw_batch_size: int = 4,
                       dtype=torch.bfloat16, check: bool = True) -> dict:
    """BASELINE.json configs[4] (predict_whs.py:72-99): C-Swin-UNETR over a synthetic 512 x 512 x 300 MR volume, roi 96^3,
    overlap 0.5 -> 10 * 10 * 6 = 600 windows, dealt round-robin to the ranks by `sliding_window_inference` (one all-reduce
    of the blended maps at the end), modality 1 expanded to the window batch."""
    import torch.distributed as dist

    net, fused = build_model("swin_unetr", "ours", pkg)
    net = net.to(dev).eval()
    g = torch.Generator().manual_seed(7)
    vol = torch.randn(1, 1, *volume, generator=g).to(dev)
    mod = torch.tensor([1], device=dev)
    nwin = len(pkg.window_slices(volume, (ROI,) * 3, 0.5))

    def predictor(w, modalities=None):
        with torch.autocast("cuda", dtype=dtype):
            return net(w, modalities)

    def run():
        with torch.no_grad():
            return pkg.sliding_window_inference(vol, (ROI,) * 3, sw_batch_size, predictor, overlap=0.5, modalities=mod)

    # warm-up on a small volume (cuDNN autotune, workspaces)
    with torch.no_grad():
        pkg.sliding_window_inference(vol[..., :ROI, :ROI, :ROI], (ROI,) * 3, sw_batch_size, predictor, overlap=0.5,
                                     modalities=mod, shard=False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    # the final exchange alone: all-reduce of the [1, 6, 512, 512, 300] fp32 logit map + the count map
    ar_ms = None
    if world > 1:
        buf = torch.zeros_like(out, dtype=torch.float32)
        cnt = torch.zeros((1, 1) + tuple(out.shape[2:]), device=dev)
        dist.all_reduce(buf)
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dist.all_reduce(buf)
        dist.all_reduce(cnt)
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([ms, a.elapsed_time(b)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ar_ms = float(t[0].item()), float(t[1].item())
        del buf, cnt
    res = {"what": f"C-Swin-UNETR (ours: install + install_plain + fuse_blocks, bf16 autocast) sliding-window inference "
                   f"over a synthetic {volume[0]}x{volume[1]}x{volume[2]} MR volume, roi 96^3, overlap 0.5, "
                   f"sw_batch_size {sw_batch_size}, windows dealt round-robin to {world} rank(s)",
           "windows": nwin, "n_gpus": world, "s_per_volume": ms * 1e-3, "windows_per_s": nwin / (ms * 1e-3),
           "final_allreduce_ms": ar_ms, "logit_map_gb": out.numel() * 4 / 1e9, "blocks_fused": fused,
           "finite": bool(torch.isfinite(out).all().item())}
    if check and rank == 0:
        # spot check against unsharded, unbatched evaluation of the same net on one interior region: the window at the
        # volume's origin corner is covered by exactly one window, so the blended result there IS that window's logits
        with torch.no_grad():
            corner = vol[..., :ROI, :ROI, :ROI]
            ref = predictor(corner, modalities=mod).float()
        # voxels [0, 48)^3 of the corner are covered only by window (0,0,0)
        h = ROI // 2
        a_, b_ = out[..., :h, :h, :h].float(), ref[..., :h, :h, :h]
        res["corner_rel_err_vs_single_window"] = float((a_ - b_).abs().max() / b_.abs().max())
    del net, vol, out
    torch.cuda.empty_cache()
    return res
