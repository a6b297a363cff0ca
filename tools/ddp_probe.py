"""Experiment: where does the DDP step lose time?  torchrun --nproc-per-node N tools/ddp_probe.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import mi_seg_b200 as pkg  # noqa: E402
import model_bench as MB  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
variants = [("", ("ours", "reference")), ("bucket_view", ("ours",)), ("bucket_view,bucket_cap=300", ("ours",)),
            ("no_find_unused,bucket_view", ("ours",))]
os.environ["MICN_PROFILE_DDP"] = "1"
for v, which in variants:
    os.environ["MICN_DDP_OPTS"] = v
    for variant in which:
        r = MB.train_step_bench("swin_unetr", variant, pkg, dev, world, rank, 8, 3, 1)
        if rank == 0:
            print(json.dumps({"ddp_opts": v, "variant": variant, "ms_per_step": r["ms_per_step"], "norm_share": r.get("norm_share")}), flush=True)
dist.destroy_process_group()
