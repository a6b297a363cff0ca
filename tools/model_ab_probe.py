"""A/B of the C-Swin-UNETR training step (N=1) under library / binding variants, same process, same box."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import mi_seg_b200 as pkg  # noqa: E402
import model_bench as MB  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
kinds = sys.argv[1:] or ["swin_unetr"]
for kind, batch in (("swin_unetr", 1), ("unetr", 4)):
    if kind not in kinds:
        continue
    for name, binding, opts in (("cpp", "cpp", {}), ("ctypes", "ctypes", {}), ("cpp res_off", "cpp", {"res_off": 1}),
                                ("cpp", "cpp", {}), ("ctypes res_off", "ctypes", {"res_off": 1})):
        pkg.set_binding(binding)
        for k, v in opts.items():
            pkg._lib.set_option(k, v)
        r = MB.train_step_bench(kind, "ours", pkg, dev, 1, 0, 10, 3, batch)
        for k in opts:
            pkg._lib.set_option(k, -1)
        print(json.dumps({"kind": kind, "variant": name, "ms_per_step": r["ms_per_step"], "norm_share": r.get("norm_share")}), flush=True)
