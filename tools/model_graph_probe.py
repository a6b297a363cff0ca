import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "baseline"))
import mi_seg_b200 as pkg
import model_bench as MB
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
for kind, b in (("swin_unetr", 1), ("unetr", 4)):
    try:
        r = MB.graphed_step_bench(kind, pkg, dev, 10, 3, b)
    except Exception as e:
        import traceback; traceback.print_exc(); r = {"error": repr(e)[:300]}
    print(kind, json.dumps(r)[:600], flush=True)
    r2 = MB.train_step_bench(kind, "ours", pkg, dev, 1, 0, 10, 3, b, profile_share=False)
    print(kind, "eager ours", r2["ms_per_step"], flush=True)
