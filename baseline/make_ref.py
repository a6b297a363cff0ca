#!/usr/bin/env python
"""Recipe for the two git-ignored reference copies that travel to the GPU box with the gpurun snapshot
(`/root/reference` does not exist there).  Run by `__graft_entry__.build()` whenever /root/reference is present.

  oracle/_ref/conditional_instance_norm.py   the UNMODIFIED reference module of the hot path
                                              (networks/norms/conditional_instance_norm.py; needs only torch): what
                                              `bench.py --impl reference` and `cpu_baseline` time (kind "reference").
  baseline/_ref/networks/                     the UNMODIFIED reference `networks/` package (nets, blocks, layers, norms):
                                              what the model-level legs of bench.py build C-UNet / C-UNETR /
                                              C-Swin-UNETR from, through baseline/monai_stub.py (MONAI is absent).

Nothing is copied into tracked paths: both targets are listed in .gitignore, reference sources never enter history."""
from __future__ import annotations

import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def main() -> int:
    src_norm = os.path.join(REF, "networks", "norms", "conditional_instance_norm.py")
    if not os.path.isfile(src_norm):
        print("make_ref: /root/reference is not present; keeping whatever copies exist")
        return 0
    dst = os.path.join(ROOT, "oracle", "_ref")
    os.makedirs(dst, exist_ok=True)
    shutil.copyfile(src_norm, os.path.join(dst, "conditional_instance_norm.py"))
    net_dst = os.path.join(ROOT, "baseline", "_ref", "networks")
    if os.path.isdir(net_dst):
        for root, dirs, _files in os.walk(net_dst):
            os.chmod(root, 0o755)
        shutil.rmtree(net_dst)
    shutil.copytree(os.path.join(REF, "networks"), net_dst,
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "lightning_monai.py"))
    for root, dirs, files in os.walk(os.path.join(ROOT, "baseline", "_ref")):  # (the checkout is read-only; the copies need not be)
        for name in dirs + files:
            os.chmod(os.path.join(root, name), 0o755 if name in dirs else 0o644)
    os.chmod(os.path.join(dst, "conditional_instance_norm.py"), 0o644)
    print(f"make_ref: {dst}/conditional_instance_norm.py and {net_dst}/ written")
    return 0


if __name__ == "__main__":
    sys.exit(main())
