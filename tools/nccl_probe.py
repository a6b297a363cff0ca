"""bring-up: what does a tiny NCCL all-reduce cost on this box, alone and between our kernels?
torchrun --nproc-per-node 2 tools/nccl_probe.py"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank = dist.get_rank()
g = torch.zeros(2, 2, 48, device=dev)
big = torch.zeros(64 << 20, device=dev)


def timeit(fn, n=100):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


t_small = timeit(lambda: dist.all_reduce(g))
t_big = timeit(lambda: dist.all_reduce(big), 10)
x = torch.randn(1 << 24, device=dev)
t_k = timeit(lambda: x.mul_(1.0001))


def mixed():
    x.mul_(1.0001)
    dist.all_reduce(g)


t_mixed = timeit(mixed)
if rank == 0:
    print(f"all_reduce 192 floats: {t_small:.1f} us; all_reduce 256 MB: {t_big:.1f} us "
          f"({2 * big.numel() * 4 / t_big * 1e-3 * (dist.get_world_size() - 1) / dist.get_world_size():.1f} GB/s bus); "
          f"elementwise kernel: {t_k:.1f} us; kernel + small all_reduce: {t_mixed:.1f} us")
dist.destroy_process_group()
