"""PCIe ceiling for the host-buffer path (micn_fwd_bwd_host): pinned 170 MB up, 170 MB down, alone and together."""
import time
import torch

dev = torch.device("cuda:0")
nbytes = 2 * 48 * 96 ** 3 * 2  # x + dy of the headline workload, bf16
h_up = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
h_dn = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
d_up = torch.empty(nbytes, dtype=torch.uint8, device=dev)
d_dn = torch.empty(nbytes, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, down, chunks=1, reps=10):
    step = nbytes // chunks
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for k in range(chunks):
            sl = slice(k * step, (k + 1) * step)
            if up:
                with torch.cuda.stream(s1):
                    d_up[sl].copy_(h_up[sl], non_blocking=True)
            if down:
                with torch.cuda.stream(s2):
                    h_dn[sl].copy_(d_dn[sl], non_blocking=True)
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


for name, up, down, chunks in (("h2d", 1, 0, 1), ("d2h", 0, 1, 1), ("both", 1, 1, 1), ("both/12", 1, 1, 12), ("both/24", 1, 1, 24)):
    run(up, down, chunks, 2)
    t = run(up, down, chunks)
    print(f"{name:8s} {t * 1e3:7.3f} ms  {nbytes / t / 1e9:6.1f} GB/s per direction"
          f"  -> e2e ceiling {425e6 / t / 1e9:6.1f} GB/s", flush=True)


def chained(chunks, reps=10, gate_us=0.0):
    """down copy k may only start once up copy k has landed (the host path's dependency pattern, no kernels)"""
    step = nbytes // chunks
    evs = [torch.cuda.Event() for _ in range(chunks)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for k in range(chunks):
            sl = slice(k * step, (k + 1) * step)
            with torch.cuda.stream(s1):
                d_up[sl].copy_(h_up[sl], non_blocking=True)
                evs[k].record(s1)
            with torch.cuda.stream(s2):
                s2.wait_event(evs[k])
                h_dn[sl].copy_(d_dn[sl], non_blocking=True)
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


for chunks in (6, 12, 24, 48):
    chained(chunks, 2)
    t = chained(chunks)
    print(f"chained/{chunks:<3d} {t * 1e3:7.3f} ms  {nbytes / t / 1e9:6.1f} GB/s per direction", flush=True)
