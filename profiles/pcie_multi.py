"""Aggregate pinned-copy bandwidth with every GPU of the box copying at once (run under torchrun):
the platform ceiling of the N-GPU `e2e` number.  Each rank moves 150 MB each way per iteration."""
import os
import subprocess
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("gloo")
dev = torch.device("cuda", local)
n = 150 * 1000 * 1000
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device=dev)
d_out = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def loop(iters, h2d=True, d2h=True):
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(iters):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    s1.synchronize()
    s2.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


loop(5)
for name, a, b in (("H2D only", True, False), ("D2H only", False, True), ("H2D + D2H", True, True)):
    dt = loop(40, a, b)
    if rank == 0:
        per_dir = world * 40 * n / dt / 1e9
        print(f"{world} GPUs, {name}: {per_dir:.1f} GB/s aggregate per direction ({per_dir / world:.1f} per GPU)"
              + (f" -> e2e ceiling {world * 40 * 4096 * 4096 / dt / 1e6:.0f} MPix/s" if a and b else ""))
if rank == 0:
    print(subprocess.run("nvidia-smi topo -m | head -14; lscpu | grep -E 'NUMA|^CPU\\(s\\)|Model name'", shell=True, capture_output=True, text=True).stdout)
dist.destroy_process_group()
