"""Host <-> device copy bandwidth with 1, 2, 4, ... ranks active at once (no kernels): where does the end-to-end
path of bench.py saturate? Run under torchrun with as many ranks as GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 tools/prof_pcie.py

Every rank pins its process to its GPU's NUMA-local CPUs (as bench.py does), allocates pinned buffers, and for each
group size n the first n ranks copy concurrently (the others wait at the barrier): H2D alone, D2H alone, both at once on
two streams. Prints one table on rank 0 (per-rank GB/s: min / mean over the active ranks, and the aggregate)."""
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
numa = bench.bind_to_gpu_numa_node(local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
MB = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n = MB * 1024 * 1024
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device=dev)
d_out = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
REPS = 10


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def run(mode):
    barrier()
    t0 = time.perf_counter()
    for _ in range(REPS):
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return time.perf_counter() - t0


sizes = [g for g in (1, 2, 4, 8) if g <= world]
rows = []
for g in sizes:
    for mode in ("h2d", "d2h", "both"):
        active = rank < g
        if active:
            run(mode)           # warm-up
            dt = run(mode)
        else:
            barrier(); barrier()
            dt = 0.0
        nbytes = n * REPS * (2 if mode == "both" else 1)
        gbs = nbytes / dt / 1e9 if active else 0.0
        t = torch.tensor([gbs], device=dev, dtype=torch.float64)
        if world > 1:
            allv = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allv, t)
            vals = [float(v[0]) for v in allv][:g]
        else:
            vals = [gbs]
        rows.append((g, mode, min(vals), sum(vals) / len(vals), sum(vals)))
        barrier()
if rank == 0:
    print(f"pinned buffers of {MB} MiB, {REPS} copies per measurement; numa binding of rank 0: {numa}")
    print("| active ranks | direction | per-rank GB/s (min) | per-rank GB/s (mean) | aggregate GB/s |")
    print("|---|---|---|---|---|")
    for g, mode, mn, mean, tot in rows:
        print(f"| {g} | {mode} | {mn:.1f} | {mean:.1f} | {tot:.1f} |")
if world > 1:
    dist.destroy_process_group()
