"""Live per-stage timing of one b_sae step (no profiler): CUDA events recorded by the library between the launches of
qsae_bsae_forward (qsae_set_stage_events). python tools/prof_stages.py [B] [k] [iters] [exact]"""
import ctypes as C
import statistics
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from quantizedsae_b200 import _lib as L  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
k = int(sys.argv[2]) if len(sys.argv) > 2 else 32
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 30
exact = len(sys.argv) > 4 and sys.argv[4] == "1"
H, D = 32768, 512
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
xs = [torch.randn((B, D), device=dev, generator=g).bfloat16().float() for _ in range(3)]
W = ((torch.rand((H, D), device=dev, generator=g) * 2 - 1) * (6.0 / (H + D)) ** 0.5).bfloat16().float()
b = torch.zeros(H, device=dev)
Wb = L.cast_bf16(W)
sample = L.prepare_sample(Wb, b)
packed = torch.randint(0, 256, (H, D // 2), dtype=torch.uint8, device=dev)
bd = torch.randn(D, device=dev)
lib = L.load()
NAMES = ["prior (cast + sample pre-pass + prior)", "sweep", "merge (+ fused decode)", "tail", "decode (separate)"]


def step(i):
    return L.bsae_forward(xs[i % 3], Wb, W if exact else None, b, k, packed, 4, 0.5, bd, exact=exact, sample=sample)


for i in range(5):
    step(i)
torch.cuda.synchronize()
acc = [[] for _ in NAMES]
tot = []
for i in range(iters):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    for e in ev:
        e.record()
    arr = (C.c_void_p * 6)(*[e.cuda_event for e in ev])
    L.check(lib.qsae_set_stage_events(arr, 6))
    step(i)
    L.check(lib.qsae_set_stage_events(None, 0))
    torch.cuda.synchronize()
    last = 4
    for j in range(4):
        acc[j].append(ev[j].elapsed_time(ev[j + 1]) * 1e3)
    try:
        d = ev[4].elapsed_time(ev[5]) * 1e3
        if d > 0:
            acc[4].append(d)
            last = 5
    except Exception:
        pass
    tot.append(ev[0].elapsed_time(ev[last]) * 1e3)
print(f"B={B} k={k} exact={exact}: stage medians in us (events between launches add ~1-2 us each)")
for n, a in zip(NAMES, acc):
    if a:
        print(f"  {n:45s} {statistics.median(a):8.1f}")
print(f"  {'sum of stages (first to last event)':45s} {statistics.median(tot):8.1f}")
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for i in range(iters):
    step(i)
e.record()
torch.cuda.synchronize()
print(f"  {'back-to-back step without events':45s} {s.elapsed_time(e) / iters * 1e3:8.1f}")
