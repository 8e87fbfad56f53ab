"""Sweep kernel time at B=65536 (events around the sweep launch) and step time. python tools/prof_sweep65536.py [B] [k]"""
import statistics, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from quantizedsae_b200 import _lib as L
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
k = int(sys.argv[2]) if len(sys.argv) > 2 else 32
H, D = 32768, 512
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
xs = [torch.randn((B, D), device=dev, generator=g).bfloat16().float() for _ in range(3)]
W = ((torch.rand((H, D), device=dev, generator=g) * 2 - 1) * (6.0 / (H + D)) ** 0.5).bfloat16().float()
b = torch.zeros(H, device=dev)
Wb = L.cast_bf16(W); sample = L.prepare_sample(Wb, b)
packed = torch.randint(0, 256, (H, D // 2), dtype=torch.uint8, device=dev); bd = torch.randn(D, device=dev)
lib = L.load()
def step(i): return L.bsae_forward(xs[i % 3], Wb, None, b, k, packed, 4, 0.5, bd, sample=sample)
for i in range(5): step(i)
torch.cuda.synchronize()
ks = []
for i in range(20):
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); c.record()
    L.check(lib.qsae_set_encode_kernel_events(a.cuda_event, c.cuda_event))
    step(i)
    L.check(lib.qsae_set_encode_kernel_events(None, None))
    torch.cuda.synchronize()
    ks.append(a.elapsed_time(c))
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for i in range(20): step(i)
e.record(); torch.cuda.synchronize()
print(f"B={B} k={k}: sweep kernel {statistics.median(ks)*1e3:.1f} us, step {s.elapsed_time(e)/20*1e3:.1f} us")
