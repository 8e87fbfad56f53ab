"""Timing of the analysis kernels at the headline shape (GPU box): python tests/prof_analysis.py [B] [k]"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from quantizedsae_b200 import _lib as L
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
k = int(sys.argv[2]) if len(sys.argv) > 2 else 32
H, D = 32768, 512
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
# Zipf-like feature popularity (a few hot latents), like trained SAEs
p = 1.0 / torch.arange(1, H + 1, device=dev, dtype=torch.float32) ** 0.7
idx = torch.multinomial(p.expand(1024, H), k, replacement=False, generator=g).to(torch.int32).repeat(B // 1024, 1).contiguous()
vals = torch.rand((B, k), device=dev, generator=g)
x = torch.randn((B, D), device=dev, generator=g)
r = torch.randn((B, D), device=dev, generator=g)
counts = torch.zeros(H, dtype=torch.int64, device=dev)
cooc = torch.zeros((H, H), dtype=torch.int32, device=dev)
acc = torch.zeros((), dtype=torch.float64, device=dev)


def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n


t1 = timed(lambda: L.activation_counts(idx, vals, counts))
t2 = timed(lambda: L.coactivation(idx, vals, cooc))
t3 = timed(lambda: L.sq_error_accumulate(r, x, acc))
dense_flops = 2.0 * H * H * B
print(f"B={B} k={k} H={H}: activation_counts {t1*1e3:.0f} us, coactivation {t2*1e3:.0f} us "
      f"({B*k*k/t2/1e6:.1f} G increments/s; the reference's dense mask^T mask would be {dense_flops/1e12:.0f} TFLOP), "
      f"sq_error {t3*1e3:.0f} us ({2*B*D*4/t3/1e6:.0f} GB/s)", flush=True)
