"""Timing / profiling driver for the large-k path on one dictionary shard (GPU box).
python tests/prof_largek.py B H k iters [exact]   -- encode_topk + pack + merge(1 shard) + int4 decode"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from quantizedsae_b200 import _lib as L
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
H = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
k = int(sys.argv[3]) if len(sys.argv) > 3 else 2097
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 10
exact = len(sys.argv) > 5 and sys.argv[5] == "exact"
D = 512
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn((B, D), device=dev, generator=g).bfloat16().float()
W = ((torch.rand((H, D), device=dev, generator=g) * 2 - 1) * (6.0 / (H + D)) ** 0.5).bfloat16().float()
b = torch.zeros(H, device=dev)
Wb = L.cast_bf16(W)
packed = torch.randint(0, 256, (H, D // 2), dtype=torch.uint8, device=dev, generator=g)
sample = L.prepare_sample(Wb, b)


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        out = fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters, out


t_enc, (vals, idx, flags) = timed(lambda: L.encode_topk(x, Wb, W if exact else None, b, k, exact=exact, want_flags=True, sample=sample))
t_dec, _ = timed(lambda: L.decode_int4(vals, idx, packed, H, D, 0.5, None))
cand = L.pack_candidates(vals, idx).unsqueeze(0).contiguous()
t_mrg, _ = timed(lambda: L.merge_candidates(cand, H, k))
print(f"B={B} H={H} k={k} exact={exact}: encode_topk {t_enc*1e3:.0f} us ({2.0*B*H*D/t_enc/1e9:.0f} TFLOP/s), "
      f"merge(1 shard) {t_mrg*1e3:.0f} us, decode_int4 {t_dec*1e3:.0f} us; flagged rows {int((flags != 0).sum())}", flush=True)
