"""python tests/prof_qsae.py [B] [iters] [exact] -- q_sae forward timing driver (GPU box)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import quantizedsae_b200 as Q  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
exact = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
dev = torch.device("cuda:0")
D, H = 512, 32768
with torch.device(dev):
    torch.manual_seed(0)
    q = Q.QuantizedMatryoshkaSAE(D, H, 32, 4.0, 4)
    with torch.no_grad():
        q.encoder[0].weight.copy_(q.encoder[0].weight.bfloat16().float())
        q.encoder[0].bias.fill_(-0.543)
q.eval()
q.exact = exact
x = torch.randn((B, D), device=dev).bfloat16().float()
with torch.no_grad():
    for _ in range(2):
        q(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        q(x)
    e1.record()
    torch.cuda.synchronize()
print(f"q_sae B={B} exact={exact}: {e0.elapsed_time(e1) / iters:.3f} ms/forward")
