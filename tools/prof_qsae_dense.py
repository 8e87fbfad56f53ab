"""python tools/prof_qsae_dense.py [B] [iters] [exact] -- q_sae forward of an UNTRAINED model (~50 % of the latents
active: the dense path, quantizedsae_b200/sae/quantized_matryoshka.py::_forward_dense) timing driver (GPU box)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import quantizedsae_b200 as Q  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
exact = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
dev = torch.device("cuda:0")
D, H = 512, 32768
with torch.device(dev):
    torch.manual_seed(0)
    q = Q.QuantizedMatryoshkaSAE(D, H, 32, 4.0, 4)
q.eval()
q.exact = exact
q.dense_mode = "always"
x = torch.randn((B, D), device=dev)
if not exact:
    x = x.bfloat16().float()
with torch.no_grad():
    for _ in range(2):
        out = q(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        q(x)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"q_sae dense path B={B} exact={exact}: {ms:.3f} ms/forward = {B / ms / 1e3:.2f} M tokens/s, "
      f"{4.0 * B * H * D / ms / 1e9:.0f} TFLOP/s of 4 B D H, path={q.last_path}")
