"""python tests/prof_rqsae.py [B] [iters] [exact] -- rq_sae forward timing driver (GPU box)."""
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
    m = Q.ResidualQuantizedSAE(D, H, 32, 4.0, 4)
    with torch.no_grad():
        for s in m.saes:
            s.encoder[0].weight.copy_(s.encoder[0].weight.bfloat16().float())
            s.encoder[0].bias.fill_(-0.6)
m.eval()
m.exact = exact
x = torch.randn((B, D), device=dev).bfloat16().float()
# calibrate the encoder biases stage by stage so that ~0.1 % of a stage's latents fire (a trained model's regime;
# at the default initialisation more than half of the latents are active and every stage takes the dense path)
with torch.no_grad():
    r = x[:4096]
    for s in m.saes:
        z = r @ s.encoder[0].weight.t()
        s.encoder[0].bias.fill_(float(-3.1 * z.std()))
        _, lv = s(r)
        r = (r - lv[-1]) * 2
with torch.no_grad():
    for _ in range(2):
        g, r = m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g, r = m(x)
    e1.record()
    torch.cuda.synchronize()
t = e0.elapsed_time(e1) / iters
print(f"rq_sae B={B} exact={exact}: {t:.3f} ms/forward = {B / t / 1e3:.2f} Mtok/s; L0 per stage {[round(float(v), 1) for v in g]}, "
      f"paths {[s.last_path for s in m.saes]}")
