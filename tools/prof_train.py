"""One call of each training-side entry at the BASELINE shapes (for an ncu launch list / --set full capture).
python tools/prof_train.py"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

import quantizedsae_b200 as Q
from quantizedsae_b200 import _lib as L

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
D, H, B = 512, 32768, 4096
g = torch.Generator(device=dev).manual_seed(1)
m = Q.BinarySAE(D, H, 4.0, 4).to(dev)
with torch.no_grad():
    m.decoder.weight.copy_(torch.randn((H, D * 4), device=dev, generator=g) * 1.5)
m.autograd, m.return_dense = True, False
x = torch.randn((B, D), device=dev, generator=g)
for _ in range(2):
    m.zero_grad(set_to_none=True)
    _, recon, pol = m(x)
    (0.5 * torch.nn.functional.mse_loss(recon, x) + 0.1 * pol).backward()
torch.cuda.synchronize()
w = torch.randn((D, H), device=dev, generator=g) * 0.4824
mask = torch.ones_like(w)
L.rigl_init_mask(w, mask, int(0.7 * w.numel()))
a_mean = torch.rand(H, device=dev, generator=g) + 1e-3
d_mean = torch.randn(D, device=dev, generator=g)
n = int(0.1 * 0.3 * w.numel())
for _ in range(2):
    L.rigl_update_mask(w, mask, a_mean, d_mean, n, n)
torch.cuda.synchronize()
print("done")
