"""Forward timings of every module of the hot path through the nn.Module API (GPU box):

    python tests/prof_modules.py [iters]

CUDA events on the current stream, 3 warm-up forwards, inputs rotate over 3 buffers. Prints one JSON
line per configuration (BASELINE.json configs 1-4 shapes; config 5 is bench.py --variant dict-sharded)."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import quantizedsae_b200 as Q  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda:0")
D, H = 512, 32768


def bf16r(t):
    return t.bfloat16().float()


def time_forward(m, B, label, flops_per_token, **note):
    g = torch.Generator(device=dev).manual_seed(1)
    xs = [bf16r(torch.randn((B, D), device=dev, generator=g)) for _ in range(3)]
    with torch.no_grad():
        for i in range(3):
            m(xs[i])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            m(xs[i % 3])
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(json.dumps({"config": label, "batch": B, "ms_per_forward": round(ms, 4), "tokens_per_s": round(B / ms * 1e3),
                      "tflops_algorithmic": round(flops_per_token * B / ms / 1e9, 1), **note}), flush=True)


with torch.device(dev):
    torch.manual_seed(0)
    # config 1: b_sae, polarised logits, k = 65 (reference default) and k = 32
    b = Q.BinarySAE(D, H, 4.0, 4)
    with torch.no_grad():
        b.encoder[0].weight.copy_(bf16r(b.encoder[0].weight))
        b.decoder.weight.copy_(torch.where(torch.rand_like(b.decoder.weight) < 0.5, 110.0, -110.0))
        b.decoder.bias.normal_()
    b.eval()
    for exact in (False, True):
        for kk in (32, 65):
            for rd in (False, True):
                for B in (4096, 65536):
                    if rd and B == 65536:
                        continue          # dense [65536, 32768] fp32 latents = 8.6 GB per forward: skipped
                    b.k, b.exact, b.return_dense = kk / H, exact, rd
                    time_forward(b, B, "b_sae 512->32768 n_bits=4", 2.0 * D * H, k=kk, exact=exact, dense_latents=rd)
    del b
    # config 2: baseline_sae
    bl = Q.BaselineSparseAutoencoder(D, H)
    with torch.no_grad():
        bl.encoder[0].weight.copy_(bf16r(bl.encoder[0].weight))
    bl.eval()
    bl.return_dense = False
    for exact in (False, True):
        bl.exact = exact
        time_forward(bl, 65536, "baseline_sae 512->32768 top_k=32", 2.0 * D * H, exact=exact, dense_latents=False)
    del bl
    # config 3: t_sae
    t = Q.TernarySparseAutoencoder(D, H)
    with torch.no_grad():
        t.encoder[0].weight.copy_(bf16r(t.encoder[0].weight))
        t.decoder.weight.normal_(0, 0.4824)
    t.eval()
    t.exact = False
    for B in (4096, 16384):
        time_forward(t, B, "t_sae 512->32768 dense", 4.0 * D * H, exact=False)
    t.exact = True
    time_forward(t, 1024, "t_sae 512->32768 dense", 4.0 * D * H, exact=True)
    del t
    # config 4: q_sae, trained-like sparsity (encoder bias -0.543) and untrained (dense path)
    q = Q.QuantizedMatryoshkaSAE(D, H, 32, 4.0, 4)
    with torch.no_grad():
        q.encoder[0].weight.copy_(bf16r(q.encoder[0].weight))
        q.encoder[0].bias.fill_(-0.543)
    q.eval()
    for exact in (False, True):
        q.exact = exact
        for B in (4096, 65536):
            time_forward(q, B, "q_sae 512->32768 n_bits=4 L0~34", 2.0 * D * H, exact=exact, path="sparse")
    with torch.no_grad():
        q.encoder[0].bias.zero_()
    q.exact, q.dense_mode = False, "always"
    time_forward(q, 4096, "q_sae 512->32768 n_bits=4 untrained (50% active)", 4.0 * D * H, exact=False, path="dense")
