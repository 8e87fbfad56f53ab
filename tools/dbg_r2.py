"""Round-2 debugging helper: baseline_sae non-exact timing at B=65536 and q_sae parity spot check."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import quantizedsae_b200 as Q
from quantizedsae_b200 import _lib as L
from oracle import qsae_oracle as O
dev = torch.device("cuda:0")
D, H = 512, 32768
def tm(fn, n=10, w=3):
    for _ in range(w): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
what = sys.argv[1] if len(sys.argv) > 1 else "baseline"
if what == "baseline":
    B = 65536
    torch.manual_seed(0)
    with torch.device(dev):
        m = Q.BaselineSparseAutoencoder(D, H)
    with torch.no_grad():
        m.encoder[0].weight.copy_(m.encoder[0].weight.bfloat16().float())
    m.eval(); m.return_dense = False
    x = torch.randn((B, D), device=dev).bfloat16().float()
    lin = m.encoder[0]
    with torch.no_grad():
        for exact in (False, True, False):
            m.exact = exact
            print("module forward exact", exact, tm(lambda: m(x)), "ms")
        wb, sample = m._w_bf16(), m._sample()
        print("encode_topk k=32", tm(lambda: L.encode_topk(x, wb, None, lin.bias.detach(), 32, sample=sample)))
        print("encode_topk k=32 zero bias", tm(lambda: L.encode_topk(x, wb, None, torch.zeros_like(lin.bias), 32, sample=(sample[0], torch.zeros_like(sample[1])))))
        v, i, _ = L.encode_topk(x, wb, None, lin.bias.detach(), 32, sample=sample)
        rows = m._dec_rows()
        print("decode_rows_f32", tm(lambda: L.decode_rows_f32(v, i, rows, H, D, 1.0, m.decoder.bias.detach())))
else:
    B = 4096
    torch.manual_seed(0)
    with torch.device(dev):
        m = Q.QuantizedMatryoshkaSAE(D, H, 32, abs_range=4.0, n_bits=4)
    with torch.no_grad():
        m.encoder[0].weight.copy_(m.encoder[0].weight.bfloat16().float())
        m.encoder[0].bias.fill_(-0.543)
    m.eval(); m.exact = False
    x = torch.randn((B, D), device=dev).bfloat16().float()
    with torch.no_grad():
        groups, levels = m(x)
    rows = np.arange(0, B, B // 16)[:16]
    lg, res, act = O.qsae_forward(x[rows].cpu().numpy(), m.encoder[0].weight.detach().cpu().numpy(), m.encoder[0].bias.detach().cpu().numpy(),
                                  m.decoder.weight.detach().cpu().numpy(), m.decoder.weight_mirror.detach().cpu().numpy(),
                                  m.decoder.bias.detach().cpu().numpy(), n_bits=4, abs_range=4.0)
    for i in range(4):
        got = levels[i][rows].cpu().numpy()
        d = np.abs(got - res[i])
        print("level", i, "max abs diff", d.max(), "rms ref", np.sqrt((res[i].astype(np.float64) ** 2).mean()), "rows bad", np.nonzero(d.max(1) > 1e-3)[0])
    print("oracle L0 per row", act.sum(1))
    wb = m._w_bf16()
    be = m.encoder[0].bias.detach()
    ztc = L.encode_dense_tc(x[rows].contiguous(), wb, be).cpu().numpy()
    znp = O.encode_pre(x[rows].cpu().numpy(), m.encoder[0].weight.detach().cpu().numpy(), be.cpu().numpy())
    print("z max abs diff", np.abs(ztc - znp).max())
    flip = (ztc >= 8.94e-8) != (znp >= 8.94e-8)
    print("flipped latents (row, col):", np.argwhere(flip)[:10], ztc[flip][:10], znp[flip][:10])
    # which latents does the module consider active? use forward_active
    with torch.no_grad():
        fa = m.forward_active(x, 256)
    ai = fa["active_idx"][rows].cpu().numpy(); ac = fa["active_cnt"][rows].cpu().numpy()
    for r in (3, 10):
        mine = set(ai[r][:ac[r]].tolist()); ref = set(np.nonzero(act[r])[0].tolist())
        print("row", r, "only ours", sorted(mine - ref), "only oracle", sorted(ref - mine), [znp[r, c] for c in sorted(ref - mine)], [ztc[r, c] for c in sorted(ref - mine)])
    fl = fa["reconstruction_levels"]
    for i in range(4):
        print("forward_active level", i, "max diff vs oracle", np.abs(fl[i][rows].cpu().numpy() - res[i]).max(),
              "vs forward()", float((fl[i] - levels[i]).abs().max()))
    with torch.no_grad():
        g2, l2 = m(x)
    print("second forward vs first", [float((l2[i] - levels[i]).abs().max()) for i in range(4)])
    bad = (levels[1] - fl[1]).abs().max(1).values.nonzero().flatten()
    print("rows where forward() != forward_active() at level 1:", bad[:20].tolist(), "count", bad.numel(), "of", B)
    r = 3 * (B // 16)
    d = (levels[1][r] - fl[1][r]).cpu().numpy()
    print("row", r, "nonzero diff positions", np.count_nonzero(d), "values", np.unique(np.round(d[d != 0], 5))[:6])
