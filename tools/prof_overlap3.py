"""Experiment: b_sae step split into encode (prior, sweep, merge, tail) on two alternating streams and the int4 decode on a
third stream (optionally lower priority), B = 4096.   python tools/prof_overlap3.py [B] [k] [steps]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

import bench
from quantizedsae_b200 import _lib as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
k = int(sys.argv[2]) if len(sys.argv) > 2 else 32
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
We, be, logits, bd = bench.make_weights(torch, dev)
w_bf16 = L.cast_bf16(We)
sample = L.prepare_sample(w_bf16, be)
packed, _, _ = L.pack_bitplanes(logits, bench.D, bench.N_BITS)
del logits
n_in = bench.n_rotating(B)
n_in -= n_in % 2
xs = [bench.make_x(torch, dev, B, s) for s in range(n_in)]
q = bench.GAMMA / 2 ** (bench.N_BITS - 1)
H, D = bench.H, bench.D


def enc(i):
    return L.encode_topk(xs[i % n_in], w_bf16, None, be, k, L.ACT_NONE, False, sample=sample)


def fused(i):
    return L.bsae_forward(xs[i % n_in], w_bf16, None, be, k, packed, bench.N_BITS, q, bd, exact=False, sample=sample)


for i in range(4):
    v, ix, _ = enc(i)
    L.decode_int4(v, ix, packed, H, D, q, bd)
    fused(i)
torch.cuda.synchronize()
ref = fused(0)
ref_recon = ref[3].clone()

lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
for name, prio_dec in (("decode stream at equal priority", 0), ("decode stream at LOW priority, encode streams HIGH", 1)):
    sA = [torch.cuda.Stream(priority=-1 if prio_dec else 0) for _ in range(2)]
    sC = torch.cuda.Stream(priority=0)
    pool = torch.cuda.graph_pool_handle()
    g_enc, g_dec = [], []
    for i in range(n_in):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, pool=pool, stream=sA[i % 2]):
            v, ix, _ = enc(i)
        g_enc.append((g, v, ix))
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g2, pool=pool, stream=sC):
            rec = L.decode_int4(v, ix, packed, H, D, q, bd)
        g_dec.append((g2, rec))
    main = torch.cuda.current_stream()

    def run(nsteps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for s in sA + [sC]:
            s.wait_event(e0)
        for i in range(nsteps):
            s = sA[i % 2]
            with torch.cuda.stream(s):
                g_enc[i % n_in][0].replay()
                ev = torch.cuda.Event()
                ev.record(s)
            sC.wait_event(ev)
            with torch.cuda.stream(sC):
                g_dec[i % n_in][0].replay()
            if i >= n_in - 2:                       # buffer reuse: encode i + n_in overwrites (v, ix) of step i: wait for its decode
                evd = torch.cuda.Event()
                evd.record(sC)
                sA[i % 2].wait_event(evd)
        for s in sA + [sC]:
            ev = torch.cuda.Event()
            ev.record(s)
            main.wait_event(ev)
        e1.record(main)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    run(20)
    ms = min(run(steps) for _ in range(3)) / steps
    ok = torch.equal(g_dec[0][1], ref_recon)
    print(f"{name}: {ms * 1e3:.1f} us/step = {B / ms / 1e3:.2f} M tokens/s, recon identical to the fused step: {ok}", flush=True)
    del g_enc, g_dec
