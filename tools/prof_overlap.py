"""Experiment: b_sae steps of consecutive batches in flight on S streams (graph replay), B = 4096.
python tools/prof_overlap.py [B] [k] [steps]"""
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
xs = [bench.make_x(torch, dev, B, s) for s in range(n_in)]
q = bench.GAMMA / 2 ** (bench.N_BITS - 1)


def step(i):
    return L.bsae_forward(xs[i % n_in], w_bf16, None, be, k, packed, bench.N_BITS, q, bd, exact=False, sample=sample)


for i in range(5):
    step(i)
torch.cuda.synchronize()
ref = [t.clone() for t in step(0) if t is not None]

for S in (1, 2, 3, 4):
    streams = [torch.cuda.Stream() for _ in range(S)]
    pool = torch.cuda.graph_pool_handle()
    graphs = []
    for i in range(n_in * S // S):            # graph i is captured on (and replayed on) stream i % S
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, pool=pool, stream=streams[i % S]):
            out = step(i)
        graphs.append((g, out))
    n = len(graphs) - len(graphs) % S         # keep graph -> stream assignment fixed over the rotation
    main = torch.cuda.current_stream()

    def run(nsteps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for s in streams:
            s.wait_event(e0)
        for i in range(nsteps):
            with torch.cuda.stream(streams[i % S]):
                graphs[i % n][0].replay()
        for s in streams:
            ev = torch.cuda.Event()
            ev.record(s)
            main.wait_event(ev)
        e1.record(main)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    run(20)
    ms = min(run(steps) for _ in range(3)) / steps
    got = graphs[0][1]
    ok = all(torch.equal(a, b) for a, b in zip(ref, [t for t in got if t is not None]))
    print(f"S={S}: {ms * 1e3:.1f} us/step = {B / ms / 1e3:.2f} M tokens/s, results identical to the single-stream step: {ok}", flush=True)
    del graphs
