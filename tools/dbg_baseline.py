import sys, ctypes as C
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench
import quantizedsae_b200 as Q
from quantizedsae_b200 import _lib as L
dev = torch.device("cuda:0"); D, H, B = 512, 32768, 65536
torch.manual_seed(0)
with torch.device(dev):
    m = Q.BaselineSparseAutoencoder(D, H)
with torch.no_grad():
    m.encoder[0].weight.copy_(m.encoder[0].weight.bfloat16().float())
m.eval(); m.return_dense, m.exact = False, False
rot = int(sys.argv[1]) if len(sys.argv) > 1 else 3
xs = [bench.make_x(torch, dev, B, 50 + s) for s in range(rot)]
lib = L.load()
names = ["prior", "sweep", "merge", "tail"]
with torch.no_grad():
    for i in range(4): m(xs[i % rot])
    torch.cuda.synchronize()
    for i in range(6):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        for e in ev: e.record()
        arr = (C.c_void_p * 6)(*[e.cuda_event for e in ev])
        L.check(lib.qsae_set_stage_events(arr, 6))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); lat, rec = m(xs[i % rot]); e1.record()
        L.check(lib.qsae_set_stage_events(None, 0))
        torch.cuda.synchronize()
        print(i % rot, " ".join(f"{n} {ev[j].elapsed_time(ev[j+1])*1e3:.0f}" for j, n in enumerate(names)), "total", round(e0.elapsed_time(e1) * 1e3))
