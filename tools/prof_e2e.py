"""Host-buffer pipeline timing (GPU box): python tests/prof_e2e.py B k chunk [chunk ...]
Prints the PCIe copy times of the same buffers next to the pipeline's wall time."""
import ctypes as C
import sys
import time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench as BN
from quantizedsae_b200 import _lib as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
k = int(sys.argv[2]) if len(sys.argv) > 2 else 32
chunks = [int(a) for a in sys.argv[3:]] or [8192, 16384]
dev = torch.device("cuda:0")
lib = L.load()
We, be, logits, bd = BN.make_weights(torch, dev)
D, H = BN.D, BN.H
hx = BN.make_x(torch, dev, B, 0).cpu().pin_memory()
hv = torch.empty((B, k), dtype=torch.float32).pin_memory()
hi = torch.empty((B, k), dtype=torch.int32).pin_memory()
hr = torch.empty((B, D), dtype=torch.float32).pin_memory()
dx = torch.empty((B, D), device=dev)
dr = torch.empty((B, D), device=dev)


def wall(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


t_in = wall(lambda: dx.copy_(hx, non_blocking=True))
t_out = wall(lambda: hr.copy_(dr, non_blocking=True))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def both():
    with torch.cuda.stream(s1):
        dx.copy_(hx, non_blocking=True)
    with torch.cuda.stream(s2):
        hr.copy_(dr, non_blocking=True)


t_both = wall(both)
print(f"PCIe: H2D {hx.numel()*4/1e6:.0f} MB {t_in:.2f} ms ({hx.numel()*4/t_in/1e6:.1f} GB/s), D2H {hr.numel()*4/1e6:.0f} MB {t_out:.2f} ms "
      f"({hr.numel()*4/t_out/1e6:.1f} GB/s), both directions at once {t_both:.2f} ms", flush=True)
for ch in chunks:
    plan = C.c_void_p()
    L.check(lib.qsae_bsae_plan_create(We.data_ptr(), be.data_ptr(), logits.data_ptr(), bd.data_ptr(), H, D, BN.N_BITS,
                                      C.c_float(BN.GAMMA), k, ch, C.byref(plan)))
    t = wall(lambda: L.check(lib.qsae_bsae_forward_host(plan, hx.data_ptr(), B, hv.data_ptr(), hi.data_ptr(), hr.data_ptr())))
    lib.qsae_bsae_plan_destroy(plan)
    print(f"forward_host B={B} k={k} max chunk {ch}: {t:.2f} ms = {B/t/1e3:.2f} Mtok/s", flush=True)
