"""Pure-write and copy bandwidth of the device (is a store-bound kernel at the memory system's write limit?).
python tools/prof_hbm_write.py"""
import torch

dev = torch.device("cuda", 0)
n = 1 << 30      # 1 Gi floats = 4 GiB: far beyond the 126 MB L2
a = torch.empty(n, dtype=torch.float32, device=dev)
b = torch.empty(n, dtype=torch.float32, device=dev)


def t(fn, it=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(it):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


ms = t(lambda: a.fill_(1.0))
print(f"fill_ (pure write, 4 GiB): {ms:.3f} ms = {a.numel() * 4 / ms / 1e6:.0f} GB/s")
ms = t(lambda: a.zero_())
print(f"zero_ (memset, 4 GiB): {ms:.3f} ms = {a.numel() * 4 / ms / 1e6:.0f} GB/s")
ms = t(lambda: b.copy_(a))
print(f"copy_ (read + write, 4 + 4 GiB): {ms:.3f} ms = {2 * a.numel() * 4 / ms / 1e6:.0f} GB/s")
ms = t(lambda: a.sum())
print(f"sum (pure read, 4 GiB): {ms:.3f} ms = {a.numel() * 4 / ms / 1e6:.0f} GB/s")
