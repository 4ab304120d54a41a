"""Measure write-only, read-only and copy HBM bandwidth with torch ops (context for the HBM-bound kernels' achieved GB/s)."""
import torch
n = 1 << 30            # 1 Gi floats = 4 GiB
x = torch.empty(n, dtype=torch.float32, device="cuda")
y = torch.empty(n, dtype=torch.float32, device="cuda")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: x.fill_(1.0)); print(f"write-only (fill_)   : {4 * n / ms / 1e6:8.0f} GB/s")
ms = t(lambda: x.sum());      print(f"read-only  (sum)     : {4 * n / ms / 1e6:8.0f} GB/s")
ms = t(lambda: y.copy_(x));   print(f"copy (read + write)  : {8 * n / ms / 1e6:8.0f} GB/s")
ms = t(lambda: torch.add(x, 1.0, out=y)); print(f"1 read + 1 write     : {8 * n / ms / 1e6:8.0f} GB/s")
ms = t(lambda: torch.add(x, y, out=y));   print(f"2 reads + 1 write    : {12 * n / ms / 1e6:8.0f} GB/s")
