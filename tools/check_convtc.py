"""Development aid: tcgen05 conv block vs exact-fp32 path (prints relative errors and timings)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import transformer_clip_eeg_b200 as pkg
from transformer_clip_eeg_b200 import clip_model as cm, _lib

torch.manual_seed(0)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 320
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = "cuda"
blk = cm.BasicBlock(64, 64, kernel_size=64, time_dimension=T).to(dev).eval()
x = torch.randn(B, T, 64, device=dev); skip = torch.randn(B, T, 64, device=dev); w = torch.randn(B, T, 64, device=dev)
res = {}
for m in ("fp32", "bf16x3", "bf16"):
    _lib.set_default_math(m)
    xx = x.clone().requires_grad_(True)
    blk.zero_grad()
    y = blk.forward_time_major(xx, skip)
    (y * w).sum().backward()
    torch.cuda.synchronize()
    res[m] = (y.detach(), xx.grad.detach(), blk.conv.weight.grad.detach().clone(), blk.conv.bias.grad.detach().clone())
    print(m, "done", flush=True)
def rel(a, b): return float((a - b).norm() / b.norm())
for m in ("bf16x3", "bf16"):
    print(m, {n: f"{rel(a, b):.2e}" for a, b, n in zip(res[m], res["fp32"], ("y", "dx", "dw", "db"))})
if B >= 64:
    for m in ("fp32", "bf16x3", "bf16"):
        _lib.set_default_math(m)
        for phase in ("fwd", "fwd+bwd"):
            ts = []
            for it in range(4):
                xx = x.clone().requires_grad_(True)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                y = blk.forward_time_major(xx, skip)
                if phase != "fwd":
                    (y * w).sum().backward()
                e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            print(f"{m} {phase}: {min(ts):.3f} ms (B={B}, T={T})")
