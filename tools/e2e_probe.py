"""Development aid: where does the end-to-end step time go (H2D copy alone, step alone, prefetched loop)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import transformer_clip_eeg_b200 as pkg  # noqa: F401
from transformer_clip_eeg_b200 import train_clip_final as t
from transformer_clip_eeg_b200.optim import AdamW

B, T = 256, 320
dev = torch.device("cuda")
args = t.build_parser().parse_args([])
model = t.build_model(args, T, 10000, dev)
opt = AdamW(model.parameters(), lr=1e-3, weight_decay=0.01)
model.train()
host = [(torch.randn(B, T, 64).pin_memory(), torch.randn(B, T, 1024).pin_memory(), (torch.randperm(10000)[:B] + 1).pin_memory()) for _ in range(2)]
res = [tuple(x.to(dev) for x in h) for h in host]
for i in range(3):
    t.train_step(model, opt, *res[i % 2])
torch.cuda.synchronize()
def ev():
    return torch.cuda.Event(enable_timing=True)
# (a) copy alone
e0, e1 = ev(), ev(); e0.record()
for i in range(5):
    d = tuple(x.to(dev, non_blocking=True) for x in host[i % 2])
e1.record(); torch.cuda.synchronize()
print(f"H2D alone: {e0.elapsed_time(e1) / 5:.2f} ms per batch ({sum(x.numel() * x.element_size() for x in host[0]) / 1e6:.0f} MB)")
# (b) step alone, wall and device
t0 = time.perf_counter(); e0, e1 = ev(), ev(); e0.record()
for i in range(5):
    t.train_step(model, opt, *res[i % 2])
e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"step alone: device {e0.elapsed_time(e1) / 5:.2f} ms, host enqueue {(t1 - t0) / 5 * 1e3:.2f} ms, wall {(t2 - t0) / 5 * 1e3:.2f} ms")
class HB:
    def __init__(self, n): self.n = n
    def __iter__(self):
        for i in range(self.n):
            e, s, ids = host[i % 2]
            yield e, [s], ids, None
for with_item in (False, True):
    for _ in range(2):
        n = 10
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for eeg, sp, ids in t.DevicePrefetcher(HB(n), dev):
            l, _, _ = t.train_step(model, opt, eeg, sp, ids)
            if with_item:
                l.item()
        torch.cuda.synchronize(); t1 = time.perf_counter()
        print(f"prefetched loop (item={with_item}): {(t1 - t0) / n * 1e3:.2f} ms per step")
