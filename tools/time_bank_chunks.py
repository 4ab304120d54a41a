import os, sys
sys.path.insert(0, os.getcwd())
import torch
import transformer_clip_eeg_b200
from transformer_clip_eeg_b200 import train_clip_helper_functions as H
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
N, D, M = 4096, 2560, 100000
E = torch.nn.functional.normalize(torch.randn(N, D, device="cuda"), dim=1)
Bk = torch.randn(M, D, device="cuda")
print("logits only", timed(lambda: H.bank_logits(E, Bk)))
for ch in (2048, 4096, 8192, 16384, 32768, 100000):
    print(ch, timed(lambda: H.bank_topk(E, Bk, 100, chunk=ch)))
x = H.bank_logits(E, Bk)
print("row_topk alone on full logits", timed(lambda: H.row_topk(x, 100)))
