"""Device timing of the match-mismatch scoring kernels (BASELINE config 4): candidate row-dots and the N x M bank similarity."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import transformer_clip_eeg_b200 as pkg  # noqa: F401
from transformer_clip_eeg_b200 import train_clip_helper_functions as H

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

N, D = 4096, 2560
torch.manual_seed(0)
E = torch.randn(N, D, device="cuda")
for K in (2, 5, 100):
    C = torch.randn(N, K, D, device="cuda")
    ms = timed(lambda: H.mm_scores(E, C))
    print(f"row-dots N={N} K={K} D={D}: {ms * 1e3:8.1f} us, {N * K * D * 4 / ms / 1e6:7.0f} GB/s of candidates")
    del C
for M in (1000, 10000, 100000):
    Bk = torch.randn(M, D, device="cuda")
    ms = timed(lambda: H.bank_logits(E, Bk))
    ms_k = timed(lambda: H.bank_topk(E, Bk, 100))
    print(f"bank N={N} M={M} D={D}: logits {ms:7.2f} ms ({2 * N * M * D / ms / 1e9:6.1f} TFLOP/s algorithmic), with top-100 {ms_k:7.2f} ms")
    del Bk
