"""Device timing of the VLAAI baseline forward / forward+backward (BASELINE config 5; development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import transformer_clip_eeg_b200 as pkg  # noqa: F401
from transformer_clip_eeg_b200 import vlaai, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
maths = sys.argv[2].split(",") if len(sys.argv) > 2 else ["bf16x3"]
FWD, FWDBWD = 32920e6, 98758e6          # FLOP per sample
torch.manual_seed(0)
model = vlaai.VLAAI().to("cuda").train()
x = torch.randn(B, 320, 64, device="cuda")
for m in maths:
    _lib.set_default_math(m)
    def fwd():
        with torch.no_grad():
            return model(x)
    def fb():
        model.zero_grad(set_to_none=True)
        model(x).sum().backward()
    for name, fn, flop in (("fwd", fwd, FWD), ("fwd+bwd", fb, FWDBWD)):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 3
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"VLAAI {m} {name}: B={B}: {ms:.2f} ms -> {B / ms * 1e3:.0f} samples/s, {B * flop / ms / 1e9:.1f} TFLOP/s algorithmic")
