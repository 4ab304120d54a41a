"""One bench-configuration train step bracketed by cudaProfilerStart/Stop (for `ncu --profile-from-start off`).

    python tools/profile_step.py [B] [T] [speech_encoder] [warmup_steps]

Same model / step as bench.py (EEGConformerInterleaved depth 10 + speech tower + CLIPSimNoLatentProj, train mode);
prints the step's device time when run without a profiler (never a bench value when run under one).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import transformer_clip_eeg_b200 as pkg  # noqa: F401
from transformer_clip_eeg_b200 import train_clip_final as t, _lib
from transformer_clip_eeg_b200.optim import AdamW

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 320
speech = sys.argv[3] if len(sys.argv) > 3 else "convLSTM"
warm = int(sys.argv[4]) if len(sys.argv) > 4 else 3
torch.manual_seed(0)
args = t.build_parser().parse_args(["--speech_encoder", speech])
dev = torch.device("cuda")
model = t.build_model(args, T, 10000, dev)
opt = AdamW(model.parameters(), lr=1e-3, weight_decay=0.01)
eeg = torch.randn(B, T, 64, device=dev)
sp = torch.randn(B, T, 1024, device=dev)
ids = torch.arange(1, B + 1, device=dev)
model.train()
for _ in range(warm):
    t.train_step(model, opt, eeg, sp, ids)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record()
t.train_step(model, opt, eeg, sp, ids)
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"profiled step: B={B} T={T} speech={speech}: {e0.elapsed_time(e1):.2f} ms, eegclip launches so far {_lib.load().eegclip_launch_count()}")
