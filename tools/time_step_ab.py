"""A/B device timing of the train step under one eegclip_tune_set knob, alternating on the same box (development aid).

    python tools/time_step_ab.py <knob index> [value_a] [value_b] [B] [T] [steps]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import transformer_clip_eeg_b200 as pkg  # noqa: F401
from transformer_clip_eeg_b200 import train_clip_final as t, _lib
from transformer_clip_eeg_b200.optim import AdamW

knob = int(sys.argv[1])
va = int(sys.argv[2]) if len(sys.argv) > 2 else 0
vb = int(sys.argv[3]) if len(sys.argv) > 3 else 1
B = int(sys.argv[4]) if len(sys.argv) > 4 else 256
T = int(sys.argv[5]) if len(sys.argv) > 5 else 320
N = int(sys.argv[6]) if len(sys.argv) > 6 else 10
dev = torch.device("cuda")
model = t.build_model(t.build_parser().parse_args([]), T, 10000, dev).train()
opt = AdamW(model.parameters(), lr=1e-3, weight_decay=0.01)
batches = [(torch.randn(B, T, 64, device=dev), torch.randn(B, T, 1024, device=dev), torch.arange(1, B + 1, device=dev)) for _ in range(2)]
for kv in filter(None, os.environ.get("FIX", "").split(",")):     # FIX="15=1,7=0": knobs held fixed during the A/B
    k_, v_ = kv.split("=")
    _lib.call("eegclip_tune_set", int(k_), int(v_))
for rnd in range(3):
    for v in (va, vb):
        _lib.call("eegclip_tune_set", knob, v)
        for i in range(3):
            t.train_step(model, opt, *batches[i % 2])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(N):
            t.train_step(model, opt, *batches[i % 2])
        e1.record()
        torch.cuda.synchronize()
        print(f"round {rnd} tune[{knob}]={v}: {e0.elapsed_time(e1) / N:.3f} ms/step", flush=True)
_lib.call("eegclip_tune_set", knob, 0)
