"""A/B device timing of the train step with the speech tower on a side stream (EEGCLIP_TWO_STREAMS=1, default) against both towers
on one stream (=0), alternating on the same box (development aid; bench.py is the contract)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import transformer_clip_eeg_b200 as pkg  # noqa: F401
from transformer_clip_eeg_b200 import train_clip_final as t
from transformer_clip_eeg_b200.optim import AdamW

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 320
N = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = torch.device("cuda")
args = t.build_parser().parse_args([])
model = t.build_model(args, T, 10000, dev).train()
opt = AdamW(model.parameters(), lr=1e-3, weight_decay=0.01)
batches = [(torch.randn(B, T, 64, device=dev), torch.randn(B, T, 1024, device=dev), torch.arange(1, B + 1, device=dev)) for _ in range(2)]


def run(n):
    for i in range(n):
        t.train_step(model, opt, *batches[i % 2])


for rnd in range(3):
    for mode in ("0", "1"):
        os.environ["EEGCLIP_TWO_STREAMS"] = mode
        run(3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(N)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / N
        print(f"round {rnd} two_streams={mode}: {ms:.3f} ms/step -> {B / ms * 1e3:.0f} samples/s", flush=True)
