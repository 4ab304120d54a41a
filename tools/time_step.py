"""Quick device timing of one train step (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import transformer_clip_eeg_b200 as pkg
from transformer_clip_eeg_b200 import train_clip_final as t, _lib
from transformer_clip_eeg_b200.optim import AdamW

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 320
speech = sys.argv[3] if len(sys.argv) > 3 else "convLSTM"
args = t.build_parser().parse_args(["--speech_encoder", speech])
dev = torch.device("cuda")
model = t.build_model(args, T, 10000, dev)
opt = AdamW(model.parameters(), lr=1e-3, weight_decay=0.01)
eeg = torch.randn(B, T, 64, device=dev); sp = torch.randn(B, T, 1024, device=dev); ids = torch.arange(1, B + 1, device=dev)
model.train()
for mode in ("train", "eval"):
    model.train(mode == "train")
    for m in model.modules():
        if isinstance(m, torch.nn.LSTM):
            m.train()   # cuDNN refuses RNN backward in eval mode
    for i in range(2):
        t.train_step(model, opt, eeg, sp, ids)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 3
    for i in range(n):
        t.train_step(model, opt, eeg, sp, ids)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{mode}: B={B} T={T} speech={speech}: {ms:.2f} ms/step -> {B / ms * 1e3:.0f} samples/s")
# tower only
x = eeg.clone().requires_grad_(True)
model.eval()
for i in range(2):
    model.eegModel(x).sum().backward()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(3):
    model.eegModel(x).sum().backward()
e1.record(); torch.cuda.synchronize()
print(f"eeg tower fwd+bwd eval: {e0.elapsed_time(e1) / 3:.2f} ms")
