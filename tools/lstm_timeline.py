"""Development aid: per-step phase timeline of CTA (0,0) of the H = 128 LSTM forward recurrence."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import transformer_clip_eeg_b200 as pkg  # noqa: F401
from transformer_clip_eeg_b200 import clip_model as cm, _lib

B, T = 256, 320
lstm = torch.nn.LSTM(64, 128, batch_first=True, bidirectional=True).to("cuda")
x = torch.randn(B, T, 64, device="cuda")
for _ in range(2):
    cm._bilstm(lstm, x)
dbg = torch.zeros(768, dtype=torch.int64, device="cuda")
_lib.call("eegclip_debug_buffer", dbg.data_ptr())
cm._bilstm(lstm, x); torch.cuda.synchronize()
_lib.call("eegclip_debug_buffer", None)
d = dbg.cpu()
n = int(d[255]); t0 = int(d[1])
names = {0: "step start", 1: "product done", 2: "sync1 passed", 3: "gates+stores issued"}
prev = t0
for i in range(n):
    t = int(d[2 * i + 1])
    print(f"{(t - t0) / 1e3:8.2f} us (+{(t - prev) / 1e3:5.2f})  {names[int(d[2 * i])]}")
    prev = t
