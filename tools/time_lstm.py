"""Development aid: time of the H = 128 bi-LSTM recurrences at the benchmarked shape + phase timeline of the forward kernel."""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import transformer_clip_eeg_b200 as pkg  # noqa: F401
from transformer_clip_eeg_b200 import clip_model as cm, _lib

B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 320
lstm = torch.nn.LSTM(64, 128, batch_first=True, bidirectional=True).to("cuda")
x = torch.randn(B, T, 64, device="cuda", requires_grad=True)
w = torch.randn(B, T, 256, device="cuda")


def prof(fn, n=10):
    fn()
    _lib.call("eegclip_profile_begin")
    for _ in range(n):
        fn()
    ms = (C.c_double * 12)(); cnt = (C.c_longlong * 12)()
    _lib.call("eegclip_profile_end", ms, cnt, 12)
    return ms[8] / n * 1e3


def fwd_only():
    with torch.no_grad():
        cm._bilstm(lstm, x)


both = prof(lambda: cm._bilstm(lstm, x).backward(w))
fwd = prof(fwd_only)
print(f"recurrence kernels at B = {B}: forward {fwd:.0f} us, backward {both - fwd:.0f} us")

# phase timeline of CTA (0,0) of the forward recurrence (first steps)
dbg = torch.zeros(768, dtype=torch.int64, device="cuda")
_lib.call("eegclip_debug_buffer", dbg.data_ptr())
fwd_only()
torch.cuda.synchronize()
_lib.call("eegclip_debug_buffer", None)
d = dbg.cpu()
n = int(d[255]); t0 = int(d[1])
names = {0: "step start (cluster barrier passed)", 2: "MMAs done", 3: "gates done, h written (local + peer)", 4: "global stores issued"}
prev = t0
for i in range(min(n, 24)):
    t = int(d[2 * i + 1])
    print(f"{(t - t0) / 1e3:8.2f} us (+{(t - prev) / 1e3:5.2f})  {names[int(d[2 * i])]}")
    prev = t
