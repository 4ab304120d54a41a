"""Development aid: forward / backward time of the H = 128 bi-LSTM at the benchmarked shape, tensor-core vs FFMA2 recurrence."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import transformer_clip_eeg_b200 as pkg  # noqa: F401
from transformer_clip_eeg_b200 import clip_model as cm, _lib

B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 320
lstm = torch.nn.LSTM(64, 128, batch_first=True, bidirectional=True).to("cuda")
x = torch.randn(B, T, 64, device="cuda", requires_grad=True)
w = torch.randn(B, T, 256, device="cuda")


def run(n):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for _ in range(n):
        ev[0].record()
        y = cm._bilstm(lstm, x)
        ev[1].record()
        y.backward(w)
        ev[2].record()
        torch.cuda.synchronize()
        tf += ev[0].elapsed_time(ev[1]); tb += ev[1].elapsed_time(ev[2])
    return tf / n, tb / n, y.detach()


keys = (0, 1) if len(sys.argv) <= 2 else tuple(int(v) for v in sys.argv[2].split(","))
outs = {}
for key in keys:
    _lib.call("eegclip_tune_set", 14, key)
    run(3)
    import ctypes as C
    _lib.call("eegclip_profile_begin")
    tf, tb, y = run(10)
    ms = (C.c_double * 12)(); cnt = (C.c_longlong * 12)()
    _lib.call("eegclip_profile_end", ms, cnt, 12)
    print(f"  recurrence kernels: {ms[8] / 10 * 1e3:.0f} us per forward+backward pair ({cnt[8]} launches)")
    outs[key] = (y, x.grad.clone())
    x.grad = None
    print(f"tune14={key} ({'tensor-core' if key == 0 else 'FFMA2'} recurrence): forward {tf:.3f} ms, backward {tb:.3f} ms (incl. projections / weight gradients)")
if len(outs) == 2:
    print("max |y_mma - y_fp32| =", float((outs[0][0] - outs[1][0]).abs().max()), " rel dx =", float((outs[0][1] - outs[1][1]).norm() / outs[1][1].norm()))

# phase timeline of CTA (0,0) of the tensor-core forward recurrence (first steps)
_lib.call("eegclip_tune_set", 14, 0)
dbg = torch.zeros(768, dtype=torch.int64, device="cuda")
_lib.call("eegclip_debug_buffer", dbg.data_ptr())
with torch.no_grad():
    cm._bilstm(lstm, x)
torch.cuda.synchronize()
_lib.call("eegclip_debug_buffer", None)
d = dbg.cpu()
n = int(d[255]); t0 = int(d[1])
names = {0: "step start", 1: "x-proj read, copy-out issued", 2: "MMAs done", 3: "gates + stage stores done", 4: "cp.async landed", 5: "cp.async issued", 6: "x-proj in registers"}
prev = t0
for i in range(min(n, 60)):
    t = int(d[2 * i + 1])
    print(f"{(t - t0) / 1e3:8.2f} us (+{(t - prev) / 1e3:5.2f})  {names[int(d[2 * i])]}")
    prev = t
