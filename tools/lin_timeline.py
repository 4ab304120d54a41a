"""Development aid: role timeline of CTA 0 of the token-GEMM kernel (producer / MMA / epilogue) for one launch.

    python tools/lin_timeline.py [N] [K] [M]
"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import transformer_clip_eeg_b200 as pkg  # noqa: F401
from transformer_clip_eeg_b200 import _lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 192
K = int(sys.argv[2]) if len(sys.argv) > 2 else 64
M = int(sys.argv[3]) if len(sys.argv) > 3 else 81920
dev = "cuda"
x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * 0.1; b = torch.randn(N, device=dev)
out = torch.empty(M, N, device=dev)
nb = ctypes.c_size_t()
_lib.call("eegclip_linear_workspace", M, N, K, ctypes.byref(nb))
scratch = torch.empty(nb.value, dtype=torch.uint8, device=dev)
def run():
    _lib.call("eegclip_linear_forward", _lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(out), M, N, K, 1, _lib.ptr(scratch), _lib.stream())
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
print(f"N={N} K={K} M={M}: {e0.elapsed_time(e1) * 1e3:.1f} us (pack + gemm)")
dbg = torch.zeros(3 * 256, dtype=torch.int64, device=dev)
_lib.call("eegclip_debug_buffer", dbg.data_ptr())
run(); torch.cuda.synchronize()
_lib.call("eegclip_debug_buffer", None)
d = dbg.cpu().view(3, 256)
t0 = min(int(d[w_, 1]) for w_ in range(3) if int(d[w_, 255]) > 0)
names = {0: "prod:start", 1: "prod:slot-free", 2: "prod:filled", 10: "mma:weights", 11: "mma:acc-free", 12: "mma:stage-full", 20: "epi:start",
         21: "epi:acc-full", 22: "epi:done", 23: "epi:  tmem loaded", 24: "epi:  staged", 25: "epi:  stored"}
ev = []
for w_ in range(3):
    for i in range(int(d[w_, 255])):
        ev.append((int(d[w_, 2 * i + 1]) - t0, names[int(d[w_, 2 * i])]))
for t, n in sorted(ev):
    print(f"{t / 1e3:9.2f} us  {n}")
ref = x @ w.t() + b
print("max rel err", float((out - ref).abs().max() / ref.abs().max()))
