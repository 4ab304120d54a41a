"""Development aid: role timeline of CTA 0 of the token weight-gradient kernel (producer warp 3 / MMA warp 0 / epilogue).

    python tools/wgrad_timeline.py [N] [K] [M]
"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import transformer_clip_eeg_b200 as pkg  # noqa: F401
from transformer_clip_eeg_b200 import _lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
K = int(sys.argv[2]) if len(sys.argv) > 2 else 64
M = int(sys.argv[3]) if len(sys.argv) > 3 else 81920
dev = "cuda"
x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * 0.1; dy = torch.randn(M, N, device=dev)
dw = torch.empty(N, K, device=dev); db = torch.empty(N, device=dev)
nb = ctypes.c_size_t()
_lib.call("eegclip_linear_workspace", M, N, K, ctypes.byref(nb))
scratch = torch.empty(nb.value, dtype=torch.uint8, device=dev)
def run():
    _lib.call("eegclip_linear_backward", _lib.ptr(x), _lib.ptr(w), _lib.ptr(dy), None, _lib.ptr(dw), _lib.ptr(db), M, N, K, 1,
              _lib.ptr(scratch), _lib.stream())
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
print(f"N={N} K={K} M={M}: {e0.elapsed_time(e1) * 1e3:.1f} us (wgrad + reduce)")
dbg = torch.zeros(3 * 256, dtype=torch.int64, device=dev)
_lib.call("eegclip_debug_buffer", dbg.data_ptr())
run(); torch.cuda.synchronize()
_lib.call("eegclip_debug_buffer", None)
d = dbg.cpu().view(3, 256)
t0 = min(int(d[w_, 1]) for w_ in range(3) if int(d[w_, 255]) > 0)
names = {0: "prod:start", 1: "prod:  data landed", 2: "prod:  operand slot free", 3: "prod:  converted+arrived", 4: "prod:done",
         12: "mma:stage full", 13: "mma:issued", 20: "epi:start", 21: "epi:acc full", 30: "tma:slot free, copies issued"}
ev = []
for w_ in range(3):
    for i in range(int(d[w_, 255])):
        ev.append((int(d[w_, 2 * i + 1]) - t0, names.get(int(d[w_, 2 * i]), str(int(d[w_, 2 * i])))))
for t, n in sorted(ev):
    print(f"{t / 1e3:9.2f} us  {n}")
ref = dy.t() @ x
print("max rel err", float((dw - ref).abs().max() / ref.abs().max()))
