"""Development aid: device time of one TransformerEncoderBlock forward+backward (B=256, T=320, train mode) per kernel
class, swept over the tuning knobs of eegclip_tune_set.  Not a bench value (bench.py is the contract).

    python tools/bench_xfblock.py [B] [T] ["k0=v0,k1=v1;k0=v0,..."]
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import transformer_clip_eeg_b200 as pkg  # noqa: F401
from transformer_clip_eeg_b200 import clip_model as cm, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 320
sweeps = sys.argv[3] if len(sys.argv) > 3 else "0=2,1=0;0=3,1=0;0=2,1=1;0=3,1=1"
names = ["conv", "conv_wgrad", "attn_fwd", "attn_bwd", "ln_ct", "gemm_f32", "lin_tc", "lin_wgrad", "lstm"]
torch.manual_seed(0)
dev = "cuda"
blk = cm.TransformerEncoderBlock(64).to(dev).train()
x = torch.randn(B, T, 64, device=dev)
w = torch.randn(B, T, 64, device=dev)
lib = _lib.load()
for sw in sweeps.split(";"):
    for kv in sw.split(","):
        if kv:
            k, v = kv.split("=")
            _lib.call("eegclip_tune_set", int(k), int(v))
    def step():
        xx = x.clone().requires_grad_(True)
        y = blk(xx)
        (y * w).sum().backward()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    n = 5
    _lib.call("eegclip_profile_begin")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = (ctypes.c_double * 12)()
    cnt = (ctypes.c_longlong * 12)()
    _lib.call("eegclip_profile_end", ctypes.cast(ms, ctypes.c_void_p), ctypes.cast(cnt, ctypes.c_void_p), 12)
    parts = ", ".join(f"{names[i]} {ms[i] / n:.3f} ms/{cnt[i] // n}" for i in range(len(names)) if cnt[i])
    print(f"[{sw}] block fwd+bwd {e0.elapsed_time(e1) / n:.3f} ms :: {parts}", flush=True)
