"""Development aid: phase timeline (stage / MMA / epilogue) of CTA 0 of the conv kernel, forward and backward."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import transformer_clip_eeg_b200 as pkg  # noqa: F401
from transformer_clip_eeg_b200 import clip_model as cm, _lib

T = int(sys.argv[1]) if len(sys.argv) > 1 else 320
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = "cuda"
blk = cm.BasicBlock(64, 64, kernel_size=64, time_dimension=T).to(dev).train()
x = torch.randn(B, T, 64, device=dev); skip = torch.randn(B, T, 64, device=dev)
for _ in range(3):
    blk.forward_time_major(x, skip)
modes = [int(m) for m in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
ss_list = [int(m) for m in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0]   # 1 = TS-form kernel (g_tune[6])
for ss, mode in [(q, m) for q in ss_list for m in modes]:     # mode: unused (the exp_mode experiments are recorded in conv_tc.cuh)
    _lib.call("eegclip_tune_set", 5, mode)
    pass  # (the TS-form conv experiment was removed in round 2; DESIGN.md keeps its measurement)
    dbg = torch.zeros(768, dtype=torch.int64, device=dev)
    blk.forward_time_major(x, skip)
    _lib.call("eegclip_debug_buffer", dbg.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); blk.forward_time_major(x, skip); e1.record(); torch.cuda.synchronize()
    _lib.call("eegclip_debug_buffer", None)
    d = dbg.cpu()
    t = [int(d[i]) - int(d[0]) for i in range(4)]
    print(f"{'TS' if ss else 'SS'} exp_mode {mode}: conv block fwd (pack + conv + LN) {e0.elapsed_time(e1) * 1e3:.1f} us; CTA0: staged {t[1] / 1e3:.2f} us, mma done {t[2] / 1e3:.2f} us, end {t[3] / 1e3:.2f} us")
_lib.call("eegclip_tune_set", 5, 0)
pass
