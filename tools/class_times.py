"""Per-kernel-class CUDA-event times of the train step (events around every launch, single stream) under two values of one
eegclip_tune_set knob, alternating on the same box (development aid).   python tools/class_times.py <knob> [va] [vb]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["EEGCLIP_TWO_STREAMS"] = "0"
import torch
import transformer_clip_eeg_b200 as pkg  # noqa: F401
from transformer_clip_eeg_b200 import train_clip_final as t, _lib
from transformer_clip_eeg_b200.optim import AdamW

knob = int(sys.argv[1]); va = int(sys.argv[2]) if len(sys.argv) > 2 else 0; vb = int(sys.argv[3]) if len(sys.argv) > 3 else 1
B, T, N = 256, 320, 5
dev = torch.device("cuda")
model = t.build_model(t.build_parser().parse_args([]), T, 10000, dev).train()
opt = AdamW(model.parameters(), lr=1e-3, weight_decay=0.01)
batches = [(torch.randn(B, T, 64, device=dev), torch.randn(B, T, 1024, device=dev), torch.arange(1, B + 1, device=dev)) for _ in range(2)]
names = ["conv_fwd_dgrad_tc", "conv_wgrad_tc", "attn_fwd", "attn_bwd", "ln_ct", "head_similarity_tc", "lin_tc", "lin_wgrad_tc", "lstm_recurrence"]
_lib.call("eegclip_tune_set", 8, 0)
for rnd in range(2):
    for v in (va, vb):
        _lib.call("eegclip_tune_set", knob, v)
        for i in range(3):
            t.train_step(model, opt, *batches[i % 2])
        torch.cuda.synchronize()
        _lib.call("eegclip_profile_begin")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(N):
            t.train_step(model, opt, *batches[i % 2])
        e1.record()
        torch.cuda.synchronize()
        ms = (ctypes.c_double * 12)(); n = (ctypes.c_longlong * 12)()
        _lib.call("eegclip_profile_end", ctypes.cast(ms, ctypes.c_void_p), ctypes.cast(n, ctypes.c_void_p), 12)
        print(f"round {rnd} tune[{knob}]={v}: step {e0.elapsed_time(e1) / N:.3f} ms; " +
              ", ".join(f"{nm} {ms[i] / N:.3f}" for i, nm in enumerate(names)), flush=True)
_lib.call("eegclip_tune_set", knob, 0)
