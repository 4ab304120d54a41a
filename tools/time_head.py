"""Device timing of the symmetric InfoNCE head fwd+bwd at BASELINE config 3's size (B = 4096, D = 2560) on one GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import transformer_clip_eeg_b200 as pkg  # noqa: F401
from transformer_clip_eeg_b200.parallel import infonce_loss

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
D = int(sys.argv[2]) if len(sys.argv) > 2 else 2560
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
torch.manual_seed(0)
E = torch.randn(B, D, device="cuda", requires_grad=True)
S = torch.randn(B, D, device="cuda", requires_grad=True)
tau = torch.tensor(0.075, device="cuda", requires_grad=True)

def fb():
    E.grad = S.grad = tau.grad = None
    infonce_loss(E, S, tau).backward()

from transformer_clip_eeg_b200 import _lib
for wide in (0, 1):   # eegclip_tune_set(14, 1): always 256-column tiles (the 64-column small-batch form off)
    _lib.call("eegclip_tune_set", 14, wide)
    for _ in range(2):
        fb()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fb()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"head fwd+bwd B={B} D={D} {'256-wide tiles only' if wide else 'default tiles'}: {ms:.3f} ms -> {6.0 * B * B * D / ms / 1e9:.1f} TFLOP/s algorithmic")
_lib.call("eegclip_tune_set", 14, 0)
