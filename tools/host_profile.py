"""Development aid: where the HOST time of a train step goes (cProfile over a few steps, no device sync inside)."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import transformer_clip_eeg_b200 as pkg  # noqa: F401
from transformer_clip_eeg_b200 import train_clip_final as t
from transformer_clip_eeg_b200.optim import AdamW

B, T = 256, 320
dev = torch.device("cuda")
model = t.build_model(t.build_parser().parse_args([]), T, 10000, dev)
opt = AdamW(model.parameters(), lr=1e-3, weight_decay=0.01)
model.train()
eeg, sp, ids = torch.randn(B, T, 64, device=dev), torch.randn(B, T, 1024, device=dev), torch.arange(1, B + 1, device=dev)
for _ in range(3):
    t.train_step(model, opt, eeg, sp, ids)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    t.train_step(model, opt, eeg, sp, ids)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
