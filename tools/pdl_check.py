import sys, os
sys.path.insert(0, "/root/repo")
import torch
import transformer_clip_eeg_b200 as pkg
from transformer_clip_eeg_b200 import clip_model as cm, _lib
_lib.load()
torch.manual_seed(11)
model = cm.EEGConformerInterleaved(output_dim=8, time_dimension=192, depth=2).to("cuda").eval()
x = torch.randn(4, 192, 64, device="cuda")
def run(off):
    _lib.call("eegclip_tune_set", 7, off)
    xx = x.clone().requires_grad_(True)
    model.zero_grad()
    y = model(xx)
    y.square().sum().backward()
    torch.cuda.synchronize()
    return [y.detach().clone(), xx.grad.detach().clone()] + [p.grad.detach().clone() for p in model.parameters()]
names = ["y", "dx"] + [n for n, _ in model.named_parameters()]
runs = [(o, run(o)) for o in (1, 0, 1, 0, 0, 1)]
base = runs[0][1]
for i, (o, R) in enumerate(runs):
    dy = float((R[0] - base[0]).abs().max())
    print("run", i, "pdl", "off" if o else "on", "max|y - y0| =", dy, " y norm", float(base[0].abs().max()))
# forward only, no backward in between
_lib.call("eegclip_tune_set", 7, 1); y1 = model(x).detach().clone()
_lib.call("eegclip_tune_set", 7, 0); y2 = model(x).detach().clone(); y3 = model(x).detach().clone()
_lib.call("eegclip_tune_set", 7, 1); y4 = model(x).detach().clone()
torch.cuda.synchronize()
print("fwd only: off-on", float((y1 - y2).abs().max()), "on-on", float((y2 - y3).abs().max()), "off-off", float((y1 - y4).abs().max()))
# per-layer: find the first differing intermediate through the conv block alone
blk = model.conv_0
_lib.call("eegclip_tune_set", 7, 1); c1 = blk.forward_time_major(x, x).detach().clone()
_lib.call("eegclip_tune_set", 7, 0); c2 = blk.forward_time_major(x, x).detach().clone()
print("conv block alone off-on", float((c1 - c2).abs().max()))
_lib.call("eegclip_tune_set", 7, 0)
