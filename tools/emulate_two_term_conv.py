"""Round-2 experiment: would a 2-term tensor-core product be accurate enough for the conv?

Emulates, on the CPU oracle in fp64, a conv whose activation-side operand (input in the forward, output gradient in the data
gradient, input in the weight gradient) is rounded ONCE to fp16 (11-bit mantissa, RN, per-tensor power-of-two scale) while the
weight-side operand stays exact -- i.e. a_hi * [w_hi | w_lo], two MMAs per MAC instead of three -- through the depth-10, T = 320
EEG tower, and prints output / gradient errors against the exact evaluation.  Result: outputs 2.0e-4, input gradient 3.4e-4, whole
parameter-gradient vector 3.6e-4, worst tensor (conv_0.conv.weight) 9.1e-4: no margin under the 1e-3 per-tensor gate.

    python tools/emulate_two_term_conv.py [fp16a|fp16b|exact]     (oracle/ is test infrastructure; this is a development aid)
"""
import sys, torch, math
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import torch.nn.functional as F
from oracle import eegclip_oracle as O, synth
torch.set_num_threads(16)

def r16(x, scale_pow2=True):
    # round to fp16 after per-tensor power-of-two scaling (amax -> ~2^14)
    amax = float(x.abs().max())
    if amax == 0: return x
    k = math.floor(math.log2(16384.0/amax))
    s = 2.0**k
    return (x*s).half().to(x.dtype)/s

MODE = sys.argv[1] if len(sys.argv)>1 else 'fp16a'

class ConvEmu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        # x (B,C,T) fp64; emulate: A=x rounded fp16 single, W exact (hi+lo)
        ctx.save_for_backward(x, w)
        xr = r16(x) if MODE!='exact' else x
        return O.conv1d_same(xr, w, b)
    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dyr = r16(dy) if MODE!='exact' else dy
        xr = r16(x) if MODE!='exact' else x
        k = w.shape[-1]; left=(k-1)//2
        with torch.enable_grad():
            xx = x.detach().requires_grad_(True); ww = w.detach().requires_grad_(True)
            y = O.conv1d_same(xx, ww, None)
            # dgrad: uses rounded dy, exact w
            dx, = torch.autograd.grad(y, xx, dyr, retain_graph=True)
            # wgrad: x rounded single (fp16), dy exact (split)  [variant]
            xx2 = xr.detach().requires_grad_(True)
            y2 = O.conv1d_same(xx2, ww, None)
            dw, = torch.autograd.grad(y2, ww, dy if MODE=='fp16a' else dyr)
        db = dy.sum((0,2))
        return dx, dw, db

orig = O.conv1d_same
def patched_basic_block(sd, pre, x, drop=O.EVAL, layer=0, p=0.2, act="gelu"):
    y = ConvEmu.apply(x, sd[pre + "conv.weight"], sd[pre + "conv.bias"])
    y = drop(y, p, layer, O.SITE_CONV, order=(0, 2, 1))
    g, b = sd[pre + "normalization.weight"], sd[pre + "normalization.bias"]
    y = F.layer_norm(y, tuple(g.shape), g, b, 1e-5)
    return F.gelu(y)

depth, T, B = 10, 320, 4
sd0 = synth.make_state_dict(synth.interleaved_shapes(depth, T), 900)
x = synth.randn(901, B, T, 64).double(); w = synth.randn(902, B, T, 8).double()
def run(patch):
    sd = {k: v.double().requires_grad_(True) for k, v in sd0.items()}
    xo = x.clone().requires_grad_(True)
    bb = O.basic_block
    if patch: O.basic_block = patched_basic_block
    try:
        y = O.eeg_conformer_interleaved(sd, xo, depth)
        gs = torch.autograd.grad((y*w).sum(), [xo]+list(sd.values()), allow_unused=True)
    finally:
        O.basic_block = bb
    return y.detach(), gs, list(sd.keys())
y0, g0, keys = run(False)
y1, g1, _ = run(True)
print('out rel', float((y1-y0).norm()/y0.norm()))
print('dx rel', float((g1[0]-g0[0]).norm()/g0[0].norm()))
tot = sum(float(g.norm())**2 for g in g0[1:] if g is not None)**0.5
num = sum(float((a-b).norm())**2 for a,b in zip(g1[1:],g0[1:]) if a is not None)**0.5
print('param grad vector rel', num/tot)
worst = max(((float((a-b).norm()/max(float(b.norm()),1e-3*tot)), k) for a,b,k in zip(g1[1:],g0[1:],keys) if a is not None))
print('worst tensor', worst)
