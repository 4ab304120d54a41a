"""Shared helpers for parity tests."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import synth  # noqa: E402


def check_digest(t, dig, rtol, name="", floor=0.0):
    """Compare a tensor with a golden digest (norm / sum / strided samples).

    ``floor`` is the absolute floor of SURVEY H3: tensors whose reference norm is
    below it (mathematically-zero gradients such as keys.bias) are compared in
    absolute terms against the floor instead of relatively.
    """
    d = synth.grad_digest(t)
    assert d["shape"] == dig["shape"], f"{name}: shape {d['shape']} != {dig['shape']}"
    scale = max(dig["norm"], floor, 1e-30)
    n = max(1, len(dig["samples"]))
    numel = 1
    for s in dig["shape"]:
        numel *= s
    assert abs(d["norm"] - dig["norm"]) <= rtol * scale, f"{name}: norm {d['norm']} vs {dig['norm']}"
    # per-sample tolerance: error of one element relative to the rms element size (x8 slack)
    rms = scale / (numel ** 0.5)
    for a, b in zip(d["samples"], dig["samples"]):
        assert abs(a - b) <= 8 * rtol * max(rms, abs(b)) + 1e-12, f"{name}: sample {a} vs {b} (rms {rms})"
    assert abs(d["sum"] - dig["sum"]) <= rtol * max(dig["asum"], floor) * 4 + 1e-12, f"{name}: sum {d['sum']} vs {dig['sum']}"


def rel_err(a, b, floor=0.0):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / max(float(b.norm()), floor, 1e-30))


def global_grad_norm(dig_map):
    return sum(v["norm"] ** 2 for v in dig_map.values()) ** 0.5


def poison_cuda_cache(mbytes=512):
    """Fill the caching allocator's free blocks with NaN, so that an output element a kernel fails to write cannot come out
    right by accident (torch.empty reusing the block that an earlier, correct run of the same shape has just freed)."""
    if not torch.cuda.is_available():
        return
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    big = torch.full((mbytes * 262144,), float("nan"), device="cuda")                      # large pool
    small = [torch.full((131072,), float("nan"), device="cuda") for _ in range(64)]        # small pool (< 1 MB requests)
    torch.cuda.synchronize()
    del big, small
