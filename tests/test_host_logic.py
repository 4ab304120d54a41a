"""CPU: boundary / host logic that needs no GPU -- ABI surface, state_dict schema, sharded-InfoNCE plumbing."""
import os
import re
import subprocess
import sys

import pytest
import torch

from _util import ROOT, synth
from oracle import eegclip_oracle as O


def _header_functions():
    src = open(os.path.join(ROOT, "include", "eegclip.h")).read()
    return sorted(set(re.findall(r"EEGCLIP_API\s+[\w\s\*]+?\b(eegclip_\w+)\s*\(", src)))


def test_abi_library_exports_every_declared_symbol():
    import transformer_clip_eeg_b200 as pkg
    from transformer_clip_eeg_b200 import _lib
    pkg.build()
    lib = _lib.load()               # loads without a GPU; no compute call is made here
    declared = _header_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/eegclip.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == declared
    assert lib.eegclip_abi_version() == 4


def test_cpu_tensors_fail_loudly():
    """There is no CPU fallback: CPU inputs must raise, not silently run somewhere else."""
    from transformer_clip_eeg_b200 import clip_model as cm, _lib
    model = cm.EEGConformerInterleaved(depth=1, time_dimension=64)
    with pytest.raises(_lib.EegclipError):
        model(torch.zeros(2, 64, 64))
    from transformer_clip_eeg_b200.parallel import infonce_loss
    with pytest.raises(_lib.EegclipError):
        infonce_loss(torch.zeros(4, 8), torch.zeros(4, 8), torch.tensor(0.0))


def test_state_dict_schema_matches_reference(golden):
    """Key names and shapes equal the reference's (observed when the golden file was made)."""
    from transformer_clip_eeg_b200 import clip_model as cm, vlaai
    g = golden["full_d2_T192"]
    T = g["T"]
    eeg = cm.EEGConformerInterleaved(output_dim=8, time_dimension=T, depth=g["depth"])
    sp = cm.EEGConvLSTM(units_lstm=128, output_dim=8, dropout_rate=0.4, eeg_dim=1024, filters=(64,), kernels=(32,),
                        input_channels=1024, time_dimension=T)
    model = cm.CLIPSimNoLatentProj(eeg, sp, cm.memoryBank(g["bank"], torch.device("cpu"), T * 8), temperature=0.075, window_length=T)
    named = {k: list(v.shape) for k, v in model.named_parameters()}
    assert named == {k: v["shape"] for k, v in g["grads"].items()}
    assert "eegMemoryBank.memory" in model.state_dict() and model.state_dict()["eegMemoryBank.memory"].shape == (g["bank"] + 1, T * 8)
    assert set(eeg.state_dict()) == set(synth.interleaved_shapes(g["depth"], T))
    assert set(cm.EEGConformer(depth=2, time_dimension=T).state_dict()) == set(synth.conformer_shapes(2, 2, T))
    assert {k: tuple(v.shape) for k, v in vlaai.VLAAI().state_dict().items()} == synth.vlaai_shapes(320)
    assert set(cm.SpeechSmallConv(output_dim=8, ks_temporal=16, time_dimension=T).state_dict()) == set(synth.small_conv_shapes(T))
    assert eeg.get_output_dim(T) == T * 8


def test_abi_param_order_matches_header():
    from transformer_clip_eeg_b200 import clip_model as cm
    m = cm.EEGConformerInterleaved(depth=3, time_dimension=64)
    ps = m.abi_params()
    assert len(ps) == 2 + 4 * 3 + 16 * 3 + 2
    assert ps[0] is m.eeg_spatial_mapping.weight and ps[2] is m.conv_0.conv.weight and ps[5] is m.conv_0.normalization.bias
    blk = m.conformer_0[0]
    base = 2 + 4 * 3
    assert ps[base + 2] is blk[0].fn[1].queries.weight and ps[base + 4] is blk[0].fn[1].keys.weight
    assert ps[base + 12] is blk[1].fn[1][0].weight and ps[-2] is m.final_layer.weight


def test_cli_flags_match_reference_defaults():
    from transformer_clip_eeg_b200 import train_clip_final as t
    a = t.build_parser().parse_args([])
    assert (a.eeg_encoder, a.speech_encoder, a.model_arch) == ("EEGConformerInterleaved", "convLSTM", "clip_sim_no_latent_proj")
    assert (a.attention_depth, a.batch_size, a.latent_dim, a.temperature) == (10, 128, 8, 0.075)
    assert (a.learning_rate, a.weight_decay, a.beta1, a.beta2, a.lambda_sim_loss) == (1e-3, 0.01, 0.9, 0.999, 0.0)
    eeg = t.load_eeg_encoder("EEGConformerInterleaved", 128, "valid", 128, 1, 192, 8, 10)
    assert sum(p.numel() for p in eeg.parameters()) == 3372360   # == the reference class at depth 10, T=192 (instantiated)
    with pytest.raises(UnboundLocalError):
        t.load_eeg_encoder("transformerEncoder", 128, "valid", 128, 1, 192, 8, 10)


class OracleHeadOps:
    """Checker implementation of the head ops (CPU, from oracle/): lets the sharded plumbing run under gloo."""

    def l2norm_fwd(self, x):
        n = x.norm(dim=1).clamp_min(1e-12)
        return x / n[:, None], 1.0 / n

    def l2norm_bwd(self, xn, inv, dxn):
        return inv[:, None] * (dxn - xn * (xn * dxn).sum(1, keepdim=True))

    def lse(self, S_all, E_all, tau, b, row0, one_sided):
        sc = torch.exp(tau)
        rows = (S_all[row0:row0 + b] @ E_all.T) * sc
        cols = (S_all @ E_all[row0:row0 + b].T) * sc
        diag = (S_all[row0:row0 + b] * E_all[row0:row0 + b]).sum(1) * sc
        return torch.stack([torch.logsumexp(rows, 1), torch.logsumexp(cols, 0), diag])

    def loss(self, vec_all, one_sided):
        return ((vec_all[0] - vec_all[2]).mean() + (vec_all[1] - vec_all[2]).mean()) / 2

    def backward(self, S_all, E_all, tau, vec_all, b, row0, dloss, one_sided):
        Bg = E_all.shape[0]
        sc = torch.exp(tau)
        eye = torch.zeros(b, Bg, dtype=E_all.dtype)
        eye[torch.arange(b), torch.arange(row0, row0 + b)] = 1
        Lr = (S_all[row0:row0 + b] @ E_all.T) * sc
        Gr = (torch.exp(Lr - vec_all[0][row0:row0 + b, None]) + torch.exp(Lr - vec_all[1][None, :]) - 2 * eye) / (2 * Bg) * dloss
        Lc = (S_all @ E_all[row0:row0 + b].T) * sc
        Gc = (torch.exp(Lc - vec_all[0][:, None]) + torch.exp(Lc - vec_all[1][None, row0:row0 + b]) - 2 * eye.T) / (2 * Bg) * dloss
        return sc * Gr @ E_all, sc * Gc.T @ S_all, (Gr * Lr).sum()


def test_infonce_plumbing_world1_matches_oracle():
    from transformer_clip_eeg_b200.parallel import infonce_loss
    E = synth.randn(1, 12, 40).double().requires_grad_(True)
    S = synth.randn(2, 12, 40).double().requires_grad_(True)
    tau = torch.tensor(0.3, dtype=torch.float64, requires_grad=True)
    loss = infonce_loss(E, S, tau, ops=OracleHeadOps())
    ref = O.symmetric_infonce(E, S, tau)
    assert abs(float(loss - ref)) < 1e-12
    g1 = torch.autograd.grad(loss, [E, S, tau])
    g2 = torch.autograd.grad(ref, [E, S, tau])
    for a, b in zip(g1, g2):
        assert float((a - b).abs().max()) < 1e-10


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["EEGCLIP_ROOT"]); sys.path.insert(0, os.path.join(os.environ["EEGCLIP_ROOT"], "tests"))
from test_host_logic import OracleHeadOps
from _util import synth
from oracle import eegclip_oracle as O
from transformer_clip_eeg_b200.parallel import infonce_loss, allreduce_gradients
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
B, D = 8 * world, 24
E_full, S_full = synth.randn(1, B, D).double(), synth.randn(2, B, D).double()
W = synth.randn(3, D, D).double().requires_grad_(True)          # a shared "tower" weight
b = B // world
sl = slice(rank * b, (rank + 1) * b)
tau = torch.tensor(0.2, dtype=torch.float64, requires_grad=True)
E = (E_full[sl] @ W)
loss = infonce_loss(E, S_full[sl].clone(), tau, group=dist.group.WORLD, ops=OracleHeadOps())
loss.backward()
allreduce_gradients([W, tau], dist.group.WORLD)                 # SUM, not MEAN (SURVEY 8(e))
Wr = W.detach().clone().requires_grad_(True); tr = tau.detach().clone().requires_grad_(True)
ref = O.symmetric_infonce(E_full @ Wr, S_full, tr)
gW, gt = torch.autograd.grad(ref, [Wr, tr])
assert abs(float(loss - ref)) < 1e-12, (float(loss), float(ref))
assert float((W.grad - gW).abs().max()) < 1e-10
assert abs(float(tau.grad - gt)) < 1e-10
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_sharded_infonce_two_ranks_gloo(tmp_path):
    """world_size-2 gloo run: R ranks x b rows == single process at batch R*b (loss and SUM-reduced grads)."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, EEGCLIP_ROOT=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("ok") == 2


_TOPK_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["EEGCLIP_ROOT"]); sys.path.insert(0, os.path.join(os.environ["EEGCLIP_ROOT"], "tests"))
from _util import synth
from transformer_clip_eeg_b200.train_clip_helper_functions import bank_topk
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
N, D = 9, 16
for M, k in [(37, 5), (11, 8), (3, 100)]:               # ragged slices; k above a slice; k above the whole bank
    E, Bk = synth.randn(1, N, D).double(), synth.randn(2, M, D).double()
    bounds = [0, M // 3, M] if world == 2 else [0, M]
    loc = Bk[bounds[rank]:bounds[rank + 1]]
    # checker ops injected (the product ops are CUDA-only); chunk=4 exercises the chunked running-top-k merge as well
    v, i = bank_topk(E, loc, k, group=dist.group.WORLD, logits_fn=lambda e, b: e @ b.T,
                     topk_fn=lambda x, kk: tuple(torch.topk(x, kk, dim=1)), chunk=4)
    rv, ri = torch.topk(E @ Bk.T, min(k, M), dim=1)
    assert torch.equal(i, ri), (M, k, i, ri)
    assert torch.allclose(v, rv, rtol=1e-12, atol=1e-12)
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_sharded_bank_topk_two_ranks_gloo(tmp_path):
    """BASELINE config 4 host logic: candidate bank sharded over 2 ranks, local top-k + all-gather + merge == global top-k."""
    script = tmp_path / "worker_topk.py"
    script.write_text(_TOPK_WORKER)
    env = dict(os.environ, EEGCLIP_ROOT=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29633", str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("ok") == 2
