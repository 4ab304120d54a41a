"""GPU: the CUDA path (through the C ABI / boundary modules) against the golden fixtures and the live oracle."""
import os
import tempfile

import numpy as np
import pytest
import torch

from _util import check_digest, global_grad_norm, rel_err, synth
from oracle import eegclip_oracle as O

pytestmark = pytest.mark.gpu

# Tolerances (north star: loss and gradients <= 1e-3 relative, fp32 accumulate).
OUT_TOL = 1e-3
GRAD_TOL = 1e-3


@pytest.fixture(scope="module")
def pkg():
    import transformer_clip_eeg_b200 as p
    assert torch.cuda.is_available()
    from transformer_clip_eeg_b200 import _lib
    _lib.load()
    return p


@pytest.fixture(scope="module")
def cm(pkg):
    from transformer_clip_eeg_b200 import clip_model
    return clip_model


DEV = "cuda"


def _lstm_train(model):
    """cuDNN refuses RNN backward in eval mode; the LSTMs have no dropout, so train() changes nothing numerically."""
    for m in model.modules():
        if isinstance(m, torch.nn.LSTM):
            m.train()


# Absolute floor of SURVEY H3: a tensor whose reference gradient norm is below FLOOR_EPS x the global gradient norm
# (mathematically-zero gradients such as keys.bias, or the log-scale's gradient of an untrained model) is compared
# against that floor instead of its own norm -- it then contributes < FLOOR_EPS * tol to the global relative error.
FLOOR_EPS = 1e-3


def _grads_ok(model, gdig, tol, prefix=""):
    floor = FLOOR_EPS * global_grad_norm(gdig)
    for k, p in model.named_parameters():
        assert p.grad is not None, k
        check_digest(p.grad.cpu(), gdig[prefix + k], tol, k, floor=floor)


@pytest.mark.parametrize("name", ["tower_d2_T192_eval", "tower_d1_T320_eval", "tower_d2_T192_train"])
def test_tower_interleaved_golden(cm, golden, name, monkeypatch):
    g = golden[name]
    from transformer_clip_eeg_b200 import _lib
    model = cm.EEGConformerInterleaved(output_dim=8, time_dimension=g["T"], depth=g["depth"])
    model.load_state_dict(synth.make_state_dict(synth.interleaved_shapes(g["depth"], g["T"]), g["seed"]))
    model.to(DEV)
    if g["train"]:
        model.train()
        monkeypatch.setattr(_lib, "new_seed", lambda: g["drop_seed"])
    else:
        model.eval()
    x = synth.randn(g["seed"] + 1, g["B"], g["T"], 64).to(DEV).requires_grad_(True)
    w = synth.randn(g["seed"] + 2, g["B"], g["T"], 8).to(DEV)
    y = model(x)
    check_digest(y.cpu(), g["out"], OUT_TOL, "out")
    (y * w).sum().backward()
    check_digest(x.grad.cpu(), g["dx"], GRAD_TOL, "dx")
    _grads_ok(model, g["grads"], GRAD_TOL)


def test_tower_conformer_golden(cm, golden):
    g = golden["conformer_c2_d2_T192_eval"]
    model = cm.EEGConformer(output_dim=8, filters=(64,) * g["n_conv"], kernels=(64,) * g["n_conv"], time_dimension=g["T"],
                            depth=g["depth"])
    model.load_state_dict(synth.make_state_dict(synth.conformer_shapes(g["n_conv"], g["depth"], g["T"]), g["seed"]))
    model.to(DEV).eval()
    x = synth.randn(g["seed"] + 1, g["B"], g["T"], 64).to(DEV).requires_grad_(True)
    w = synth.randn(g["seed"] + 2, g["B"], g["T"], 8).to(DEV)
    y = model(x)
    check_digest(y.cpu(), g["out"], OUT_TOL, "out")
    (y * w).sum().backward()
    check_digest(x.grad.cpu(), g["dx"], GRAD_TOL, "dx")
    _grads_ok(model, g["grads"], GRAD_TOL)


def test_tower_vs_live_oracle_fp64(cm):
    """Full gradient vectors (not digests) against the oracle evaluated in fp64."""
    depth, T, B = 2, 192, 4
    sd = synth.make_state_dict(synth.interleaved_shapes(depth, T), 77)
    model = cm.EEGConformerInterleaved(output_dim=8, time_dimension=T, depth=depth)
    model.load_state_dict(sd)
    model.to(DEV).eval()
    x = synth.randn(78, B, T, 64)
    w = synth.randn(79, B, T, 8)
    xg = x.to(DEV).requires_grad_(True)
    y = model(xg)
    (y * w.to(DEV)).sum().backward()
    sdo = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    xo = x.double().requires_grad_(True)
    yo = O.eeg_conformer_interleaved(sdo, xo, depth)
    go = torch.autograd.grad((yo * w.double()).sum(), [xo] + list(sdo.values()), allow_unused=True)
    assert rel_err(y, yo) < OUT_TOL
    assert rel_err(xg.grad, go[0]) < GRAD_TOL
    total = sum(float(t.norm()) ** 2 for t in go[1:] if t is not None) ** 0.5
    named = dict(model.named_parameters())
    for k, gr in zip(sdo, go[1:]):
        gr = gr if gr is not None else torch.zeros_like(sdo[k])
        assert rel_err(named[k].grad, gr, floor=FLOOR_EPS * total) < GRAD_TOL, k
    # the whole gradient vector
    num = sum(float((named[k].grad.cpu().double() - (g if g is not None else 0 * sdo[k])).norm()) ** 2 for k, g in zip(sdo, go[1:])) ** 0.5
    assert num / total < GRAD_TOL


@pytest.mark.parametrize("name", ["head_B64_D2560", "head_B16_D1536", "head_B96_D200"])
def test_head_golden(pkg, golden, name):
    from transformer_clip_eeg_b200.parallel import infonce_loss
    g = golden[name]
    E = synth.randn(g["seed"], g["B"], g["D"])
    S = 0.5 * synth.randn(g["seed"] + 1, g["B"], g["D"]) + 0.5 * E
    Eg, Sg = E.to(DEV).requires_grad_(True), S.to(DEV).requires_grad_(True)
    tau = torch.tensor(g["tau"], device=DEV, requires_grad=True)
    loss = infonce_loss(Eg, Sg, tau)
    assert abs(float(loss) - g["loss"]) <= 1e-5 * max(1.0, abs(g["loss"]))
    loss.backward()
    check_digest(Eg.grad.cpu(), g["dE"], GRAD_TOL, "dE")
    check_digest(Sg.grad.cpu(), g["dS"], GRAD_TOL, "dS")
    assert abs(float(tau.grad) - g["dtau"]) <= GRAD_TOL * max(abs(g["dtau"]), 1e-3)


@pytest.mark.parametrize("name", ["clipsim_lam0", "clipsim_lam1"])
def test_clipsim_golden(cm, golden, name):
    g = golden[name]
    B, T, bank = g["B"], g["T"], g["bank"]
    mb = cm.memoryBank(bank_size=bank, device=torch.device(DEV), dim=T * 8)
    mb.memory.copy_(synth.randn(g["seed"] + 5, bank + 1, T * 8).abs().to(DEV))
    model = cm.CLIPSimNoLatentProj(torch.nn.Identity(), torch.nn.Identity(), mb, temperature=0.075, window_length=T,
                                   lambda_clip=1, lambda_average=g["lam_avg"]).to(DEV)
    model.eval()
    ef = synth.randn(g["seed"], B, T, 8)
    sf = 0.5 * synth.randn(g["seed"] + 1, B, T, 8) + 0.5 * ef
    efg, sfg = ef.to(DEV).requires_grad_(True), sf.to(DEV).requires_grad_(True)
    ids = torch.tensor(g["ids"], device=DEV)
    l_ce, l_avg, l_tot = model(efg, sfg, ids)
    for got, key in ((l_ce, "loss_ce"), (l_avg, "avg_loss"), (l_tot, "loss_total")):
        assert got.dim() == 0
        assert abs(float(got) - g[key]) <= 1e-5 * max(1.0, abs(g[key])), key
    l_tot.backward()
    check_digest(efg.grad.cpu(), g["d_eeg"], GRAD_TOL, "d_eeg")
    check_digest(sfg.grad.cpu(), g["d_speech"], GRAD_TOL, "d_speech")
    assert abs(float(model.temperature.grad) - g["dtau"]) <= GRAD_TOL * max(abs(g["dtau"]), 1e-3)
    assert model.temperature_eeg.grad is not None  # tensor(0.) when lambda_average == 0, never None
    assert abs(float(model.temperature_eeg.grad) - g["dtau_eeg"]) <= GRAD_TOL * max(abs(g["dtau_eeg"]), 1e-3)
    check_digest(mb.memory.cpu(), g["memory_after"], 1e-6, "memory")


@pytest.mark.parametrize("name", ["speech_smallConv", "speech_convLSTM"])
def test_speech_golden(cm, golden, name):
    g = golden[name]
    T = g["T"]
    if g["which"] == "smallConv":
        model = cm.SpeechSmallConv(output_dim=8, ks_temporal=16, dropout_rate=0.4, speech_dim=1024, time_dimension=T)
        model.load_state_dict(synth.make_state_dict(synth.small_conv_shapes(T), g["seed"]))
    else:
        model = cm.EEGConvLSTM(units_lstm=128, output_dim=8, dropout_rate=0.4, eeg_dim=1024, filters=(64,), kernels=(32,),
                               input_channels=1024, time_dimension=T)
        model.load_state_dict(synth.make_state_dict(synth.conv_lstm_shapes(T), g["seed"]))
    model.to(DEV).eval()
    _lstm_train(model)
    x = synth.randn(g["seed"] + 1, g["B"], T, 1024).to(DEV).requires_grad_(True)
    w = synth.randn(g["seed"] + 2, g["B"], T, 8).to(DEV)
    y = model(x)
    check_digest(y.cpu(), g["out"], OUT_TOL, "out")
    (y * w).sum().backward()
    check_digest(x.grad.cpu(), g["dx"], GRAD_TOL, "dx")
    _grads_ok(model, g["grads"], GRAD_TOL)


def test_full_model_golden(cm, golden):
    g = golden["full_d2_T192"]
    T, B, depth = g["T"], g["B"], g["depth"]
    eeg_m = cm.EEGConformerInterleaved(output_dim=8, time_dimension=T, depth=depth)
    eeg_m.load_state_dict(synth.make_state_dict(synth.interleaved_shapes(depth, T), g["seed"]))
    sp_m = cm.EEGConvLSTM(units_lstm=128, output_dim=8, dropout_rate=0.4, eeg_dim=1024, filters=(64,), kernels=(32,),
                          input_channels=1024, time_dimension=T)
    sp_m.load_state_dict(synth.make_state_dict(synth.conv_lstm_shapes(T), g["seed"] + 1))
    mb = cm.memoryBank(bank_size=g["bank"], device=torch.device(DEV), dim=T * 8)
    model = cm.CLIPSimNoLatentProj(eeg_m, sp_m, mb, temperature=0.075, window_length=T, lambda_clip=1, lambda_average=0.0).to(DEV)
    mb.memory.copy_(synth.randn(g["seed"] + 5, g["bank"] + 1, T * 8).abs().to(DEV))
    model.eval()
    _lstm_train(model)
    l_ce, l_avg, l_tot = model(synth.randn(g["seed"] + 10, B, T, 64).to(DEV), synth.randn(g["seed"] + 11, B, T, 1024).to(DEV),
                               torch.arange(1, B + 1, device=DEV))
    assert abs(float(l_ce) - g["loss_ce"]) <= 1e-5 * max(1.0, abs(g["loss_ce"]))
    assert abs(float(l_avg) - g["avg_loss"]) <= 1e-5 * max(1.0, abs(g["avg_loss"]))
    l_tot.backward()
    # state_dict schema identical to the reference's (checkpoint round-trip)
    assert set(dict(model.named_parameters())) == set(g["grads"])
    _grads_ok(model, g["grads"], GRAD_TOL)


def test_vlaai_golden(pkg, golden):
    from transformer_clip_eeg_b200 import vlaai
    g = golden["vlaai_B2"]
    model = vlaai.VLAAI()
    model.load_state_dict(synth.make_state_dict(synth.vlaai_shapes(320), g["seed"]))
    model.to(DEV).eval()
    x = synth.randn(g["seed"] + 1, g["B"], 320, 64).to(DEV).requires_grad_(True)
    w = synth.randn(g["seed"] + 2, g["B"], 64, 320).to(DEV)
    y = model(x)
    check_digest(y.cpu(), g["out"], OUT_TOL, "out")
    (y * w).sum().backward()
    # LeakyReLU kink: fp32-vs-fp64 runs of identical code differ by ~1 % in gradients (tests/test_oracle_golden.py)
    check_digest(x.grad.cpu(), g["dx"], 3e-2, "dx")
    _grads_ok(model, g["grads"], 3e-2)


def test_adamw_golden(pkg, golden):
    from transformer_clip_eeg_b200.optim import AdamW
    g = golden["adamw"]
    p = torch.nn.Parameter(synth.randn(g["seed"], 257).to(DEV))
    q = torch.nn.Parameter(torch.zeros(5, device=DEV))  # second, tiny tensor exercises the table
    opt = AdamW([p, q], lr=1e-3, betas=(0.9, 0.999), weight_decay=0.01)
    for s in range(g["steps"]):
        opt.zero_grad()
        p.grad = synth.randn(g["seed"] + 1 + s, 257).to(DEV)
        q.grad = torch.zeros(5, device=DEV)
        opt.step()
    assert float((p.detach().cpu().double() - torch.tensor(g["p"], dtype=torch.float64)).abs().max()) < 2e-6


def test_mm_decisions_golden(cm, golden):
    """Identical match-mismatch decisions on the fixed probe set (north star), via the reference's on-disk format."""
    from oracle.make_golden import write_mm_fixture
    from transformer_clip_eeg_b200 import train_clip_helper_functions as H
    g = golden["mm_T192"]
    T, seed = g["T"], g["seed"]
    eeg_m = cm.EEGConformerInterleaved(output_dim=8, time_dimension=T, depth=1)
    eeg_m.load_state_dict(synth.make_state_dict(synth.interleaved_shapes(1, T), seed))
    sp_m = cm.SpeechSmallConv(output_dim=8, ks_temporal=16, dropout_rate=0.4, speech_dim=1024, time_dimension=T)
    sp_m.load_state_dict(synth.make_state_dict(synth.small_conv_shapes(T), seed + 1))
    model = cm.CLIPSimNoLatentProj(eeg_m, sp_m, None, temperature=0.075, window_length=T, lambda_clip=1, lambda_average=0.0).to(DEV)
    with tempfile.TemporaryDirectory() as root:
        write_mm_fixture(root, T, n_sub=g["n_sub"], n_seg=g["n_seg"], n_stim=g["n_stim"], seed=seed + 20)
        ev, ev_logits, ev_top, ev_top_logits = H.evaluate_model_challenge_2023_mm(model, torch.device(DEV), speech_feature="wav2vec_19",
                                                                                 eeg_folder=root)
    assert set(ev) == set(g["evaluation"])
    for k in ev:
        assert abs(ev[k] - g["evaluation"][k]) < 1e-6, k              # accuracy == identical decisions
    for sub, per in g["logits"].items():
        for sid, (ref_scores, lab) in per.items():
            got_scores, got_lab = ev_logits[sub][sid]
            assert got_lab == lab
            assert int(np.argmax(got_scores)) == int(np.argmax(ref_scores)), (sub, sid)   # the decision itself
            assert np.allclose(got_scores, ref_scores, atol=2e-4)
    for sub in g["top_x"]:
        assert np.allclose(ev_top[sub], g["top_x"][sub], atol=1e-9), sub                  # top-x curves identical
        check_digest(torch.tensor(ev_top_logits[sub]["logits"]), g["bank_logits_digest"][sub], 1e-3, "bank logits")


@pytest.mark.parametrize("T,B", [(192, 3), (320, 5), (64, 2)])
@pytest.mark.parametrize("math", ["bf16x3", "bf16"])
def test_conv_tensor_core_vs_fp32(cm, pkg, T, B, math):
    """tcgen05 implicit-GEMM conv (forward, dgrad, wgrad) against the exact-fp32 CUDA-core path, same inputs."""
    from transformer_clip_eeg_b200 import _lib
    blk = cm.BasicBlock(64, 64, kernel_size=64, time_dimension=T).to(DEV).eval()
    with torch.no_grad():
        blk.normalization.weight.add_(0.1 * torch.randn_like(blk.normalization.weight))
    x = torch.randn(B, T, 64, device=DEV)
    skip = torch.randn(B, T, 64, device=DEV)
    w = torch.randn(B, T, 64, device=DEV)
    res = {}
    for m in ("fp32", math):
        _lib.set_default_math(m)
        try:
            xx = x.clone().requires_grad_(True)
            blk.zero_grad()
            y = blk.forward_time_major(xx, skip)
            (y * w).sum().backward()
            res[m] = (y.detach(), xx.grad.detach(), blk.conv.weight.grad.detach().clone(), blk.conv.bias.grad.detach().clone())
        finally:
            _lib.set_default_math("bf16x3")
    tol = 2e-4 if math == "bf16x3" else 3e-2
    for a, b, name in zip(res[math], res["fp32"], ("y", "dx", "dw", "db")):
        assert rel_err(a, b) < tol, (name, rel_err(a, b))


@pytest.mark.parametrize("Cin,Cout,taps,T,B", [(64, 256, 64, 320, 3), (256, 128, 64, 320, 2), (128, 128, 32, 192, 3), (256, 256, 64, 64, 5)])
def test_conv_tensor_core_channel_blocks_vs_fp32(cm, pkg, Cin, Cout, taps, T, B):
    """The VLAAI-shaped convs (vlaai.py:27-33: 64..256 channels): channel-blocked tcgen05 conv == exact-fp32 path."""
    from transformer_clip_eeg_b200 import _lib
    blk = cm.BasicBlock(Cin, Cout, kernel_size=taps, time_dimension=T, dropout_rate=0.0).to(DEV).eval()
    with torch.no_grad():
        blk.normalization.weight.add_(0.1 * torch.randn_like(blk.normalization.weight))
    x = torch.randn(B, T, Cin, device=DEV)
    skip = torch.randn(B, T, Cin, device=DEV)
    w = torch.randn(B, T, Cout, device=DEV)
    res = {}
    for m in ("fp32", "bf16x3"):
        _lib.set_default_math(m)
        try:
            xx = x.clone().requires_grad_(True)
            blk.zero_grad()
            y = blk.forward_time_major(xx, skip)
            (y * w).sum().backward()
            res[m] = (y.detach(), xx.grad.detach(), blk.conv.weight.grad.detach().clone(), blk.conv.bias.grad.detach().clone(),
                      blk.normalization.weight.grad.detach().clone())
        finally:
            _lib.set_default_math("bf16x3")
    for a, b, name in zip(res["bf16x3"], res["fp32"], ("y", "dx", "dw", "db", "dgamma")):
        assert rel_err(a, b) < 2e-4, (name, rel_err(a, b))


def test_full_size_properties(cm):
    """BASELINE config 2 sizes (B=256, T=320, depth 10): properties that need no CPU oracle run."""
    torch.manual_seed(0)
    depth, T, B = 10, 320, 256
    model = cm.EEGConformerInterleaved(output_dim=8, time_dimension=T, depth=depth).to(DEV).eval()
    x = torch.randn(B, T, 64, device=DEV)
    y = model(x)
    assert torch.isfinite(y).all()
    # per-sample independence: a window's embedding does not depend on its batch (no cross-sample op on the path)
    sub = [3, 77, 200, 255]
    y2 = model(x[sub].contiguous())
    assert rel_err(y[sub], y2) < 1e-5
    # untrained symmetric InfoNCE of independent embeddings is ~ ln(B)
    from transformer_clip_eeg_b200.parallel import infonce_loss
    tau = torch.tensor(0.075, device=DEV, requires_grad=True)
    E = y.flatten(1).detach().requires_grad_(True)
    S = torch.randn(B, T * 8, device=DEV, requires_grad=True)
    loss = infonce_loss(E, S, tau)
    assert abs(float(loss) - np.log(B)) < 0.2
    loss.backward()
    # gradient of the loss w.r.t. a raw embedding is orthogonal to it (the L2 normalisation's null direction)
    dots = (E.grad * E.detach()).sum(1).abs().max()
    assert float(dots) < 1e-4 * float(E.grad.norm() * E.detach().norm(dim=1).max())
    # keys.bias gradient is identically zero (softmax invariance)
    xg = x[:8].clone().requires_grad_(True)
    model(xg).square().sum().backward()
    named = dict(model.named_parameters())
    gnorm = sum(float(p.grad.norm()) ** 2 for p in named.values()) ** 0.5
    assert float(named["conformer_3.0.0.fn.1.keys.bias"].grad.norm()) < 1e-5 * gnorm
