"""CPU: the oracle restatement against the round-2 golden cases (oracle/make_golden_r2.py: BASELINE-size towers and head,
stand-alone sub-modules, memory bank with duplicate ids, non-default loss wrappers, regression head, Adam / amsgrad)."""
import numpy as np
import pytest
import torch

from _util import check_digest, global_grad_norm, synth
from oracle import eegclip_oracle as O
from oracle.make_golden_r2 import IdTower, fill_module

TOL = 2e-4


def _leaf(sd):
    return {k: v.clone().requires_grad_(True) for k, v in sd.items()}


def _check_grads(sd, grads, gdig, tol=TOL, prefix=""):
    floor = 1e-4 * global_grad_norm(gdig)
    for (k, _), gr in zip(sd.items(), grads):
        gr = gr if gr is not None else torch.zeros_like(sd[k])
        check_digest(gr, gdig[prefix + k], tol, k, floor=floor)


@pytest.mark.parametrize("name", ["tower_d10_T320_eval", "tower_d10_T320_train"])
def test_tower_baseline_size(golden, name):
    """depth 10, T = 320 (BASELINE configs 1-3): error growth over ten layers stays inside the restatement tolerance."""
    g = golden[name]
    sd = _leaf(synth.make_state_dict(synth.interleaved_shapes(g["depth"], g["T"]), g["seed"]))
    x = synth.randn(g["seed"] + 1, g["B"], g["T"], 64).requires_grad_(True)
    w = synth.randn(g["seed"] + 2, g["B"], g["T"], 8)
    y = O.eeg_conformer_interleaved(sd, x, g["depth"], O.Drop(g["train"], g["drop_seed"]))
    check_digest(y, g["out"], TOL, "out")
    grads = torch.autograd.grad((y * w).sum(), [x] + list(sd.values()), allow_unused=True)
    check_digest(grads[0], g["dx"], TOL, "dx")
    _check_grads(sd, grads[1:], g["grads"])


def test_head_config3_size(golden):
    g = golden["head_B4096_D2560"]
    E = synth.randn(g["seed"], g["B"], g["D"]).requires_grad_(True)
    S = (0.5 * synth.randn(g["seed"] + 1, g["B"], g["D"]) + 0.5 * E.detach()).requires_grad_(True)
    tau = torch.tensor(g["tau"], requires_grad=True)
    loss = O.symmetric_infonce(E, S, tau)
    assert abs(float(loss) - g["loss"]) < 1e-5 * max(1.0, abs(g["loss"]))
    dE, dS, dt = torch.autograd.grad(loss, [E, S, tau])
    check_digest(dE, g["dE"], TOL, "dE")
    check_digest(dS, g["dS"], TOL, "dS")
    assert abs(float(dt) - g["dtau"]) < 1e-3 * max(abs(g["dtau"]), 1e-3)


def _module_sd(module, seed):
    fill_module(module, seed)
    return _leaf({k: v.detach() for k, v in module.named_parameters()})


def test_submodules(golden):
    from transformer_clip_eeg_b200 import clip_model as cm     # parameter containers only (no kernel call on CPU)
    g = golden["submodules"]
    x0, w = synth.randn(g["seed"], g["B"], g["T"], 64), synth.randn(g["seed"] + 1, g["B"], g["T"], 64)
    mods = {"mha": (cm.MultiHeadAttention(64, 8, 0.5), lambda sd, x: O.mha(sd, "", x)),
            "ffn": (cm.FeedForwardBlock(64, expansion=4, drop_p=0.5), lambda sd, x: O.feed_forward(sd, "", x)),
            "residual": (cm.ResidualAdd(torch.nn.Sequential(cm.LayerNorm(64), cm.MultiHeadAttention(64, 8, 0.5), cm.Dropout(0.5))),
                         lambda sd, x: O.residual_ln_mha(sd, "", x)),
            "block": (cm.TransformerEncoderBlock(64), lambda sd, x: O.transformer_block(sd, "", x))}
    for i, (name, (m, fn)) in enumerate(mods.items()):
        sd = _module_sd(m, g["seed"] + 10 + i)
        x = x0.clone().requires_grad_(True)
        y = fn(sd, x)
        check_digest(y, g[name]["out"], TOL, name + ".out")
        grads = torch.autograd.grad((y * w).sum(), [x] + list(sd.values()), allow_unused=True)
        check_digest(grads[0], g[name]["dx"], TOL, name + ".dx")
        _check_grads(sd, grads[1:], g[name]["grads"])


def test_membank_duplicate_ids(golden):
    g = golden["membank_dup"]
    assert g["old_rows_equal_mem0"]                       # the reference returns the PRE-batch row for every duplicate
    mem = synth.randn(g["seed"], g["bank"] + 1, g["D"])
    ids = torch.tensor(g["ids"])
    data = synth.randn(g["seed"] + 1, len(g["ids"]), g["D"])
    mem0 = mem.clone()
    old = O.memory_bank_update(mem, ids, data)
    assert torch.equal(old, mem0[ids])
    check_digest(mem, g["memory_after"], 1e-6, "memory")
    # the LAST occurrence of a duplicated id is the update that lands (index_copy_ on CPU writes in order)
    last3 = max(i for i, v in enumerate(g["ids"]) if v == 3)
    assert np.allclose(mem[3].double().numpy(), (mem0[3] * 0.9 + data[last3] * (1 - 0.9)).double().numpy(), atol=1e-7)
    assert np.allclose(mem[3].double().numpy(), np.array(g["row3"]), atol=1e-7)


def test_loss_variants(golden):
    from transformer_clip_eeg_b200 import clip_model as cm
    g = golden["loss_variants"]
    B, T, seed = g["B"], g["T"], g["seed"]
    ids = torch.tensor(g["ids"])

    def inputs(s, n_rep=1):
        ef = synth.randn(s, n_rep * B, T, 8)
        sf = 0.5 * synth.randn(s + 1, B, T, 8) + 0.5 * ef[:B]
        return ef.requires_grad_(True), sf.requires_grad_(True)

    def check(name, out, names, ef, sf, sd):
        c = g[name]
        for n, v in zip(names, out):
            assert abs(float(v) - c[n]) <= 1e-5 * max(1.0, abs(c[n])), (name, n, float(v), c[n])
        total = out[names.index("loss_total")]
        grads = torch.autograd.grad(total, [ef, sf] + list(sd.values()), allow_unused=True)
        check_digest(grads[0], c["d_eeg"], TOL, name + ".d_eeg")
        check_digest(grads[1], c["d_speech"], TOL, name + ".d_speech")
        _check_grads(sd, grads[2:], c["grads"])

    std = ("loss_ce", "aux", "loss_total")
    kl = ("loss_total", "loss_ce", "log_pmu2", "kld_z2")
    # CLIPSim
    mem = synth.randn(seed + 5, 21, 16).abs()
    m = cm.CLIPSim(IdTower(), IdTower(), None, temperature=0.075, latent_dim=16, window_length=T, lambda_clip=1, lambda_average=0.5)
    sd = _module_sd(m, seed + 7)
    ef, sf = inputs(seed)
    out = O.clip_sim(ef, sf, ids, mem, sd["latent_projection_eeg.weight"], sd["latent_projection_speech.weight"], sd["temperature"], 1.0, 0.5)
    check("clip_sim", out, std, ef, sf, sd)
    check_digest(mem, g["clip_sim"]["memory_after"], 1e-6, "memory")
    for name, adapted in (("clip_mp", False), ("clip_mp_adapted", True)):
        m = cm.CLIPSimMultiplePositives(IdTower(), IdTower(), temperature=0.075, window_length=T, lambda_clip=1, lambda_average=0.5)
        sd = _module_sd(m, seed + 7)
        ef, sf = inputs(seed + 20, 3)
        check(name, O.clip_multiple_positives(ef, sf, sd["temperature"], 1.0, 0.5, adapted), std, ef, sf, sd)
    m = cm.CLIPKLDNoLatentProj(IdTower(), IdTower(), latent_dimension=T * 8, number_of_classes=20, latent_dimension2=64, temperature=0.075,
                               window_length=T, lambda_clip=1, lambda_lower_bound=0.5, lambda_discriminative=0.5)
    sd = _module_sd(m, seed + 7)
    ef, sf = inputs(seed + 40)
    check("clip_kld", O.clip_kld(ef, sf, ids, sd, sd["temperature"], 1.0, 0.5), kl, ef, sf, sd)
    m = cm.CLIPKLDWithLatentProj(IdTower(), IdTower(), latent_dimension=16, number_of_classes=20, temperature=0.075, window_length=T,
                                 lambda_clip=1, lambda_lower_bound=0.5, lambda_discriminative=0.5)
    sd = _module_sd(m, seed + 7)
    ef, sf = inputs(seed + 60)
    check("clip_kld_latent_proj", O.clip_kld_latent_proj(ef, sf, ids, sd, sd["temperature"], 1.0, 0.5), kl, ef, sf, sd)
    ef, sf = inputs(seed + 80)
    l = O.clip_no_contrastive(ef, sf)
    check("no_contrastive", (l, l, l), std, ef, sf, {})


def test_regression_step(golden):
    g = golden["regression"]["step"]
    seed = golden["regression"]["seed"]
    from transformer_clip_eeg_b200 import train_clip_helper_functions as H
    reg = H.RegressionModel(g["Cin"], output_dim=2)
    sd = _module_sd(reg, seed)
    x = synth.randn(seed + 1, g["B"], g["Cin"], g["T"]).requires_grad_(True)
    y = synth.randn(seed + 2, g["B"], 2, g["T"]) + 0.3 * x.detach()[:, :2]
    pred = O.regression_model(sd["conv.weight"], sd["conv.bias"], x)
    check_digest(pred, g["pred"], TOL, "pred")
    loss = O.pearson_loss(pred, y)
    assert np.allclose(loss.detach().double().numpy(), np.array(g["loss"]), atol=1e-6)
    assert abs(float(loss.mean()) - g["loss_mean"]) < 1e-6
    grads = torch.autograd.grad(loss.sum(), [x] + list(sd.values()))
    check_digest(grads[0], g["dx"], TOL, "dx")
    _check_grads(sd, grads[1:], g["grads"])


@pytest.mark.parametrize("name,kw", [("adam", dict(wd=0.0)), ("adam_wd", dict(wd=0.05)), ("adamw_amsgrad", dict(wd=0.01, decoupled=True, amsgrad=True)),
                                     ("adam_amsgrad", dict(amsgrad=True))])
def test_adam_variants(golden, name, kw):
    g = golden["optim"]
    p = synth.randn(g["seed"], 257)
    m, v, vmax = torch.zeros_like(p), torch.zeros_like(p), torch.zeros_like(p)
    for s in range(g["steps"]):
        gr = synth.randn(g["seed"] + 1 + s, 257) * (3.0 if s == 1 else 1.0)
        p, m, v, vmax = O.adam_step(p, gr, m, v, vmax, s + 1, **kw)
    assert float((p.double() - torch.tensor(g[name], dtype=torch.float64)).abs().max()) < 2e-6
