"""TEST INFRASTRUCTURE ONLY -- round-2 golden cases from the UNMODIFIED reference (tests/golden/reference_golden_r2.json).

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden_r2

Adds, next to oracle/make_golden.py's cases:
  * the BASELINE-config tower and full model: depth 10, T = 320, B = 8, eval and train mode (Philox-patched dropouts);
  * the contrastive head at BASELINE config 3's size: B = 4096, D = 2560;
  * the stand-alone transformer sub-modules (MultiHeadAttention, FeedForwardBlock, ResidualAdd);
  * memoryBank with duplicate ids in one batch;
  * the non-default loss wrappers (CLIPSim, CLIPSimMultiplePositives[Adapted], CLIPKLDNoLatentProj,
    CLIPKLDWithLatentProj, CLIPNoContrastiveLearning);
  * the regression evaluation: RegressionModel + PearsonLoss single step, and the reference's own
    evaluate_model_do_regression_sub_specific run end to end on a synthetic EEGDatasetSimdata stand-in;
  * Adam / amsgrad optimizer steps.
"""
import json
import os
import tempfile
import zlib

import numpy as np
import torch
import torch.nn as nn

from . import eegclip_oracle as O
from . import synth
from .make_golden import (OUT, REF, _import_reference, build_full_model, digest_grads, patch_dropouts_interleaved)


def fill_module(model, seed, scale=None):
    """Deterministic weights for modules whose parameters the reference initialises from torch's RNG: parameters in sorted-name
    order, uniform(-b, b) with b = 1/sqrt(fan_in) (matrices) or 0.1 (vectors); 0-dim parameters (temperatures) are kept."""
    rs = np.random.RandomState(seed)
    named = dict(model.named_parameters())
    with torch.no_grad():
        for k in sorted(named):
            p = named[k]
            if p.dim() == 0:
                continue
            b = scale if scale is not None else (0.1 if p.dim() == 1 else 1.0 / np.sqrt(int(np.prod(p.shape[1:]))))
            p.copy_(torch.from_numpy(rs.uniform(-b, b, size=tuple(p.shape)).astype(np.float32)))
    return model


class IdTower(nn.Module):
    """Identity 'tower' with the get_output_dim the loss wrappers ask for."""

    def __init__(self, out_dim=8):
        super().__init__()
        self.out_dim = out_dim

    def get_output_dim(self, input_window_size):
        return input_window_size * self.out_dim

    def forward(self, x):
        return x


# ---- BASELINE-size towers -------------------------------------------------------------------------------------------------
def case_tower_big(cm, depth, T, B, seed, train):
    model = cm.EEGConformerInterleaved(output_dim=8, time_dimension=T, depth=depth)
    model.load_state_dict(synth.make_state_dict(synth.interleaved_shapes(depth, T), seed))
    x = synth.randn(seed + 1, B, T, 64).requires_grad_(True)
    w = synth.randn(seed + 2, B, T, 8)
    if train:
        model.train()
        patch_dropouts_interleaved(model, depth, O.Drop(True, seed=seed + 3))
    else:
        model.eval()
    y = model(x)
    (y * w).sum().backward()
    return {"kind": "interleaved", "depth": depth, "T": T, "B": B, "seed": seed, "train": train, "drop_seed": seed + 3,
            "out": synth.grad_digest(y), "dx": synth.grad_digest(x.grad), "grads": digest_grads(model)}


def patch_dropouts_convlstm(model, drop):
    for i in range(model.n_blocks):
        mod = getattr(model, f"conv_{i}").dropout
        mod.forward = lambda x, _p=mod.p, _l=i: drop(x, _p, _l, O.SITE_CONV, order=(0, 2, 1))


def case_full_big(cm, depth, T, B, bank, seed, train):
    model = build_full_model(cm, depth, T, bank, seed)
    if train:
        model.train()
        drop = O.Drop(True, seed=seed + 3)
        patch_dropouts_interleaved(model.eegModel, depth, drop)
        patch_dropouts_convlstm(model.speechModel, drop)
    else:
        model.eval()
    eeg, sp = synth.randn(seed + 10, B, T, 64), synth.randn(seed + 11, B, T, 1024)
    ids = torch.arange(1, B + 1, dtype=torch.int64)
    l_ce, l_avg, l_tot = model(eeg, sp, ids)
    l_tot.backward()
    return {"kind": "full", "depth": depth, "T": T, "B": B, "bank": bank, "seed": seed, "train": train, "drop_seed": seed + 3,
            "loss_ce": float(l_ce), "avg_loss": float(l_avg), "loss_total": float(l_tot), "grads": digest_grads(model)}


def case_head_big(cm, B, D, tau, seed):
    model = cm.CLIP(nn.Identity(), nn.Identity(), temperature=tau)
    E = synth.randn(seed, B, D).requires_grad_(True)
    S = (0.5 * synth.randn(seed + 1, B, D) + 0.5 * E.detach()).requires_grad_(True)
    loss = model(E, S)
    loss.backward()
    return {"kind": "head", "B": B, "D": D, "tau": tau, "seed": seed, "loss": float(loss),
            "dE": synth.grad_digest(E.grad), "dS": synth.grad_digest(S.grad), "dtau": float(model.temperature.grad)}


# ---- stand-alone transformer sub-modules -----------------------------------------------------------------------------------
def case_submodules(cm, B, T, seed):
    out = {"kind": "submodules", "B": B, "T": T, "seed": seed}
    x0 = synth.randn(seed, B, T, 64)
    w = synth.randn(seed + 1, B, T, 64)
    mods = {
        "mha": cm.MultiHeadAttention(64, 8, 0.5),
        "ffn": cm.FeedForwardBlock(64, expansion=4, drop_p=0.5),
        "residual": cm.ResidualAdd(nn.Sequential(nn.LayerNorm(64), cm.MultiHeadAttention(64, 8, 0.5), nn.Dropout(0.5))),
        "block": cm.TransformerEncoderBlock(64),
    }
    for i, (name, m) in enumerate(mods.items()):
        fill_module(m, seed + 10 + i)
        m.eval()
        x = x0.clone().requires_grad_(True)
        y = m(x)
        (y * w).sum().backward()
        out[name] = {"out": synth.grad_digest(y), "dx": synth.grad_digest(x.grad), "grads": digest_grads(m)}
    return out


def case_membank_dup(cm, seed):
    bank, D, B = 12, 64, 10
    mb = cm.memoryBank(bank_size=bank, device=torch.device("cpu"), dim=D)
    mem0 = synth.randn(seed, bank + 1, D)
    mb.memory.copy_(mem0)
    ids = torch.tensor([3, 7, 3, 1, 7, 7, 12, 0, 3, 5], dtype=torch.int64)       # duplicates: 3 x3, 7 x3
    data = synth.randn(seed + 1, B, D)
    old = mb(ids, data)
    return {"kind": "membank_dup", "seed": seed, "bank": bank, "D": D, "ids": ids.tolist(), "old": synth.grad_digest(old),
            "old_rows_equal_mem0": bool(torch.equal(old, mem0[ids])), "memory_after": synth.grad_digest(mb.memory),
            "row3": mb.memory[3].double().tolist(), "row7": mb.memory[7].double().tolist()}


# ---- non-default loss wrappers ---------------------------------------------------------------------------------------------
def _loss_inputs(seed, B, T, n_rep=1):
    ef = synth.randn(seed, n_rep * B, T, 8)
    sf = 0.5 * synth.randn(seed + 1, B, T, 8) + 0.5 * ef[:B]
    return ef.requires_grad_(True), sf.requires_grad_(True)


def _loss_case(model, ef, sf, ids, seed, names):
    fill_module(model, seed + 7)
    model.eval()
    out = model(ef, sf, ids)
    total = out[names.index("loss_total")]
    total.backward()
    d = {n: float(v) for n, v in zip(names, out)}
    d.update({"d_eeg": synth.grad_digest(ef.grad), "d_speech": synth.grad_digest(sf.grad), "grads": digest_grads(model)})
    return d


def case_loss_variants(cm, B, T, seed):
    out = {"kind": "loss_variants", "B": B, "T": T, "seed": seed, "bank": 20, "latent": 16, "n_rep": 3, "classes": 20, "latent2": 64}
    ids = torch.from_numpy(np.random.RandomState(seed + 2).permutation(20)[:B] + 1).to(torch.int64)
    out["ids"] = ids.tolist()
    std = ("loss_ce", "aux", "loss_total")
    # CLIPSim: bank rows live in the projected (latent) space
    mb = cm.memoryBank(bank_size=20, device=torch.device("cpu"), dim=16)
    mb.memory.copy_(synth.randn(seed + 5, 21, 16).abs())
    ef, sf = _loss_inputs(seed, B, T)
    m = cm.CLIPSim(IdTower(), IdTower(), mb, temperature=0.075, latent_dim=16, window_length=T, lambda_clip=1, lambda_average=0.5)
    out["clip_sim"] = _loss_case(m, ef, sf, ids, seed, std)
    out["clip_sim"]["memory_after"] = synth.grad_digest(mb.memory)
    for name, cls in (("clip_mp", cm.CLIPSimMultiplePositives), ("clip_mp_adapted", cm.CLIPSimMultiplePositivesAdapted)):
        ef, sf = _loss_inputs(seed + 20, B, T, n_rep=3)
        m = cls(IdTower(), IdTower(), temperature=0.075, window_length=T, lambda_clip=1, lambda_average=0.5)
        out[name] = _loss_case(m, ef, sf, ids, seed, std)
    ef, sf = _loss_inputs(seed + 40, B, T)
    m = cm.CLIPKLDNoLatentProj(IdTower(), IdTower(), latent_dimension=T * 8, number_of_classes=20, latent_dimension2=64,
                               temperature=0.075, window_length=T, lambda_clip=1, lambda_lower_bound=0.5, lambda_discriminative=0.5)
    out["clip_kld"] = _loss_case(m, ef, sf, ids, seed, ("loss_total", "loss_ce", "log_pmu2", "kld_z2"))
    ef, sf = _loss_inputs(seed + 60, B, T)
    m = cm.CLIPKLDWithLatentProj(IdTower(), IdTower(), latent_dimension=16, number_of_classes=20, temperature=0.075,
                                 window_length=T, lambda_clip=1, lambda_lower_bound=0.5, lambda_discriminative=0.5)
    out["clip_kld_latent_proj"] = _loss_case(m, ef, sf, ids, seed, ("loss_total", "loss_ce", "log_pmu2", "kld_z2"))
    ef, sf = _loss_inputs(seed + 80, B, T)
    m = cm.CLIPNoContrastiveLearning(IdTower(), IdTower(), window_length=T)
    out["no_contrastive"] = _loss_case(m, ef, sf, ids, seed, ("loss_ce", "aux", "loss_total"))
    return out


# ---- regression evaluation -------------------------------------------------------------------------------------------------
class SynthRegressionDataset:
    """Stand-in for dataset_loader.EEGDatasetSimdata (same constructor shape, same batch tuples): one batch of `windows` windows
    per EEG file, generated from the file's name; the envelope is a noisy mixture of two EEG channels so that a regression
    head has something to find."""
    windows = 12

    def __init__(self, files, audio_files, window_length, hop_length, batch_size=128, **kw):
        self.files, self.T = sorted(files), int(window_length)

    def __iter__(self):
        for f in self.files:
            name = os.path.basename(f)
            rs = np.random.RandomState(zlib.crc32(name.encode()) & 0x7fffffff)
            eeg = rs.standard_normal((self.windows, self.T, 64)).astype(np.float32)
            env = (0.7 * eeg[:, :, 3:4] - 0.4 * eeg[:, :, 17:18] + 0.5 * rs.standard_normal((self.windows, self.T, 1))).astype(np.float32)
            speech = rs.standard_normal((self.windows, self.T, 28)).astype(np.float32)
            yield name.split("_")[0], name.split("-audio-")[-1].split("_eeg")[0], eeg, speech, env


def regression_files():
    mk = lambda story: f"/synthetic/sub-001_-_run_-audio-{story}_eeg.npy"
    au = lambda story: f"/synthetic/{story}_-_env.npy"
    stories = {"train": ["st1", "st2", "st3"], "val": ["st4"], "test": ["st5", "st6"]}
    return {k: ([mk(s) for s in v], [au(s) for s in v]) for k, v in stories.items()}


def case_regression(cm, helpers, seed):
    out = {"kind": "regression", "seed": seed}
    # single step: RegressionModel + PearsonLoss, loss and gradients
    B, Cin, T = 6, 8, 320
    reg = helpers.RegressionModel(Cin, output_dim=2)
    fill_module(reg, seed)
    x = synth.randn(seed + 1, B, Cin, T).requires_grad_(True)
    y = synth.randn(seed + 2, B, 2, T) + 0.3 * x.detach()[:, :2]
    pred = reg(x)
    loss = helpers.PearsonLoss()(pred, y)
    loss.sum().backward()
    out["step"] = {"B": B, "Cin": Cin, "T": T, "pred": synth.grad_digest(pred), "loss": loss.double().tolist(),
                   "loss_mean": float(helpers.PearsonLossMean()(pred.detach(), y)), "dx": synth.grad_digest(x.grad), "grads": digest_grads(reg)}
    # end to end: the reference's evaluate_model_do_regression_sub_specific on the synthetic dataset stand-in
    model = build_full_model(cm, 1, 320, 0, seed + 30, speech="smallConv")
    files = regression_files()
    helpers.EEGDatasetSimdata = SynthRegressionDataset
    with tempfile.TemporaryDirectory() as root:
        torch.manual_seed(seed)
        ev = helpers.evaluate_model_do_regression_sub_specific(model, files["train"][0], files["val"][0], files["test"][0],
                                                              files["train"][1], files["val"][1], files["test"][1],
                                                              torch.device("cpu"), root, window_length=5, fs=64)
        lines = open(os.path.join(root, "loss_regression.txt")).read().strip().splitlines()
    out["fit"] = {"evaluation": ev, "epochs": len(lines), "first": lines[0], "last": lines[-1], "model_seed": seed + 30,
                  "torch_seed": seed}
    return out


def case_optim(seed):
    out = {"kind": "optim", "seed": seed, "steps": 4}
    for name, cls, kw in (("adam", torch.optim.Adam, dict(weight_decay=0.0)), ("adam_wd", torch.optim.Adam, dict(weight_decay=0.05)),
                          ("adamw_amsgrad", torch.optim.AdamW, dict(weight_decay=0.01, amsgrad=True)),
                          ("adam_amsgrad", torch.optim.Adam, dict(amsgrad=True))):
        p = nn.Parameter(synth.randn(seed, 257).clone())
        opt = cls([p], lr=1e-3, betas=(0.9, 0.999), **kw)
        for s in range(4):
            p.grad = synth.randn(seed + 1 + s, 257) * (3.0 if s == 1 else 1.0)     # a spike so that amsgrad's max matters
            opt.step()
        out[name] = p.detach().double().tolist()
    return out


def main():
    cm, vl, helpers = _import_reference()
    torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))
    cases = {
        "tower_d10_T320_eval": case_tower_big(cm, 10, 320, 8, 900, False),
        "tower_d10_T320_train": case_tower_big(cm, 10, 320, 8, 910, True),
        "full_d10_T320_eval": case_full_big(cm, 10, 320, 8, 16, 920, False),
        "full_d10_T320_train": case_full_big(cm, 10, 320, 8, 16, 930, True),
        "head_B4096_D2560": case_head_big(cm, 4096, 2560, 0.075, 940),
        "submodules": case_submodules(cm, 3, 192, 950),
        "membank_dup": case_membank_dup(cm, 960),
        "loss_variants": case_loss_variants(cm, 8, 192, 970),
        "regression": case_regression(cm, helpers, 980),
        "optim": case_optim(990),
    }
    meta = {"torch": torch.__version__, "numpy": np.__version__, "reference": REF,
            "note": "round-2 cases: outputs of the unmodified reference modules / functions on oracle/synth.py inputs"}
    path = os.path.join(OUT, "reference_golden_r2.json")
    with open(path, "w") as f:
        json.dump({"meta": meta, "cases": cases}, f, indent=1)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
