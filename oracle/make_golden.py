"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.json from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden

It imports clip_model.py / vlaai.py / train_clip_helper_functions.py from
/root/reference, loads the deterministic synthetic weights of oracle/synth.py into
the reference modules, runs them on CPU and records small digests (norm, sum,
abs-sum, strided samples) of every output and gradient.  The GPU box has no
/root/reference; tests there regenerate the same inputs from the seeds stored in
the golden file and compare against these recorded digests.

Train-mode cases patch each nn.Dropout *instance* of the reference model so that
it multiplies by the Philox mask of oracle/philox_ref.py (the reference's own
torch-RNG masks are not reproducible by any other implementation).
"""
import json
import os
import pickle
import sys
import tempfile
import types

import numpy as np
import torch
import torch.nn as nn

REF = os.environ.get("EEGCLIP_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

from . import eegclip_oracle as O
from . import synth


def _import_reference():
    sys.path.insert(0, REF)
    if "braindecode" not in sys.modules:  # dataset_loader.py:5 pulls it in; unused on this path
        bd = types.ModuleType("braindecode")
        aug = types.ModuleType("braindecode.augmentation")
        for n in ("SignFlip", "FTSurrogate", "FrequencyShift", "BandstopFilter", "GaussianNoise",
                  "SmoothTimeMask", "ChannelsDropout", "ChannelsShuffle"):
            setattr(aug, n, type(n, (), {}))
        bd.augmentation = aug
        sys.modules["braindecode"] = bd
        sys.modules["braindecode.augmentation"] = aug
    import clip_model
    import vlaai
    import train_clip_helper_functions as helpers
    return clip_model, vlaai, helpers


def patch_dropouts_interleaved(model, depth, drop, conformer_attr="conformer_{i}", offset=0):
    """Route every nn.Dropout of an EEGConformerInterleaved through the Philox policy."""
    def bind(mod, layer, site, order=None):
        p = mod.p
        mod.forward = lambda x, _p=p, _l=layer, _s=site, _o=order: drop(x, _p, _l, _s, order=_o)

    for i in range(depth):
        bind(getattr(model, f"conv_{i}").dropout, i, O.SITE_CONV, (0, 2, 1))
        blk = getattr(model, f"conformer_{i}")[0]
        bind(blk[0].fn[1].att_drop, i, O.SITE_ATTN)
        bind(blk[0].fn[2], i, O.SITE_PROJ)
        bind(blk[1].fn[1][2], i, O.SITE_FFN_HID)
        bind(blk[1].fn[2], i, O.SITE_FFN_OUT)


def digest_grads(model, prefix=""):
    return {prefix + k: synth.grad_digest(p.grad if p.grad is not None else torch.zeros_like(p))
            for k, p in model.named_parameters()}


def case_tower_interleaved(cm, depth, T, B, seed, train):
    model = cm.EEGConformerInterleaved(output_dim=8, time_dimension=T, depth=depth)
    sd = synth.make_state_dict(synth.interleaved_shapes(depth, T), seed)
    model.load_state_dict(sd)
    x = synth.randn(seed + 1, B, T, 64).requires_grad_(True)
    w = synth.randn(seed + 2, B, T, 8)
    if train:
        model.train()
        patch_dropouts_interleaved(model, depth, O.Drop(True, seed=seed + 3))
    else:
        model.eval()
    y = model(x)
    (y * w).sum().backward()
    return {
        "kind": "interleaved", "depth": depth, "T": T, "B": B, "seed": seed, "train": train,
        "drop_seed": seed + 3,
        "out": synth.grad_digest(y), "dx": synth.grad_digest(x.grad), "grads": digest_grads(model),
    }


def case_tower_conformer(cm, n_conv, depth, T, B, seed):
    model = cm.EEGConformer(output_dim=8, filters=(64,) * n_conv, kernels=(64,) * n_conv, time_dimension=T, depth=depth)
    sd = synth.make_state_dict(synth.conformer_shapes(n_conv, depth, T), seed)
    model.load_state_dict(sd)
    model.eval()
    x = synth.randn(seed + 1, B, T, 64).requires_grad_(True)
    w = synth.randn(seed + 2, B, T, 8)
    y = model(x)
    (y * w).sum().backward()
    return {"kind": "conformer", "n_conv": n_conv, "depth": depth, "T": T, "B": B, "seed": seed,
            "out": synth.grad_digest(y), "dx": synth.grad_digest(x.grad), "grads": digest_grads(model)}


def case_head(cm, B, D, tau, seed):
    model = cm.CLIP(nn.Identity(), nn.Identity(), temperature=tau)
    E = synth.randn(seed, B, D).requires_grad_(True)
    S = (0.5 * synth.randn(seed + 1, B, D) + 0.5 * E.detach()).requires_grad_(True)
    loss = model(E, S)
    loss.backward()
    return {"kind": "head", "B": B, "D": D, "tau": tau, "seed": seed, "loss": float(loss),
            "dE": synth.grad_digest(E.grad), "dS": synth.grad_digest(S.grad), "dtau": float(model.temperature.grad)}


def case_clipsim(cm, B, T, bank, seed, lam_avg):
    torch.manual_seed(seed)
    mb = cm.memoryBank(bank_size=bank, device=torch.device("cpu"), dim=T * 8)
    mem0 = synth.randn(seed + 5, bank + 1, T * 8).abs()
    mb.memory.copy_(mem0)
    model = cm.CLIPSimNoLatentProj(nn.Identity(), nn.Identity(), mb, temperature=0.075, window_length=T,
                                   lambda_clip=1, lambda_average=lam_avg)
    ef = synth.randn(seed, B, T, 8).requires_grad_(True)
    sf = (0.5 * synth.randn(seed + 1, B, T, 8) + 0.5 * ef.detach()).requires_grad_(True)
    ids = torch.from_numpy(np.random.RandomState(seed + 2).permutation(bank)[:B] + 1).to(torch.int64)
    model.eval()  # the bank is updated in eval mode too (clip_model.py:743 has no training guard)
    l_ce, l_avg, l_tot = model(ef, sf, ids)
    l_tot.backward()
    return {"kind": "clipsim", "B": B, "T": T, "bank": bank, "seed": seed, "lam_avg": lam_avg,
            "ids": ids.tolist(), "loss_ce": float(l_ce), "avg_loss": float(l_avg), "loss_total": float(l_tot),
            "d_eeg": synth.grad_digest(ef.grad), "d_speech": synth.grad_digest(sf.grad),
            "dtau": float(model.temperature.grad), "dtau_eeg": float(model.temperature_eeg.grad),
            "memory_after": synth.grad_digest(mb.memory)}


def case_speech(cm, which, T, B, seed):
    if which == "smallConv":
        model = cm.SpeechSmallConv(output_dim=8, ks_temporal=16, dropout_rate=0.4, speech_dim=1024, time_dimension=T)
        sd = synth.make_state_dict(synth.small_conv_shapes(T), seed)
    else:
        model = cm.EEGConvLSTM(units_lstm=128, output_dim=8, dropout_rate=0.4, eeg_dim=1024, filters=(64,), kernels=(32,),
                               input_channels=1024, time_dimension=T)
        sd = synth.make_state_dict(synth.conv_lstm_shapes(T), seed)
    model.load_state_dict(sd)
    model.eval()
    x = synth.randn(seed + 1, B, T, 1024).requires_grad_(True)
    w = synth.randn(seed + 2, B, T, 8)
    y = model(x)
    (y * w).sum().backward()
    return {"kind": "speech", "which": which, "T": T, "B": B, "seed": seed, "out": synth.grad_digest(y),
            "dx": synth.grad_digest(x.grad), "grads": digest_grads(model)}


def build_full_model(cm, depth, T, bank, seed, speech="convLSTM"):
    eeg = cm.EEGConformerInterleaved(output_dim=8, time_dimension=T, depth=depth)
    eeg.load_state_dict(synth.make_state_dict(synth.interleaved_shapes(depth, T), seed))
    if speech == "convLSTM":
        sp = cm.EEGConvLSTM(units_lstm=128, output_dim=8, dropout_rate=0.4, eeg_dim=1024, filters=(64,), kernels=(32,),
                            input_channels=1024, time_dimension=T)
        sp.load_state_dict(synth.make_state_dict(synth.conv_lstm_shapes(T), seed + 1))
    else:
        sp = cm.SpeechSmallConv(output_dim=8, ks_temporal=16, dropout_rate=0.4, speech_dim=1024, time_dimension=T)
        sp.load_state_dict(synth.make_state_dict(synth.small_conv_shapes(T), seed + 1))
    mb = None
    if bank:
        mb = cm.memoryBank(bank_size=bank, device=torch.device("cpu"), dim=T * 8)
        mb.memory.copy_(synth.randn(seed + 5, bank + 1, T * 8).abs())
    return cm.CLIPSimNoLatentProj(eeg, sp, mb, temperature=0.075, window_length=T, lambda_clip=1, lambda_average=0.0)


def case_full(cm, depth, T, B, bank, seed):
    model = build_full_model(cm, depth, T, bank, seed)
    model.eval()
    eeg = synth.randn(seed + 10, B, T, 64)
    sp = synth.randn(seed + 11, B, T, 1024)
    ids = torch.arange(1, B + 1, dtype=torch.int64)
    l_ce, l_avg, l_tot = model(eeg, sp, ids)
    l_tot.backward()
    return {"kind": "full", "depth": depth, "T": T, "B": B, "bank": bank, "seed": seed,
            "loss_ce": float(l_ce), "avg_loss": float(l_avg), "loss_total": float(l_tot), "grads": digest_grads(model)}


def case_vlaai(vl, B, seed):
    model = vl.VLAAI()
    model.load_state_dict(synth.make_state_dict(synth.vlaai_shapes(320), seed))
    model.eval()
    x = synth.randn(seed + 1, B, 320, 64).requires_grad_(True)
    w = synth.randn(seed + 2, B, 64, 320)
    y = model(x)
    (y * w).sum().backward()
    return {"kind": "vlaai", "B": B, "seed": seed, "out": synth.grad_digest(y), "dx": synth.grad_digest(x.grad),
            "grads": digest_grads(model)}


def case_adamw(seed):
    p0 = synth.randn(seed, 257)
    p = nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([p], lr=1e-3, betas=(0.9, 0.999), weight_decay=0.01)
    for s in range(3):
        p.grad = synth.randn(seed + 1 + s, 257)
        opt.step()
    return {"kind": "adamw", "seed": seed, "steps": 3, "p": p.detach().double().tolist()}


def write_mm_fixture(root, T, n_sub, n_seg, n_stim, seed):
    """On-disk layout read by train_clip_helper_functions.py:56-58,86-101,121-140."""
    rs = np.random.RandomState(seed)
    os.makedirs(os.path.join(root, "labels"), exist_ok=True)
    os.makedirs(os.path.join(root, "wav2vec_segments_wholefile_64hz"), exist_ok=True)
    stim = {f"story{j // 6}_-_seg{j}": rs.standard_normal((T, 1024)).astype(np.float32) for j in range(n_stim)}
    keys = list(stim)
    for s in range(n_stim // 6):
        part = {k: stim[k] for k in keys[6 * s:6 * s + 6]}
        with open(os.path.join(root, "wav2vec_segments_wholefile_64hz", f"x_-_story{s}_-_wav2vec_19.pkl"), "wb") as f:
            pickle.dump(part, f)
    labels = {}
    for u in range(n_sub):
        mapping = {}
        for g in range(n_seg):
            sid = f"sub-{u:03d}_{g}"
            a, b = rs.choice(n_stim, 2, replace=False)
            lab = int(rs.randint(2))
            mapping[sid] = [(rs.standard_normal((1, T, 64)) * (1 + u) + 0.3 * g).tolist(), keys[a] + ".npy", keys[b] + ".npy"]
            labels[sid] = lab
        with open(os.path.join(root, f"sub-{u:03d}.json"), "w") as f:
            json.dump(mapping, f)
    with open(os.path.join(root, "labels", "labels.json"), "w") as f:
        json.dump(labels, f)


def case_mm(cm, helpers, T, seed):
    model = build_full_model(cm, 1, T, 0, seed, speech="smallConv")
    with tempfile.TemporaryDirectory() as root:
        write_mm_fixture(root, T, n_sub=2, n_seg=7, n_stim=12, seed=seed + 20)
        ev, ev_logits, ev_top, ev_top_logits = helpers.evaluate_model_challenge_2023_mm(
            model, torch.device("cpu"), speech_feature="wav2vec_19", eeg_folder=root)
    return {"kind": "mm", "T": T, "seed": seed, "n_sub": 2, "n_seg": 7, "n_stim": 12,
            "evaluation": ev, "logits": ev_logits, "top_x": ev_top,
            "correct_keys_idx": {k: v["correct_keys_idx"] for k, v in ev_top_logits.items()},
            "bank_logits_digest": {k: synth.grad_digest(torch.tensor(v["logits"])) for k, v in ev_top_logits.items()}}


def main():
    cm, vl, helpers = _import_reference()
    torch.set_num_threads(8)
    os.makedirs(OUT, exist_ok=True)
    cases = {
        "tower_d2_T192_eval": case_tower_interleaved(cm, 2, 192, 3, 100, False),
        "tower_d1_T320_eval": case_tower_interleaved(cm, 1, 320, 2, 110, False),
        "tower_d2_T192_train": case_tower_interleaved(cm, 2, 192, 2, 120, True),
        "conformer_c2_d2_T192_eval": case_tower_conformer(cm, 2, 2, 192, 2, 130),
        "head_B64_D2560": case_head(cm, 64, 2560, 1.0, 200),
        "head_B16_D1536": case_head(cm, 16, 1536, 0.075, 210),
        "head_B96_D200": case_head(cm, 96, 200, 2.0, 220),
        "clipsim_lam0": case_clipsim(cm, 8, 192, 20, 300, 0.0),
        "clipsim_lam1": case_clipsim(cm, 8, 192, 20, 310, 1.0),
        "speech_smallConv": case_speech(cm, "smallConv", 192, 2, 400),
        "speech_convLSTM": case_speech(cm, "convLSTM", 192, 2, 410),
        "full_d2_T192": case_full(cm, 2, 192, 4, 16, 500),
        "vlaai_B2": case_vlaai(vl, 2, 600),
        "adamw": case_adamw(700),
        "mm_T192": case_mm(cm, helpers, 192, 800),
    }
    meta = {"torch": torch.__version__, "numpy": np.__version__, "reference": REF,
            "note": "outputs of the unmodified reference modules on oracle/synth.py inputs"}
    with open(os.path.join(OUT, "reference_golden.json"), "w") as f:
        json.dump({"meta": meta, "cases": cases}, f, indent=1)
    print("wrote", os.path.join(OUT, "reference_golden.json"), os.path.getsize(os.path.join(OUT, "reference_golden.json")), "bytes")


if __name__ == "__main__":
    main()
