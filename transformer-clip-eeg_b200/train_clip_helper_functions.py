"""Mirror of the evaluation entry points of the reference ``train_clip_helper_functions.py`` on the B200 kernels.

``evaluate_model_challenge_2023_mm`` keeps the reference signature, on-disk fixture format and return
values (train_clip_helper_functions.py:51-205) but scores on the GPU kernels:
  * per-subject mean-variance normalisation of the EEG windows (:133-136) is eegclip_mvn_normalize;
  * the two candidate similarities per window are row-dots (eegclip_mm_rowdots) instead of the
    reference's N x N matmul followed by ``torch.diag`` (:159-160);
  * the all-stimuli logits are one similarity GEMM (eegclip_mm_bank_logits) and the top-x ranking (:187) is
    eegclip_row_topk; ``bank_topk`` streams the bank in column chunks so the N x M logits never exist in full.
The downstream regression evaluation (:443-1103) -- ``RegressionModel`` (Conv1d(latent -> n, 32, 'same') + LeakyReLU),
``PearsonLoss`` and the Adam fit with early stopping -- runs on eegclip_conv_small_* / eegclip_pearson_* and the
package's Adam.  File discovery / JSON / pickle handling is host plumbing and follows the reference's layout
(:56-101,121-140); ``EEGDatasetSimdata`` itself (dataset_loader.py) is the caller's: pass it as ``dataset_cls``.
"""
import ctypes
import glob
import json
import os
import pickle

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib as L


def printf(s, file):
    print(s)
    with open(file, 'a') as f:
        f.write(s + '\n')


def load_labels_match_mismatch_2023(path_true_labels):
    labels = {}
    for file in glob.glob(os.path.join(path_true_labels, '*.json')):
        with open(file, 'r') as f:
            labels.update(json.load(f))
    return labels


def l2_normalize_rows(x):
    x = L.f32c(x)
    out = torch.empty_like(x)
    L.call("eegclip_l2norm_forward", L.ptr(x), L.ptr(out), None, x.shape[0], x.shape[1], L.stream())
    return out


def mm_scores(eeg_emb, cand_emb):
    """eeg (N,D), candidates (N,K,D) -> scores (K,N) and argmax over K (N,)."""
    eeg_emb, cand_emb = L.f32c(eeg_emb), L.f32c(cand_emb)
    N, K, D = cand_emb.shape
    scores = torch.empty(K, N, dtype=torch.float32, device=eeg_emb.device)
    choice = torch.empty(N, dtype=torch.int64, device=eeg_emb.device)
    L.call("eegclip_mm_rowdots", L.ptr(eeg_emb), L.ptr(cand_emb), L.ptr(scores), L.ptr(choice), N, K, D, L.stream())
    return scores, choice


def bank_logits(eeg_emb, bank_emb):
    """eeg (N,D) . bank (M,D)^T -> (N,M)."""
    eeg_emb, bank_emb = L.f32c(eeg_emb), L.f32c(bank_emb)
    N, D = eeg_emb.shape
    M = bank_emb.shape[0]
    out = torch.empty(N, M, dtype=torch.float32, device=eeg_emb.device)
    nb = ctypes.c_size_t()
    L.call("eegclip_mm_bank_workspace", N, M, D, ctypes.byref(nb))
    scratch = torch.empty(max(nb.value, 16), dtype=torch.uint8, device=eeg_emb.device)
    L.call("eegclip_mm_bank_logits", L.ptr(eeg_emb), L.ptr(bank_emb), L.ptr(out), N, M, D, L.default_math(), L.ptr(scratch), L.stream())
    return out


def mvn_per_subject(eeg):
    """(N,T,C) windows of one subject -> (x - mean_c) / std_c over all windows and time samples (population std), on the device
    (train_clip_helper_functions.py:133-136 does this with numpy on the host)."""
    eeg = L.f32c(eeg)
    C = eeg.shape[-1]
    nb = ctypes.c_size_t()
    L.call("eegclip_mvn_workspace", C, ctypes.byref(nb))
    scratch = torch.empty(max(nb.value, 16), dtype=torch.uint8, device=eeg.device)
    out = torch.empty_like(eeg)
    L.call("eegclip_mvn_normalize", L.ptr(eeg), L.ptr(out), eeg.numel() // C, C, L.ptr(scratch), L.stream())
    return out


def row_topk(x, k, col_offset=0):
    """Per-row top-k of a 2-D fp32 CUDA tensor (row stride may exceed the width): (values (N,k) descending, columns (N,k) int64)."""
    if x.dim() != 2 or x.stride(1) != 1 or x.dtype != torch.float32 or not x.is_cuda:
        raise L.EegclipError("row_topk: needs a 2-D fp32 CUDA tensor with unit column stride")
    N, M = x.shape
    vals = torch.empty(N, k, dtype=torch.float32, device=x.device)
    idx = torch.empty(N, k, dtype=torch.int64, device=x.device)
    L.call("eegclip_row_topk", x.data_ptr(), x.stride(0), N, M, k, int(col_offset), L.ptr(vals), L.ptr(idx), L.stream())
    return vals, idx


BANK_CHUNK = 16384   # bank columns scored per pass: N x 16384 logits (268 MB at N = 4096) instead of N x M (1.6 GB at M = 1e5)


def bank_topk(eeg_emb, bank_emb, k, group=None, logits_fn=None, topk_fn=None, chunk=BANK_CHUNK):
    """Top-k stimuli of the bank for every EEG window (train_clip_helper_functions.py:182-187), optionally with the bank
    SHARDED over a process group (BASELINE config 4): every rank holds all N windows and its own slice of the M
    candidates (``bank_emb`` is the local slice, slices concatenated in rank order form the bank), scores locally, keeps
    its local top-k, and the k*world candidates are all-gathered and merged.  Returns (values (N,k), global indices (N,k)),
    identical on all ranks and identical to a single-process top-k over the whole bank (ties aside).
    The local slice is scored in column chunks: similarity GEMM -> per-row top-k -> merge with the running top-k, so the full
    N x M logits matrix is never written (nor read back by a sort)."""
    import torch.distributed as dist
    logits_fn = logits_fn or bank_logits
    topk_fn = topk_fn or row_topk
    world = dist.get_world_size(group) if (group is not None and dist.is_initialized()) else 1
    m_loc = bank_emb.shape[0]
    best = None
    for c0 in range(0, m_loc, chunk):
        logits = logits_fn(eeg_emb, bank_emb[c0:c0 + chunk])
        kk = min(k, logits.shape[1])
        v, i = topk_fn(logits, kk)
        i = i + c0
        if best is None:
            best = (v, i)
        else:
            cat_v, cat_i = torch.cat([best[0], v], dim=1).contiguous(), torch.cat([best[1], i], dim=1)
            top_v, pos = topk_fn(cat_v, min(k, cat_v.shape[1]))
            best = (top_v, torch.gather(cat_i, 1, pos))
    vals, idx = best
    kk = vals.shape[1]
    if world == 1:
        return vals, idx
    rank = dist.get_rank(group)
    sizes = torch.zeros(world, dtype=torch.int64, device=vals.device)
    sizes[rank] = m_loc
    dist.all_reduce(sizes, group=group)
    offset = int(sizes[:rank].sum())
    if kk < k:   # pad so that every rank contributes k columns
        pad = k - kk
        vals = torch.cat([vals, vals.new_full((vals.shape[0], pad), float("-inf"))], dim=1)
        idx = torch.cat([idx, idx.new_zeros((idx.shape[0], pad))], dim=1)
    idx = idx + offset
    all_v = [torch.empty_like(vals) for _ in range(world)]
    all_i = [torch.empty_like(idx) for _ in range(world)]
    dist.all_gather(all_v, vals.contiguous(), group=group)
    dist.all_gather(all_i, idx.contiguous(), group=group)
    cat_v, cat_i = torch.cat(all_v, dim=1).contiguous(), torch.cat(all_i, dim=1)
    kout = min(k, int(sizes.sum()))
    top_v, pos = topk_fn(cat_v, kout)
    return top_v, torch.gather(cat_i, 1, pos)


def evaluate_model_challenge_2023_mm(model, device, subject=None, speech_feature='omsimel', eeg_folder=''):
    labels_all = load_labels_match_mismatch_2023(os.path.join(eeg_folder, 'labels'))
    stimulus_folder = os.path.join(eeg_folder, 'wav2vec_segments_wholefile_64hz/')
    model.eval()
    evaluation, evaluation_with_logits, evaluation_top_x, evaluation_top_x_with_logits = {}, {}, {}, {}

    if subject is not None:
        mappings = [os.path.join(eeg_folder, f'{subject}.json')]
        try:
            first = json.load(open(mappings[0]))
        except Exception:
            print(f'error with {mappings[0]}')
            return evaluation, evaluation_with_logits
        used = {first[k][1].split('_-_')[0] for k in first}
        stim_files = [p for p in glob.glob(os.path.join(stimulus_folder, f'*{speech_feature}.pkl'))
                      if os.path.basename(p).split('_-_')[1] in used]
    else:
        mappings = glob.glob(os.path.join(eeg_folder, 'sub*.json'))
        stim_files = glob.glob(os.path.join(stimulus_folder, f'*{speech_feature}.pkl'))

    n_raw, emb = 0, {}
    for path in stim_files:
        with open(path, 'rb') as f:
            data = pickle.load(f)
        n_raw += len(data)
        keys = list(data.keys())
        if data[keys[-1]].shape != data[keys[-2]].shape:   # ragged last segment is dropped (:99-100)
            keys = keys[:-1]
        seg = torch.from_numpy(np.stack([data[k] for k in keys])).to(device, dtype=torch.float)
        with torch.no_grad():
            e = l2_normalize_rows(torch.flatten(model.speechModel(seg), start_dim=1))
        emb.update({k: e[i] for i, k in enumerate(keys)})
    print(f'number of test stimuli: {n_raw}')
    print(f'number of test stimuli embeddings: {len(emb)}')
    stim_keys = list(emb.keys())
    if not stim_keys:
        print("No test stimuli found for match-mismatch evaluation, skipping.")
        return {}, {}, {}, {}
    bank = torch.stack([emb[k] for k in stim_keys])
    key_pos = {k: i for i, k in enumerate(stim_keys)}

    for path in mappings:
        sub = os.path.basename(path).split('.')[0]
        print(f'evaluating {sub}')
        try:
            mapping = json.load(open(path))
        except Exception:
            print(f'error with {path}')
            continue
        ids = list(mapping.keys())
        eeg = np.squeeze(np.stack([mapping[k][0] for k in ids]))
        labels = [labels_all[k] for k in ids]
        correct_keys = [mapping[k][labels_all[k] + 1].split('.')[0] for k in ids]
        with torch.no_grad():
            eeg_dev = mvn_per_subject(torch.from_numpy(eeg).to(device, dtype=torch.float))             # per-subject MVN (:136)
            e = l2_normalize_rows(torch.flatten(model.eegModel(eeg_dev), start_dim=1))
            cand = torch.stack([torch.stack([emb[mapping[k][1].split('.')[0]], emb[mapping[k][2].split('.')[0]]]) for k in ids])
            scores, choice = mm_scores(e, cand)                    # (2,N), (N,)
            truth = torch.tensor(labels, device=device, dtype=torch.int64)
            acc = (choice == truth).float().mean()
            evaluation[sub + '_mvn'] = acc.item()
            print(f"evaluation mm with mvn : {evaluation[sub + '_mvn']}, {sub}")
            sc = scores.cpu()
            evaluation_with_logits[sub] = {k: (sc[:, i].tolist(), labels[i]) for i, k in enumerate(ids)}

            idx = torch.tensor([key_pos[k] for k in correct_keys], dtype=torch.float32, device=device)
            logits = bank_logits(e, bank)
            maxtop = min(100, logits.shape[1])
            top = row_topk(logits, maxtop)[1].cpu().numpy()
            lab = np.repeat(idx.to(torch.int).cpu().numpy().astype(np.int32), maxtop).reshape(len(ids), -1)
            correct_top = np.mean(np.cumsum(np.equal(lab, top), axis=1), axis=0)
            evaluation_top_x[sub] = correct_top.tolist()
            evaluation_top_x_with_logits[sub] = {'logits': logits.tolist(), 'correct_keys_idx': idx.tolist(),
                                                 'correct_top': correct_top.tolist()}
            print(f"evaluation mm top x: {sub} : top1 {evaluation_top_x[sub][0] * 100}, top10: {evaluation_top_x[sub][9] * 100}")
    return evaluation, evaluation_with_logits, evaluation_top_x, evaluation_top_x_with_logits


# ---------------------------------------------------------------------------------------------------
# Downstream regression evaluation (train_clip_helper_functions.py:443-1103, 1107-1140)
# ---------------------------------------------------------------------------------------------------
class _ConvSmallFn(torch.autograd.Function):
    """LeakyReLU(Conv1d(Cin -> Cout, K, 'same')(x)) on channel-major (B,Cin,T)."""

    @staticmethod
    def forward(ctx, x, w, b):
        x, w = L.f32c(x), L.f32c(w)
        b = L.f32c(b) if b is not None else None
        B, Cin, T = x.shape
        Cout, _, K = w.shape
        out = torch.empty(B, Cout, T, dtype=torch.float32, device=x.device)
        L.call("eegclip_conv_small_forward", L.ptr(x), L.ptr(w), L.ptr(b), L.ptr(out), B, Cin, Cout, T, K, L.stream())
        ctx.saved, ctx.has_b = (x, w, out), b is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w, out = ctx.saved
        B, Cin, T = x.shape
        Cout, _, K = w.shape
        nb = ctypes.c_size_t()
        L.call("eegclip_conv_small_workspace", B, Cin, Cout, K, ctypes.byref(nb))
        scratch = torch.empty(max(nb.value, 16), dtype=torch.uint8, device=x.device)
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(w)
        db = torch.empty(Cout, dtype=torch.float32, device=x.device) if ctx.has_b else None
        L.call("eegclip_conv_small_backward", L.ptr(x), L.ptr(w), L.ptr(out), L.ptr(L.f32c(dout)), L.ptr(dx), L.ptr(dw), L.ptr(db), B, Cin,
               Cout, T, K, L.ptr(scratch), L.stream())
        return dx, dw, db


class _PearsonFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        x, y = L.f32c(x), L.f32c(y)
        B, C, T = x.shape
        r = torch.empty(B, C, dtype=torch.float32, device=x.device)
        loss = torch.empty(C, dtype=torch.float32, device=x.device)
        L.call("eegclip_pearson_forward", L.ptr(x), L.ptr(y), L.ptr(r), L.ptr(loss), B, C, T, L.stream())
        ctx.saved = (x, y)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        x, y = ctx.saved
        B, C, T = x.shape
        dl = L.f32c(dloss)
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dy = torch.empty_like(y) if ctx.needs_input_grad[1] else None
        if dx is not None:
            L.call("eegclip_pearson_backward", L.ptr(x), L.ptr(y), L.ptr(dl), L.ptr(dx), B, C, T, L.stream())
        if dy is not None:                                  # the correlation is symmetric in its arguments
            L.call("eegclip_pearson_backward", L.ptr(y), L.ptr(x), L.ptr(dl), L.ptr(dy), B, C, T, L.stream())
        return dx, dy


class PearsonLoss(torch.nn.Module):
    """train_clip_helper_functions.py:1107-1118: minus the batch-mean Pearson correlation along time, per channel; (B,C,T) -> (C,)."""

    def forward(self, x, y):
        if not x.is_cuda:
            raise L.EegclipError("PearsonLoss: CUDA tensors only (no CPU fallback on this path)")
        return _PearsonFn.apply(x, y)


class PearsonLossMean(torch.nn.Module):
    """train_clip_helper_functions.py:1120-1129."""

    def __init__(self):
        super().__init__()
        self.pearsonCalculator = PearsonLoss()

    def forward(self, x, y):
        return self.pearsonCalculator(x, y).mean()


class RegressionModel(torch.nn.Module):
    """train_clip_helper_functions.py:1132-1140: Conv1d(input_dim -> output_dim, receptive_field, 'same') + LeakyReLU."""

    def __init__(self, input_dim, output_dim=1, receptive_field=32):
        super().__init__()
        self.conv = torch.nn.Conv1d(input_dim, output_dim, kernel_size=receptive_field, padding='same')
        self.activation = torch.nn.LeakyReLU()
        self.model = torch.nn.Sequential(self.conv, self.activation)

    def forward(self, x):
        if not x.is_cuda:
            raise L.EegclipError("RegressionModel: CUDA tensors only (no CPU fallback on this path)")
        return _ConvSmallFn.apply(x, self.conv.weight, self.conv.bias)


def _match_time_axis(emb, n_time):
    """Stride handling of :521-535: nearest-neighbour upsampling by the integer stride, then pad with the last frame / crop."""
    if emb.shape[1] == n_time:
        return emb
    stride = int(n_time / emb.shape[1])
    emb = emb.repeat_interleave(stride, dim=1)
    if emb.shape[1] < n_time:
        emb = torch.cat([emb, emb[:, -1:, :].repeat(1, n_time - emb.shape[1], 1)], dim=1)
    return emb[:, :n_time, :]


def regression_embeddings(model, generator, device):
    """Embeddings / envelopes / speech features of every batch of an ``EEGDatasetSimdata``-style generator (:505-541), channel-major:
    batches are (sub, story, eeg (b,T,64), speech (b,T,F), env (b,T,n[,1])) numpy tuples."""
    embs, envs, mels = [], [], []
    with torch.no_grad():
        for data in generator:
            if len(data) != 5:
                print(f'error with {data[0]} {data[1]}')
                continue
            env = data[4][:, :, :, 0] if data[4].ndim == 4 else data[4]
            eeg = torch.from_numpy(np.asarray(data[2])).to(device, dtype=torch.float)
            speech = torch.from_numpy(np.asarray(data[3])).to(device, dtype=torch.float)[:eeg.shape[0]]
            env = torch.from_numpy(np.asarray(env)).to(device, dtype=torch.float)[:eeg.shape[0]]
            embs.append(_match_time_axis(model.eegModel(eeg), env.shape[1]))
            envs.append(env)
            mels.append(speech)
    cat = lambda ts: torch.cat(ts, dim=0).permute(0, 2, 1).contiguous()
    return cat(embs), cat(envs), cat(mels)


def fit_regression_head(train_x, train_y, val_x, val_y, device, ckpt_path, file_loss=None, epochs=250, patience=10, batch_size=64,
                        lr=0.001, val_batched=True):
    """The fit of :620-675 / :960-1010: RegressionModel + PearsonLoss + Adam(lr), early stopping on the validation loss with the
    best checkpoint restored.  x (N,latent,T), y (N,n,T) channel-major.  Returns (model, history [(train, val) per epoch])."""
    from .optim import Adam
    reg = RegressionModel(train_x.shape[1], output_dim=train_y.shape[1])
    reg.to(device)
    crit = PearsonLoss()
    opt = Adam(reg.parameters(), lr=lr)
    best_epoch, best_val, hist = 0, float('inf'), []
    for epoch in range(epochs):
        losses = []
        for i in range(0, train_x.shape[0], batch_size):
            loss = crit(reg(train_x[i:i + batch_size].to(device)), train_y[i:i + batch_size].to(device))
            losses.append(loss.detach())
            opt.zero_grad()
            loss.backward()
            opt.step()
        with torch.no_grad():
            if val_batched:
                vl = [crit(reg(val_x[i:i + batch_size].to(device)), val_y[i:i + batch_size].to(device)) for i in range(0, val_x.shape[0], batch_size)]
            else:
                vl = [crit(reg(val_x.to(device)), val_y.to(device))]
        tr, va = float(torch.stack(losses).mean()), float(torch.stack(vl).mean())   # one host sync per epoch (the reference: per batch)
        hist.append((tr, va))
        if file_loss:
            printf(f'epoch {epoch}, loss {tr}, val_loss {va}', file_loss)
        if va < best_val:
            best_val, best_epoch = va, epoch
            torch.save(reg.state_dict(), ckpt_path)
        elif epoch - best_epoch > patience:
            print(f'early stopping at epoch {epoch}')
            reg.load_state_dict(torch.load(ckpt_path))
            break
    return reg, hist


def _dataset_cls(dataset_cls):
    if dataset_cls is not None:
        return dataset_cls
    try:
        from dataset_loader import EEGDatasetSimdata       # the reference's own loader, if the caller has it on sys.path
        return EEGDatasetSimdata
    except Exception as e:
        raise L.EegclipError("regression evaluation needs an EEGDatasetSimdata-style generator class: pass dataset_cls=... "
                             "(dataset plumbing is outside this package)") from e


def _files_of(files, subs):
    return [x for x in files if os.path.basename(x).split("_")[0] in subs]


def _audio_of(eeg_files, audio_files):
    stimuli = {os.path.basename(x).split("-audio-")[-1].split('_eeg')[0] for x in eeg_files}
    return [x for x in audio_files if os.path.basename(x).split("_-_")[0] in stimuli]


def evaluate_model_do_regression_sub_specific(model, train_files, val_files, test_files, train_files_audio, val_files_audio,
                                              test_files_audio, device, result_folder, regress_to=['env', 'mel'], window_length=5,
                                              fs=64, dataset_cls=None):
    """train_clip_helper_functions.py:443-764: one regression head per subject on frozen EEG-tower embeddings; returns
    {subject: test PearsonLoss}."""
    Dataset = _dataset_cls(dataset_cls)
    os.makedirs(os.path.join(result_folder, 'sub_specific'), exist_ok=True)
    model.eval()
    evaluation = {}
    all_subs = list(set(os.path.basename(x).split("_")[0] for x in train_files))
    print(f'number of subjects {len(all_subs)}')
    n = window_length * fs
    for sub in all_subs:
        try:
            print(f'subject {sub}')
            tr, va, te = _files_of(train_files, [sub]), _files_of(val_files, [sub]), _files_of(test_files, [sub])
            if not tr:
                print(f'subject {sub} has no train files')
                continue
            if not te:
                print(f'subject {sub} has no test files')
                continue
            if not va:
                print(f'subject {sub} has no val files')
                va = te
            a_tr, a_va, a_te = _audio_of(tr, train_files_audio), _audio_of(va, val_files_audio), _audio_of(te, test_files_audio)
            x_tr, y_tr, _ = regression_embeddings(model, Dataset(tr, a_tr, n, n, batch_size=128), device)
            x_va, y_va, _ = regression_embeddings(model, Dataset(va, a_va, n, n, batch_size=128), device)
            ckpt = os.path.join(result_folder, 'sub_specific', f'regression_model_{sub}.pth')
            reg, _ = fit_regression_head(x_tr, y_tr, x_va, y_va, device, ckpt, os.path.join(result_folder, 'loss_regression.txt'))
            x_te, y_te, _ = regression_embeddings(model, Dataset(te, a_te, n, n, batch_size=128), device)
            with torch.no_grad():
                evaluation[sub] = PearsonLoss()(reg(x_te), y_te).item()
            print(f'evaluation for subject {sub} is {evaluation[sub]}')
            with open(os.path.join(result_folder, 'evaluation_regression.json'), 'w') as f:
                json.dump(evaluation, f)
        except Exception as e:   # the reference logs and moves on to the next subject (:757-761)
            print(f'error with subject {sub}')
            printf(f'error with subject {sub}', os.path.join(result_folder, 'error_regression.txt'))
            printf(str(e), os.path.join(result_folder, 'error_regression.txt'))
            continue
    return evaluation


def evaluate_model_do_regression_sub_independent(model, train_files, val_files, test_files, train_files_audio, val_files_audio,
                                                 test_files_audio, device, result_folder, regress_to='env', window_length=5, fs=64,
                                                 dataset_cls=None):
    """train_clip_helper_functions.py:767-1103: one regression head over all subjects, evaluated per test subject."""
    Dataset = _dataset_cls(dataset_cls)
    os.makedirs(result_folder, exist_ok=True)
    model.eval()
    evaluation = {}
    n = window_length * fs
    a_tr, a_va = _audio_of(train_files, train_files_audio), _audio_of(val_files, val_files_audio)
    x_tr, y_tr, _ = regression_embeddings(model, Dataset(train_files, a_tr, n, n, batch_size=128), device)
    x_va, y_va, _ = regression_embeddings(model, Dataset(val_files, a_va, n, n, batch_size=128), device)
    ckpt = os.path.join(result_folder, 'regression_model_general_env.pth')
    reg, _ = fit_regression_head(x_tr, y_tr, x_va, y_va, device, ckpt, os.path.join(result_folder, 'loss_regression_general_env.txt'),
                                 val_batched=False)
    for sub in sorted(set(os.path.basename(x).split("_")[0] for x in test_files)):
        te = _files_of(test_files, [sub])
        x_te, y_te, _ = regression_embeddings(model, Dataset(te, _audio_of(te, test_files_audio), n, n, batch_size=128), device)
        with torch.no_grad():
            evaluation[sub] = PearsonLoss()(reg(x_te), y_te).item()
        print(f'evaluation for subject {sub} is {evaluation[sub]}')
    with open(os.path.join(result_folder, 'evaluation_regression_general_env.json'), 'w') as f:
        json.dump(evaluation, f)
    return evaluation
