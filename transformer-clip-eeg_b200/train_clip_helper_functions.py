"""Mirror of the match-mismatch scoring entry point of the reference ``train_clip_helper_functions.py``.

``evaluate_model_challenge_2023_mm`` keeps the reference signature, on-disk fixture format and return
values (train_clip_helper_functions.py:51-205) but scores on the GPU kernels:
  * the two candidate similarities per window are row-dots (eegclip_mm_rowdots) instead of the
    reference's N x N matmul followed by ``torch.diag`` (:159-160);
  * the all-stimuli logits are one similarity GEMM (eegclip_mm_bank_logits), top-k as in :187.
File discovery / JSON / pickle handling is host plumbing and follows the reference's layout
(:56-101,121-140).  The regression evaluations (:208-1103) are out of scope.
"""
import glob
import json
import os
import pickle

import numpy as np
import torch

from . import _lib as L


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
    import ctypes
    eeg_emb, bank_emb = L.f32c(eeg_emb), L.f32c(bank_emb)
    N, D = eeg_emb.shape
    M = bank_emb.shape[0]
    out = torch.empty(N, M, dtype=torch.float32, device=eeg_emb.device)
    nb = ctypes.c_size_t()
    L.call("eegclip_mm_bank_workspace", N, M, D, ctypes.byref(nb))
    scratch = torch.empty(max(nb.value, 16), dtype=torch.uint8, device=eeg_emb.device)
    L.call("eegclip_mm_bank_logits", L.ptr(eeg_emb), L.ptr(bank_emb), L.ptr(out), N, M, D, L.default_math(), L.ptr(scratch), L.stream())
    return out


def bank_topk(eeg_emb, bank_emb, k, group=None, logits_fn=None):
    """Top-k stimuli of the bank for every EEG window (train_clip_helper_functions.py:182-187), optionally with the bank
    SHARDED over a process group (BASELINE config 4): every rank holds all N windows and its own slice of the M
    candidates (``bank_emb`` is the local slice, slices concatenated in rank order form the bank), scores locally, keeps
    its local top-k, and the k*world candidates are all-gathered and merged.  Returns (values (N,k), global indices (N,k)),
    identical on all ranks and identical to a single-process top-k over the whole bank (ties aside)."""
    import torch.distributed as dist
    logits_fn = logits_fn or bank_logits
    world = dist.get_world_size(group) if (group is not None and dist.is_initialized()) else 1
    logits = logits_fn(eeg_emb, bank_emb)
    m_loc = logits.shape[1]
    kk = min(k, m_loc)
    vals, idx = torch.topk(logits, k=kk, dim=1)
    if world == 1:
        return vals, idx
    rank = dist.get_rank(group)
    sizes = torch.zeros(world, dtype=torch.int64, device=logits.device)
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
    cat_v, cat_i = torch.cat(all_v, dim=1), torch.cat(all_i, dim=1)
    kout = min(k, int(sizes.sum()))
    top_v, pos = torch.topk(cat_v, k=kout, dim=1)
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
        eeg = (eeg - np.mean(eeg, axis=(0, 1), keepdims=True)) / np.std(eeg, axis=(0, 1), keepdims=True)   # per-subject MVN (:136)
        labels = [labels_all[k] for k in ids]
        correct_keys = [mapping[k][labels_all[k] + 1].split('.')[0] for k in ids]
        with torch.no_grad():
            e = l2_normalize_rows(torch.flatten(model.eegModel(torch.from_numpy(eeg).to(device, dtype=torch.float)), start_dim=1))
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
            top = torch.topk(logits, k=maxtop, dim=1).indices.cpu().numpy()
            lab = np.repeat(idx.to(torch.int).cpu().numpy().astype(np.int32), maxtop).reshape(len(ids), -1)
            correct_top = np.mean(np.cumsum(np.equal(lab, top), axis=1), axis=0)
            evaluation_top_x[sub] = correct_top.tolist()
            evaluation_top_x_with_logits[sub] = {'logits': logits.tolist(), 'correct_keys_idx': idx.tolist(),
                                                 'correct_top': correct_top.tolist()}
            print(f"evaluation mm top x: {sub} : top1 {evaluation_top_x[sub][0] * 100}, top10: {evaluation_top_x[sub][9] * 100}")
    return evaluation, evaluation_with_logits, evaluation_top_x, evaluation_top_x_with_logits
