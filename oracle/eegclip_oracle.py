"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference EEG-CLIP hot path.

Plain functional torch (CPU, fp32 or fp64 depending on the dtype of what it is
given).  Every function cites the reference lines it restates
(/root/reference/...).  Gradients come from autograd over this restatement.

Pinned by tests/test_oracle_golden.py against tests/golden/*.json, which hold
outputs of the unmodified reference modules (oracle/make_golden.py).  The
reference itself ships no golden vectors.

Dropout: the reference draws nn.Dropout masks from torch's generator, which no
other implementation can reproduce; in train mode this oracle (and the patched
reference in make_golden.py) uses the Philox masks of oracle/philox_ref.py,
i.e. the same masks the CUDA kernels regenerate from (seed, layer, site, index).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import philox_ref

# dropout site ids (must match csrc/common.cuh)
SITE_CONV, SITE_ATTN, SITE_PROJ, SITE_FFN_HID, SITE_FFN_OUT = 0, 1, 2, 3, 4


class Drop:
    """Dropout policy: eval (identity) or train with Philox masks."""

    def __init__(self, train=False, seed=0, native=False):
        self.train = bool(train)
        self.seed = int(seed)
        self.native = bool(native)   # torch's own dropout (what the reference executes): used for CPU timing only

    def __call__(self, x, p, layer, site, order=None):
        """x: tensor; mask index = flat index of x (after ``order`` permutation, if given)."""
        if not self.train or p <= 0.0:
            return x
        if self.native:
            return F.dropout(x, p, True)
        if order is not None:
            xs = x.permute(*order)
        else:
            xs = x
        keep = philox_ref.keep_mask(xs.numel(), self.seed, philox_ref.stream_id(layer, site), p)
        m = torch.from_numpy(keep.reshape(tuple(xs.shape))).to(x.dtype) / (1.0 - p)
        if order is not None:
            inv = [0] * len(order)
            for i, o in enumerate(order):
                inv[o] = i
            m = m.permute(*inv)
        return x * m


EVAL = Drop(False)


def mha(sd, pre, x, drop=EVAL, layer=0, heads=8, p=0.5):
    """MultiHeadAttention.forward -- clip_model.py:30-45 (scale = sqrt(emb_size), :39)."""
    B, T, E = x.shape
    d = E // heads

    def proj(n):
        return F.linear(x, sd[pre + n + ".weight"], sd[pre + n + ".bias"]).view(B, T, heads, d).permute(0, 2, 1, 3)

    q, k, v = proj("queries"), proj("keys"), proj("values")
    energy = torch.einsum("bhqd,bhkd->bhqk", q, k)
    att = torch.softmax(energy / math.sqrt(E), dim=-1)
    att = drop(att, p, layer, SITE_ATTN)
    out = torch.einsum("bhal,bhlv->bhav", att, v).permute(0, 2, 1, 3).reshape(B, T, E)
    return F.linear(out, sd[pre + "projection.weight"], sd[pre + "projection.bias"])


def transformer_block(sd, pre, x, drop=EVAL, layer=0, p=0.5, p_ffn=0.5):
    """TransformerEncoderBlock + ResidualAdd -- clip_model.py:75-94, 48-57, 60-67."""
    E = x.shape[-1]
    h = F.layer_norm(x, (E,), sd[pre + "0.fn.0.weight"], sd[pre + "0.fn.0.bias"], 1e-5)
    a = mha(sd, pre + "0.fn.1.", h, drop, layer, p=p)
    x = x + drop(a, p, layer, SITE_PROJ)
    h = F.layer_norm(x, (E,), sd[pre + "1.fn.0.weight"], sd[pre + "1.fn.0.bias"], 1e-5)
    f = F.gelu(F.linear(h, sd[pre + "1.fn.1.0.weight"], sd[pre + "1.fn.1.0.bias"]))
    f = drop(f, p_ffn, layer, SITE_FFN_HID)
    f = F.linear(f, sd[pre + "1.fn.1.3.weight"], sd[pre + "1.fn.1.3.bias"])
    return x + drop(f, p, layer, SITE_FFN_OUT)


def conv1d_same(x, w, b):
    """nn.Conv1d(padding='same') on (B,C,T): left pad (k-1)//2, right pad k-1-left (k=64 -> 31/32)."""
    k = w.shape[-1]
    left = (k - 1) // 2
    return F.conv1d(F.pad(x, (left, k - 1 - left)), w, b)


def basic_block(sd, pre, x, drop=EVAL, layer=0, p=0.2, act="gelu"):
    """BasicBlock.forward on (B,C,T) -- clip_model.py:234-249 (conv, dropout, LN([C,T]), GELU)."""
    y = conv1d_same(x, sd[pre + "conv.weight"], sd[pre + "conv.bias"])
    y = drop(y, p, layer, SITE_CONV, order=(0, 2, 1))  # mask indexed in (B,T,C) order
    g, b = sd[pre + "normalization.weight"], sd[pre + "normalization.bias"]
    y = F.layer_norm(y, tuple(g.shape), g, b, 1e-5)
    return F.gelu(y) if act == "gelu" else F.leaky_relu(y, 0.01)


def eeg_conformer_interleaved(sd, x, depth, drop=EVAL, pre="", p_conv=0.2):
    """EEGConformerInterleaved.forward -- clip_model.py:445-474. x: (B,T,64) -> (B,T,latent)."""
    x = x.permute(0, 2, 1)
    x = F.conv1d(x, sd[pre + "eeg_spatial_mapping.weight"], sd[pre + "eeg_spatial_mapping.bias"])
    eeg = x
    eeg_t = eeg.permute(0, 2, 1)
    for i in range(depth):
        if i != 0:
            x = x.permute(0, 2, 1)
        x = basic_block(sd, pre + f"conv_{i}.", x + eeg, drop, i, p_conv)
        x = x.permute(0, 2, 1)
        inp = x if i == depth - 1 else x + eeg_t
        x = transformer_block(sd, pre + f"conformer_{i}.0.", inp, drop, i)
    return F.linear(x, sd[pre + "final_layer.weight"], sd[pre + "final_layer.bias"])


def eeg_conformer(sd, x, n_conv, depth, drop=EVAL, pre="", p_conv=0.2):
    """EEGConformer.forward (sequential variant) -- clip_model.py:373-398."""
    x = x.permute(0, 2, 1)
    x = F.conv1d(x, sd[pre + "eeg_spatial_mapping.weight"], sd[pre + "eeg_spatial_mapping.bias"])
    eeg = x
    for i in range(n_conv):
        x = basic_block(sd, pre + f"conv_{i}.", x if i == n_conv - 1 else x + eeg, drop, i, p_conv)
    x = x.permute(0, 2, 1)
    for i in range(depth):
        x = transformer_block(sd, pre + f"transformerEncoder.{i}.", x, drop, n_conv + i)
    return F.linear(x, sd[pre + "final_layer.weight"], sd[pre + "final_layer.bias"])


def speech_small_conv(sd, x, drop=EVAL, pre="", p=0.4):
    """SpeechSmallConv.forward -- clip_model.py:224-232. x: (B,T,F) -> (B,T,out)."""
    y = conv1d_same(x.permute(0, 2, 1), sd[pre + "speech_spatial_mapping.weight"], sd[pre + "speech_spatial_mapping.bias"])
    y = drop(y, p, 0, SITE_CONV, order=(0, 2, 1))
    g, b = sd[pre + "layernorm.weight"], sd[pre + "layernorm.bias"]
    y = F.leaky_relu(F.layer_norm(y, tuple(g.shape), g, b, 1e-5), 0.01)
    return y.permute(0, 2, 1)


def _lstm_dir(x, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of nn.LSTM(batch_first) with zero initial state; gate order i,f,g,o."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    outs = [None] * T
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        g = F.linear(x[:, t], w_ih, b_ih) + F.linear(h, w_hh, b_hh)
        i, f, gg, o = g.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs[t] = h
    return torch.stack(outs, dim=1)


def bilstm(sd, pre, x):
    f = _lstm_dir(x, sd[pre + "weight_ih_l0"], sd[pre + "weight_hh_l0"], sd[pre + "bias_ih_l0"], sd[pre + "bias_hh_l0"], False)
    r = _lstm_dir(x, sd[pre + "weight_ih_l0_reverse"], sd[pre + "weight_hh_l0_reverse"],
                  sd[pre + "bias_ih_l0_reverse"], sd[pre + "bias_hh_l0_reverse"], True)
    return torch.cat([f, r], dim=2)


_LSTM_CACHE = {}


def bilstm_fast(sd, pre, x):
    """Same bi-LSTM through torch's fused nn.LSTM kernel (what the reference runs, clip_model.py:267-268,322-323)."""
    w_ih = sd[pre + "weight_ih_l0"]
    hid, inp = w_ih.shape[0] // 4, w_ih.shape[1]
    mod = _LSTM_CACHE.get((inp, hid))
    if mod is None:
        mod = _LSTM_CACHE[(inp, hid)] = torch.nn.LSTM(inp, hid, batch_first=True, bidirectional=True)
    names = [n for n, _ in mod.named_parameters()]
    out, _ = torch.func.functional_call(mod, {n: sd[pre + n] for n in names}, (x,))
    return out


def eeg_conv_lstm(sd, x, n_blocks=1, drop=EVAL, pre="", p=0.4, fast_lstm=False):
    """EEGConvLSTM.forward -- clip_model.py:302-325 (default speech tower)."""
    x = F.conv1d(x.permute(0, 2, 1), sd[pre + "eeg_spatial_mapping.weight"], sd[pre + "eeg_spatial_mapping.bias"])
    eeg = x
    for i in range(n_blocks):
        x = basic_block(sd, pre + f"conv_{i}.", x if i == n_blocks - 1 else x + eeg, drop, i, p)
    x = x.permute(0, 2, 1)
    lstm = bilstm_fast if fast_lstm else bilstm
    x = lstm(sd, pre + "speech_lstm1.", x)
    return lstm(sd, pre + "speech_lstm2.", x)


def l2_normalize(x):
    """F.normalize(p=2, dim=1, eps=1e-12) -- clip_model.py:675-676, 913-914."""
    return x / x.norm(dim=1, keepdim=True).clamp_min(1e-12)


def symmetric_infonce(E, S, tau):
    """CLIP.forward tail -- clip_model.py:675-693.  E,S: raw flattened (B,D); tau: log-scale."""
    En, Sn = l2_normalize(E), l2_normalize(S)
    logits = (Sn @ En.T) * torch.exp(tau)
    tgt = torch.arange(logits.shape[0])
    return (F.cross_entropy(logits, tgt) + F.cross_entropy(logits.T, tgt)) / 2.0


def infonce_closed_form(En, Sn, tau):
    """Same loss from already-normalised embeddings via LSE (what the fused kernel computes)."""
    L = (Sn @ En.T) * torch.exp(tau)
    d = torch.diagonal(L)
    return ((torch.logsumexp(L, 1) - d).mean() + (torch.logsumexp(L, 0) - d).mean()) / 2.0


def memory_bank_update(memory, idx, data, momentum=0.9):
    """memoryBank.forward -- clip_model.py:731-745. Mutates ``memory`` in place, returns old rows."""
    old = memory.index_select(0, idx.view(-1)).detach()
    with torch.no_grad():
        memory.index_copy_(0, idx, old * momentum + data.detach() * (1 - momentum))
    return old


def clip_sim_no_latent_proj(eeg_feat, speech_feat, ids, memory, tau, tau_eeg, lam_clip=1.0, lam_avg=0.0, momentum=0.9):
    """CLIPSimNoLatentProj.forward after the towers -- clip_model.py:892-944.

    eeg_feat/speech_feat: tower outputs (B,T,latent).  Returns (loss_ce, avg_loss, loss_total).
    """
    if speech_feat.shape[1] > speech_feat.shape[2]:
        speech_feat = speech_feat.transpose(1, 2)
    if eeg_feat.shape[1] > eeg_feat.shape[2]:
        eeg_feat = eeg_feat.transpose(1, 2)
    E = l2_normalize(eeg_feat.flatten(1))
    S = l2_normalize(speech_feat.flatten(1))
    avg = l2_normalize(memory_bank_update(memory, ids, E, momentum))
    tgt = torch.arange(E.shape[0])
    logits = (S @ E.T) * torch.exp(tau)
    loss_ce = (F.cross_entropy(logits, tgt) + F.cross_entropy(logits.T, tgt)) / 2.0
    avg_loss = F.cross_entropy((avg @ E.T) * torch.exp(tau_eeg), tgt)
    return loss_ce, avg_loss, lam_clip * loss_ce + lam_avg * avg_loss


def sharded_infonce(En, Sn, tau, world):
    """Single-process simulation of the R-rank sharded InfoNCE of SURVEY §8(e).

    Returns (loss, dEn, dSn, dtau) assembled from per-rank row/column blocks; must
    equal autograd through infonce_closed_form on the full batch.
    """
    B = En.shape[0]
    b = B // world
    scale = torch.exp(tau)
    lse_r, lse_c, diag = [], [], []
    for r in range(world):
        sl = slice(r * b, (r + 1) * b)
        Lr = (Sn[sl] @ En.T) * scale            # rows of this rank
        Lc = (Sn @ En[sl].T) * scale            # columns of this rank
        lse_r.append(torch.logsumexp(Lr, 1))
        lse_c.append(torch.logsumexp(Lc, 0))
        diag.append((Sn[sl] * En[sl]).sum(1) * scale)
    lse_r, lse_c, diag = torch.cat(lse_r), torch.cat(lse_c), torch.cat(diag)
    loss = ((lse_r - diag).mean() + (lse_c - diag).mean()) / 2.0
    dS, dE, dtau = [], [], 0.0
    for r in range(world):
        sl = slice(r * b, (r + 1) * b)
        eye = torch.zeros(b, B, dtype=En.dtype)
        eye[torch.arange(b), torch.arange(r * b, (r + 1) * b)] = 1.0
        Lr = (Sn[sl] @ En.T) * scale
        Gr = (torch.exp(Lr - lse_r[sl, None]) + torch.exp(Lr - lse_c[None, :]) - 2 * eye) / (2 * B)
        Lc = (Sn @ En[sl].T) * scale
        Gc = (torch.exp(Lc - lse_r[:, None]) + torch.exp(Lc - lse_c[None, sl]) - 2 * eye.T) / (2 * B)
        dS.append(scale * Gr @ En)
        dE.append(scale * Gc.T @ Sn)
        dtau = dtau + (Gr * Lr).sum()
    return loss, torch.cat(dE), torch.cat(dS), dtau


def mm_decisions(eeg_emb, cand_emb, bank_emb, k=100):
    """Match-mismatch scoring core -- train_clip_helper_functions.py:153-187.

    eeg_emb (N,D) normalised; cand_emb (N,K,D); bank_emb (M,D).
    Returns (scores (K,N), argmax (N,), topk indices (N,min(k,M))).
    """
    scores = torch.einsum("nd,nkd->kn", eeg_emb, cand_emb)
    choice = torch.argmax(scores, dim=0)
    logits = eeg_emb @ bank_emb.T
    top = torch.topk(logits, k=min(k, logits.shape[1]), dim=1).indices
    return scores, choice, top


def mvn_per_subject(eeg):
    """Per-channel mean/std over (segments,time) -- train_clip_helper_functions.py:136 (np.std, ddof=0)."""
    e = np.asarray(eeg)
    return (e - e.mean(axis=(0, 1), keepdims=True)) / e.std(axis=(0, 1), keepdims=True)


class LeakyTap:
    """LeakyReLU(0.01) whose branch decisions can be recorded or forced (tests/test_gpu_parity_r2.py: the CUDA path's own
    branch masks are replayed through the fp64 oracle, which isolates "an element sat within rounding distance of the kink"
    from genuine arithmetic error).  ``forced``: list of bool tensors consumed in call order; ``seen``: (pre, mask) log."""

    def __init__(self, forced=None):
        self.forced, self.seen, self.i = forced, [], 0

    def __call__(self, pre):
        own = pre > 0
        mask = own if self.forced is None else self.forced[self.i].to(pre.device)
        self.i += 1
        self.seen.append((pre.detach(), own, mask))
        return torch.where(mask, pre, 0.01 * pre)


def _vlaai_stack(sd, x, leaky):
    pre = "sequentialConvStack.0."
    x = F.conv1d(x, sd[pre + "eeg.weight"], sd[pre + "eeg.bias"])
    for j in range(5):
        x = conv1d_same(x, sd[pre + f"conv_layers.{3 * j}.weight"], sd[pre + f"conv_layers.{3 * j}.bias"])
        g, b = sd[pre + f"conv_layers.{3 * j + 1}.weight"], sd[pre + f"conv_layers.{3 * j + 1}.bias"]
        x = leaky(F.layer_norm(x, tuple(g.shape), g, b, 1e-5))
    x = F.conv1d(x, sd["sequentialConvStack.1.weight"], sd["sequentialConvStack.1.bias"])
    x = conv1d_same(x, sd["sequentialConvStack.2.conv1d.weight"], sd["sequentialConvStack.2.conv1d.bias"])
    g, b = sd["sequentialConvStack.2.normalization_fn.weight"], sd["sequentialConvStack.2.normalization_fn.bias"]
    return leaky(F.layer_norm(x, tuple(g.shape), g, b, 1e-5))


def vlaai(sd, x, nb_blocks=4, leaky=None):
    """VLAAI.forward -- vlaai.py:109-134. x: (B,T,64) -> (B,64,T); the stack weights are shared."""
    leaky = leaky or (lambda t: F.leaky_relu(t, 0.01))
    x = x.transpose(1, 2)
    eeg = x
    x = F.conv1d(x, sd["eeg.weight"], sd["eeg.bias"])
    for i in range(nb_blocks):
        x = _vlaai_stack(sd, x if (i == 0 or i == nb_blocks - 1) else x + eeg, leaky)
    return F.conv1d(x, sd["final_linear.weight"], sd["final_linear.bias"])


def adamw_step(p, g, m, v, step, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8, wd=0.01):
    """torch.optim.AdamW single-tensor update (train_clip_final.py:409-413 uses torch defaults)."""
    p = p * (1 - lr * wd)
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * m / denom
    return p, m, v


def adam_step(p, g, m, v, vmax, step, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8, wd=0.0, decoupled=False, amsgrad=False):
    """torch.optim.Adam / AdamW single-tensor update incl. amsgrad (train_clip_final.py:403-413; helpers :626)."""
    if decoupled:
        p = p * (1 - lr * wd)
    else:
        g = g + wd * p
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    if amsgrad:
        vmax = torch.maximum(vmax, v)
    denom = (vmax if amsgrad else v).sqrt() / math.sqrt(1 - b2 ** step) + eps
    return p - (lr / (1 - b1 ** step)) * m / denom, m, v, vmax


# ---- stand-alone transformer sub-modules (clip_model.py:30-67) -------------------------------------------------------------
def feed_forward(sd, pre, x, drop=EVAL, layer=0, p=0.5):
    """FeedForwardBlock -- clip_model.py:60-67."""
    f = drop(F.gelu(F.linear(x, sd[pre + "0.weight"], sd[pre + "0.bias"])), p, layer, SITE_FFN_HID)
    return F.linear(f, sd[pre + "3.weight"], sd[pre + "3.bias"])


def residual_ln_mha(sd, pre, x, drop=EVAL, layer=0, p=0.5):
    """ResidualAdd(Sequential(LayerNorm, MultiHeadAttention, Dropout)) -- clip_model.py:48-57, 83-87."""
    h = F.layer_norm(x, (x.shape[-1],), sd[pre + "fn.0.weight"], sd[pre + "fn.0.bias"], 1e-5)
    return x + drop(mha(sd, pre + "fn.1.", h, drop, layer, p=p), p, layer, SITE_PROJ)


# ---- non-default loss wrappers after the towers (clip_model.py:747-810, 948-1168, 1174-1450) --------------------------------
def _sym_ce(logits):
    tgt = torch.arange(logits.shape[0])
    return (F.cross_entropy(logits, tgt) + F.cross_entropy(logits.T, tgt)) / 2.0


def clip_sim(ef, sf, ids, memory, w_eeg, w_speech, tau, lam_clip=1.0, lam_avg=1.0, momentum=0.9):
    """CLIPSim.forward -- clip_model.py:773-808: bias-free latent projections, MSE against the normalised bank average."""
    E = l2_normalize(F.linear(ef.flatten(1), w_eeg))
    S = l2_normalize(F.linear(sf.flatten(1), w_speech))
    avg = l2_normalize(memory_bank_update(memory, ids, E, momentum))
    loss_ce = _sym_ce((S @ E.T) * torch.exp(tau))
    avg_loss = F.mse_loss(avg, E)
    return loss_ce, avg_loss, lam_clip * loss_ce + lam_avg * avg_loss


def clip_multiple_positives(ef, sf, tau, lam_clip=1.0, lam_avg=1.0, adapted=False):
    """CLIPSimMultiplePositives[Adapted].forward -- clip_model.py:1017-1078 / 1098-1168; helpers :1478-1494."""
    E, S = l2_normalize(ef.flatten(1)), l2_normalize(sf.flatten(1))
    logits = (S @ E.T) * torch.exp(tau)
    B = logits.shape[0]
    tgt = torch.arange(B)
    eeg_loss = F.cross_entropy(logits.T, torch.cat((tgt,) * (logits.shape[1] // B)))
    x = logits.reshape(B, -1, B)
    if adapted:
        loss_ce = (F.cross_entropy(x.sum(1), tgt) + eeg_loss) / 2.0
        return loss_ce, loss_ce, lam_clip * loss_ce
    lsm = x.exp().sum(-2).log() - x.exp().sum(-2).sum(-1).log().unsqueeze(-1)
    loss_ce = (F.nll_loss(lsm, tgt) + eeg_loss) / 2.0
    sim = F.nll_loss(x.sum(-2), tgt)
    return loss_ce, sim, lam_clip * loss_ce + lam_avg * sim


def _log_gauss(x, mu, logvar):
    return -0.5 * (math.log(2 * math.pi) + logvar + (x - mu) ** 2 / math.exp(logvar))


def kld_lower_bound(mu2, z_mu, z_logvar):
    """clip_model.py:1226-1239: prior N(mu2, 0.5^2) on z, N(0, 1) on mu2."""
    q_logvar = math.log(0.5 ** 2)
    log_pmu2 = _log_gauss(mu2, 0.0, 0.0).mean(1)
    kld_z2 = (-0.5 * (1 + z_logvar - q_logvar - ((z_mu - mu2) ** 2 + z_logvar.exp()) / math.exp(q_logvar))).mean(1)
    return log_pmu2, kld_z2, (-log_pmu2 + kld_z2).mean(0)


def clip_kld(ef, sf, ids, sd, tau, lam_clip=1.0, lam_lb=1.0):
    """CLIPKLDNoLatentProj.forward -- clip_model.py:1204-1268."""
    E, S = ef.flatten(1), sf.flatten(1)
    mu2 = sd["mu_eeg_lookup.weight"][ids]
    z_mu = F.linear(E, sd["eeg_mu_linear.weight"], sd["eeg_mu_linear.bias"])
    z_lv = F.linear(E, sd["eeg_logvar_linear.weight"], sd["eeg_logvar_linear.bias"])
    log_pmu2, kld_z2, lb = kld_lower_bound(mu2, z_mu, z_lv)
    loss_ce = symmetric_infonce(E, S, tau)
    return lam_clip * loss_ce + lam_lb * lb, loss_ce, log_pmu2.mean(), kld_z2.mean()


def _proj_head_linear(sd, pre, x):
    """ProjectionHeadLinear -- clip_model.py:1303-1320."""
    h = F.leaky_relu(F.linear(x, sd[pre + "projection.weight"], sd[pre + "projection.bias"]), 0.01)
    return F.linear(h, sd[pre + "last_linear.weight"], sd[pre + "last_linear.bias"])


def clip_kld_latent_proj(ef, sf, ids, sd, tau, lam_clip=1.0, lam_lb=1.0):
    """CLIPKLDWithLatentProj.forward (linear heads) -- clip_model.py:1362-1440."""
    E, S = ef.flatten(1), sf.flatten(1)
    z_lv, z_mu = _proj_head_linear(sd, "eeg_logvar_linear.", E), _proj_head_linear(sd, "eeg_mu_linear.", E)
    Sp = _proj_head_linear(sd, "speech_latent_projection.", S)
    log_pmu2, kld_z2, lb = kld_lower_bound(sd["mu_eeg_lookup.weight"][ids], z_mu, z_lv)
    loss_ce = symmetric_infonce(z_mu, Sp, tau)
    return lam_clip * loss_ce + lam_lb * lb, loss_ce, log_pmu2.mean(), kld_z2.mean()


def clip_no_contrastive(ef, sf):
    """CLIPNoContrastiveLearning.forward -- clip_model.py:959-993."""
    if sf.shape[1] > sf.shape[2]:
        sf = sf.transpose(1, 2)
    if ef.shape[1] > ef.shape[2]:
        ef = ef.transpose(1, 2)
    L = l2_normalize(sf.flatten(1)) @ l2_normalize(ef.flatten(1)).T
    match, mism = torch.diagonal(L)[:-1], torch.diagonal(L, offset=1)
    tgt = torch.stack([torch.ones_like(match), torch.zeros_like(mism)])
    return F.binary_cross_entropy_with_logits(torch.stack([match, mism]), tgt)


# ---- regression evaluation (train_clip_helper_functions.py:1107-1140) -------------------------------------------------------
def pearson_loss(x, y):
    """PearsonLoss.forward -- helpers:1111-1118.  (B,C,T) -> (C,)."""
    xc, yc = x - x.mean(2, keepdim=True), y - y.mean(2, keepdim=True)
    r = (xc * yc).sum(2) / (xc.norm(dim=2).clamp_min(1e-6) * yc.norm(dim=2).clamp_min(1e-6))
    return -r.mean(0)


def regression_model(w, b, x):
    """RegressionModel.forward -- helpers:1132-1140: Conv1d('same') + LeakyReLU on (B,Cin,T)."""
    return F.leaky_relu(conv1d_same(x, w, b), 0.01)


def to_dtype(sd, dtype):
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}


def grads_of(loss, sd):
    keys = [k for k, v in sd.items() if v.requires_grad]
    gs = torch.autograd.grad(loss, [sd[k] for k in keys], allow_unused=True)
    return {k: (g if g is not None else torch.zeros_like(sd[k])) for k, g in zip(keys, gs)}
