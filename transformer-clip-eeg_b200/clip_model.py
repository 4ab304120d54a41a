"""Drop-in mirror of the reference ``clip_model.py`` for the EEG-CLIP hot path, on B200 kernels.

Same class names, constructor / forward signatures, return arity and ``state_dict`` keys as
/root/reference/clip_model.py, so ``import clip_model`` with this directory on
``sys.path`` replaces the reference module for the path in scope:

    MultiHeadAttention, ResidualAdd, FeedForwardBlock, TransformerEncoderBlock, TransformerEncoder  (:19-99)
    BasicBlock (:234-249), EEGConformer (:327-398), EEGConformerInterleaved (:400-474)
    SpeechSmallConv (:204-232), EEGConvLSTM (:251-325)
    CLIP (:657-693), memoryBank (:697-745), CLIPSimNoLatentProj (:868-944)
    CLIPSim (:747-810), CLIPNoContrastiveLearning (:948-995), CLIPSimMultiplePositives[Adapted] (:1000-1168),
    CLIPKLDNoLatentProj (:1174-1279), CLIPKLDWithLatentProj (:1325-1450) and their helpers (:1473-1500)

The nested containers (nn.Sequential / nn.Linear / nn.LayerNorm / nn.Conv1d) only *hold* the
parameters under the reference's key names; every forward below goes through the C ABI of
libeegclip_b200.so (include/eegclip.h).  There is no eager/CPU fallback: CPU tensors raise.
"""
import ctypes
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L
from .parallel import _all_gather_rows, _world, infonce_loss, ce_rows_loss, l2_normalize

__all__ = [
    "MultiHeadAttention", "ResidualAdd", "FeedForwardBlock", "TransformerEncoderBlock", "TransformerEncoder",
    "BasicBlock", "EEGConformer", "EEGConformerInterleaved", "SpeechSmallConv", "EEGConvLSTM", "CLIP", "memoryBank",
    "CLIPSimNoLatentProj", "CLIPSim", "CLIPSimMultiplePositives", "CLIPSimMultiplePositivesAdapted", "CLIPKLDNoLatentProj",
    "CLIPKLDWithLatentProj", "CLIPNoContrastiveLearning", "ProjectionHeadLinear", "multiple_postives_loss", "simloss",
    "log_gauss", "kld", "LayerNorm", "Dropout",
]


def _bytes(n):
    return torch.empty(max(int(n), 16), dtype=torch.uint8, device="cuda")


def _flat_grads(params):
    """One contiguous gradient buffer + per-parameter views (so the backward zero-fills with one memset)."""
    sizes = [p.numel() for p in params]
    offs, o = [], 0
    for s in sizes:
        offs.append(o)
        o += (s + 3) // 4 * 4  # keep every view 16-byte aligned
    flat = torch.empty(o, dtype=torch.float32, device=params[0].device)
    views = [flat[a:a + s].view(p.shape) for a, s, p in zip(offs, sizes, params)]
    return flat, views


# ---------------------------------------------------------------------------------------------------
# autograd glue (each Function == one pair of C-ABI calls)
# ---------------------------------------------------------------------------------------------------
class _TowerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, desc, *params):
        x = L.f32c(x)
        ps = [L.f32c(p) for p in params]
        save_b, scr_b = ctypes.c_size_t(), ctypes.c_size_t()
        L.call("eegclip_tower_workspace", ctypes.byref(desc), ctypes.byref(save_b), ctypes.byref(scr_b))
        save, scratch = _bytes(save_b.value), _bytes(scr_b.value)
        out = torch.empty(desc.B, desc.T, desc.latent, dtype=torch.float32, device=x.device)
        tab = L.ptr_table(ps)
        L.call("eegclip_tower_forward", ctypes.byref(desc), tab.data_ptr(), L.ptr(x), L.ptr(out), L.ptr(save), L.ptr(scratch),
               L.stream())
        ctx.desc, ctx.save, ctx.scr_bytes, ctx.x, ctx.ps = desc, save, scr_b.value, x, ps
        return out

    @staticmethod
    def backward(ctx, dout):
        desc, ps, x = ctx.desc, ctx.ps, ctx.x
        dout = L.f32c(dout)
        sinks = L.claim_grad_sinks(ps)                 # zero-filled arena views of the optimizer, or None
        if sinks is not None:
            flat, views = None, sinks
        else:
            flat, views = _flat_grads(ps)
        scratch = _bytes(ctx.scr_bytes)
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        ptab, gtab = L.ptr_table(ps), L.ptr_table(views)
        L.call("eegclip_tower_backward", ctypes.byref(desc), ptab.data_ptr(), gtab.data_ptr(), L.ptr(flat),
               flat.numel() * 4 if flat is not None else 0, L.ptr(x), L.ptr(dout), L.ptr(dx), L.ptr(ctx.save), L.ptr(scratch), L.stream())
        ctx.save = None
        if sinks is not None:
            L.fire_sinks_written(sinks)     # data-parallel: this tower's slice of the gradient arena can be all-reduced from here on
            return (dx, None, *([None] * len(ps)))
        return (dx, None, *views)


class _XfBlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, desc, *params):
        z = L.f32c(z)
        ps = [L.f32c(p) for p in params]
        save_b, scr_b = ctypes.c_size_t(), ctypes.c_size_t()
        L.call("eegclip_xfblock_workspace", ctypes.byref(desc), ctypes.byref(save_b), ctypes.byref(scr_b))
        save, scratch = _bytes(save_b.value), _bytes(scr_b.value)
        out = torch.empty_like(z)
        tab = L.ptr_table(ps)
        L.call("eegclip_xfblock_forward", ctypes.byref(desc), tab.data_ptr(), L.ptr(z), L.ptr(out), L.ptr(save), L.ptr(scratch),
               L.stream())
        ctx.desc, ctx.save, ctx.scr_bytes, ctx.z, ctx.ps = desc, save, scr_b.value, z, ps
        return out

    @staticmethod
    def backward(ctx, dout):
        desc, ps, z = ctx.desc, ctx.ps, ctx.z
        dout = L.f32c(dout)
        sinks = L.claim_grad_sinks(ps)
        if sinks is not None:
            flat, views = None, sinks
        else:
            flat, views = _flat_grads(ps)
        scratch = _bytes(ctx.scr_bytes)
        dz = torch.empty_like(z)
        ptab, gtab = L.ptr_table(ps), L.ptr_table(views)
        L.call("eegclip_xfblock_backward", ctypes.byref(desc), ptab.data_ptr(), gtab.data_ptr(), L.ptr(flat),
               flat.numel() * 4 if flat is not None else 0, L.ptr(z), L.ptr(dout), L.ptr(dz), L.ptr(ctx.save), L.ptr(scratch), L.stream())
        ctx.save = None
        if sinks is not None:
            return (dz, None, *([None] * len(ps)))
        return (dz, None, *views)


class _ConvBlockFn(torch.autograd.Function):
    """x, skip: time-major (B,T,Cin); returns time-major (B,T,Cout)."""

    @staticmethod
    def forward(ctx, x, skip, desc, w, b, gamma, beta):
        x, w, b, gamma, beta = (L.f32c(t) for t in (x, w, b, gamma, beta))
        skip = L.f32c(skip) if skip is not None else None
        save_b, scr_b = ctypes.c_size_t(), ctypes.c_size_t()
        L.call("eegclip_convblock_workspace", ctypes.byref(desc), ctypes.byref(save_b), ctypes.byref(scr_b))
        save, scratch = _bytes(save_b.value), _bytes(scr_b.value)
        out = torch.empty(desc.B, desc.T, desc.Cout, dtype=torch.float32, device=x.device)
        L.call("eegclip_convblock_forward", ctypes.byref(desc), L.ptr(x), L.ptr(skip), L.ptr(w), L.ptr(b), L.ptr(gamma),
               L.ptr(beta), L.ptr(out), L.ptr(save), L.ptr(scratch), L.stream())
        ctx.desc, ctx.save, ctx.scr_bytes, ctx.t, ctx.bias = desc, save, scr_b.value, (x, skip, w, gamma, beta), b
        return out

    @staticmethod
    def backward(ctx, dout):
        desc = ctx.desc
        x, skip, w, gamma, beta = ctx.t
        dout = L.f32c(dout)
        scratch = _bytes(ctx.scr_bytes)
        dx = torch.empty_like(x)
        sinks = L.claim_grad_sinks([w, ctx.bias, gamma, beta]) if ctx.bias is not None else None
        if sinks is not None:
            dw, db, dg, dbe = sinks
        else:
            dw, dg, dbe = torch.empty_like(w), torch.empty_like(gamma), torch.empty_like(beta)
            db = torch.empty(desc.Cout, dtype=torch.float32, device=x.device)
        L.call("eegclip_convblock_backward", ctypes.byref(desc), L.ptr(x), L.ptr(skip), L.ptr(w), L.ptr(gamma), L.ptr(beta),
               L.ptr(dout), L.ptr(dx), L.ptr(dw), L.ptr(db), L.ptr(dg), L.ptr(dbe), L.ptr(ctx.save), L.ptr(scratch), L.stream())
        ctx.save = None
        if sinks is not None:
            return dx, (dx if skip is not None else None), None, None, None, None, None
        return dx, (dx if skip is not None else None), None, dw, db, dg, dbe


class _LinearFn(torch.autograd.Function):
    """Per-token linear (nn.Linear / 1x1 Conv1d) on (..., K) -> (..., N)."""

    @staticmethod
    def _scratch(M, N, K):
        nb = ctypes.c_size_t()
        L.call("eegclip_linear_workspace", M, N, K, ctypes.byref(nb))
        return _bytes(nb.value)

    @staticmethod
    def forward(ctx, x, w, b):
        x, w2 = L.f32c(x), L.f32c(w).reshape(w.shape[0], -1)
        b = L.f32c(b) if b is not None else None
        M, K, N = x.numel() // x.shape[-1], x.shape[-1], w2.shape[0]
        out = torch.empty(*x.shape[:-1], N, dtype=torch.float32, device=x.device)
        scratch = _LinearFn._scratch(M, N, K)
        L.call("eegclip_linear_forward", L.ptr(x), L.ptr(w2), L.ptr(b), L.ptr(out), M, N, K, L.default_math(), L.ptr(scratch),
               L.stream())
        ctx.t, ctx.wshape, ctx.has_b, ctx.math = (x, w2), w.shape, b is not None, L.default_math()
        ctx.w_orig, ctx.b_orig = L.f32c(w), b
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w2 = ctx.t
        dout = L.f32c(dout)
        M, K, N = x.numel() // x.shape[-1], x.shape[-1], w2.shape[0]
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        sinks = L.claim_grad_sinks([ctx.w_orig, ctx.b_orig]) if ctx.has_b else None
        if sinks is not None:
            dw, db = sinks[0].view(N, K), sinks[1]
        else:
            dw = torch.empty_like(w2)
            db = torch.empty(N, dtype=torch.float32, device=x.device) if ctx.has_b else None
        scratch = _LinearFn._scratch(M, N, K)
        L.call("eegclip_linear_backward", L.ptr(x), L.ptr(w2), L.ptr(dout), L.ptr(dx), L.ptr(dw), L.ptr(db), M, N, K,
               ctx.math, L.ptr(scratch), L.stream())
        if sinks is not None:
            return dx, None, None
        return dx, dw.view(ctx.wshape), db


_LSTM_ORDER = ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0",
               "weight_ih_l0_reverse", "weight_hh_l0_reverse", "bias_ih_l0_reverse", "bias_hh_l0_reverse")


class _BiLSTMFn(torch.autograd.Function):
    """nn.LSTM(batch_first=True, bidirectional=True), zero initial state: x (B,T,In) -> (B,T,2H) on the eegclip kernels."""

    @staticmethod
    def forward(ctx, x, desc, *params):
        x = L.f32c(x)
        ps = [L.f32c(p) for p in params]
        save_b, scr_b = ctypes.c_size_t(), ctypes.c_size_t()
        L.call("eegclip_bilstm_workspace", ctypes.byref(desc), ctypes.byref(save_b), ctypes.byref(scr_b))
        save, scratch = _bytes(save_b.value), _bytes(scr_b.value)
        out = torch.empty(desc.B, desc.T, 2 * desc.H, dtype=torch.float32, device=x.device)
        tab = L.ptr_table(ps)
        L.call("eegclip_bilstm_forward", ctypes.byref(desc), tab.data_ptr(), L.ptr(x), L.ptr(out), L.ptr(save), L.ptr(scratch), L.stream())
        ctx.desc, ctx.save, ctx.scr_bytes, ctx.x, ctx.ps = desc, save, scr_b.value, x, ps
        return out

    @staticmethod
    def backward(ctx, dout):
        desc, ps, x = ctx.desc, ctx.ps, ctx.x
        dout = L.f32c(dout)
        sinks = L.claim_grad_sinks(ps)
        grads = sinks if sinks is not None else [torch.empty_like(p) for p in ps]
        scratch = _bytes(ctx.scr_bytes)
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        ptab, gtab = L.ptr_table(ps), L.ptr_table(grads)
        L.call("eegclip_bilstm_backward", ctypes.byref(desc), ptab.data_ptr(), gtab.data_ptr(), L.ptr(x), L.ptr(dout), L.ptr(dx),
               L.ptr(ctx.save), L.ptr(scratch), L.stream())
        ctx.save = None
        if sinks is not None:
            return (dx, None, *([None] * len(ps)))
        return (dx, None, *grads)


def _bilstm(mod, x):
    """Run an nn.LSTM parameter container on the eegclip recurrence kernels (no second backend: uncovered shapes raise)."""
    B, T, In = x.shape
    d = L.BiLstmDesc(B=B, T=T, In=In, H=mod.hidden_size, math=L.default_math())
    if not (mod.bidirectional and mod.num_layers == 1 and mod.batch_first and L.load().eegclip_bilstm_supported(ctypes.byref(d))):
        raise L.EegclipError(f"bi-LSTM shape (input {In}, hidden {mod.hidden_size}, layers {mod.num_layers}, bidirectional "
                             f"{mod.bidirectional}) is not covered by the eegclip recurrence kernels (H=128 with In 64/128, H=4 with In "
                             "64..256; tensor-core math) and this path has no library fallback")
    return _BiLSTMFn.apply(x, d, *[getattr(mod, n) for n in _LSTM_ORDER])


def _require_cuda(x, who):
    if not x.is_cuda:
        raise L.EegclipError(f"{who}: input is on {x.device}; this implementation runs on CUDA (sm_100a) only")


def _check_ln_shape(norm, C, T, who):
    """LayerNorm([C, time_dimension]) affines are read as a (C,T) array by the kernels: the window length must be the one the
    module was built with (the reference raises a shape error from F.layer_norm in the same situation)."""
    if tuple(norm.weight.shape) != (C, T):
        raise RuntimeError(f"{who}: input has {C} channels x {T} time samples but the LayerNorm was built for "
                           f"normalized_shape={list(norm.weight.shape)} (time_dimension mismatch)")


SITE_CONV, SITE_ATTN, SITE_PROJ, SITE_FFN_HID, SITE_FFN_OUT = 0, 1, 2, 3, 4   # csrc/common.cuh


class _AttentionFn(torch.autograd.Function):
    """softmax(QK^T / sqrt(64)) (+ Philox dropout on the probabilities) . V on (B,T,192) = [q|k|v]."""

    @staticmethod
    def forward(ctx, qkv, p, train, layer, seed):
        qkv = L.f32c(qkv)
        B, T, _ = qkv.shape
        out = torch.empty(B, T, 64, dtype=torch.float32, device=qkv.device)
        lse = torch.empty(B, 8, T, dtype=torch.float32, device=qkv.device)
        L.call("eegclip_attention_forward", L.ptr(qkv), L.ptr(out), L.ptr(lse), B, T, float(p), int(train), int(layer), int(seed),
               L.default_math(), L.stream())
        ctx.saved, ctx.meta = (qkv, out, lse), (B, T, float(p), int(train), int(layer), int(seed), L.default_math())
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, out, lse = ctx.saved
        B, T, p, train, layer, seed, math = ctx.meta
        dqkv = torch.empty_like(qkv)
        L.call("eegclip_attention_backward", L.ptr(qkv), L.ptr(out), L.ptr(L.f32c(dout)), L.ptr(lse), L.ptr(dqkv), B, T, p, train,
               layer, seed, math, L.stream())
        return dqkv, None, None, None, None


class _LayerNorm64Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, g, b):
        x, g, b = L.f32c(x), L.f32c(g), L.f32c(b)
        out = torch.empty_like(x)
        L.call("eegclip_layernorm_forward", L.ptr(x), L.ptr(g), L.ptr(b), L.ptr(out), x.numel() // 64, 64, L.stream())
        ctx.saved = (x, g)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, g = ctx.saved
        dx, dg, db = torch.empty_like(x), torch.empty_like(g), torch.empty_like(g)
        L.call("eegclip_layernorm_backward", L.ptr(L.f32c(dout)), L.ptr(x), L.ptr(g), L.ptr(dx), L.ptr(dg), L.ptr(db), x.numel() // 64, 64,
               L.stream())
        return dx, dg, db


class _DropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p, layer, site, seed):
        x = L.f32c(x)
        if x.numel() % 4:
            raise L.EegclipError("dropout kernel: element count must be a multiple of 4")
        out = torch.empty_like(x)
        L.call("eegclip_dropout", L.ptr(x), L.ptr(out), x.numel(), float(p), 1, int(layer), int(site), int(seed), L.stream())
        ctx.meta = (float(p), int(layer), int(site), int(seed))
        return out

    @staticmethod
    def backward(ctx, dout):
        p, layer, site, seed = ctx.meta
        dout = L.f32c(dout)
        dx = torch.empty_like(dout)
        L.call("eegclip_dropout", L.ptr(dout), L.ptr(dx), dout.numel(), p, 1, layer, site, seed, L.stream())
        return dx, None, None, None, None


class _GeluDropFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pre, p, train, layer, site, seed):
        pre = L.f32c(pre)
        out = torch.empty_like(pre)
        L.call("eegclip_gelu_dropout_forward", L.ptr(pre), L.ptr(out), pre.numel(), float(p), int(train), int(layer), int(site),
               int(seed), L.stream())
        ctx.saved, ctx.meta = pre, (float(p), int(train), int(layer), int(site), int(seed))
        return out

    @staticmethod
    def backward(ctx, dout):
        p, train, layer, site, seed = ctx.meta
        dpre = torch.empty_like(ctx.saved)
        L.call("eegclip_gelu_dropout_backward", L.ptr(ctx.saved), L.ptr(L.f32c(dout)), L.ptr(dpre), dpre.numel(), p, train, layer, site,
               seed, L.stream())
        return dpre, None, None, None, None, None


class LayerNorm(nn.LayerNorm):
    """nn.LayerNorm(64) of the transformer block (clip_model.py:84,89) on the eegclip per-token kernel when called directly."""

    def forward(self, x):
        _require_cuda(x, "LayerNorm")
        if tuple(self.normalized_shape) != (64,) or x.shape[-1] != 64:
            raise L.EegclipError("the eegclip LayerNorm kernel covers 64 features (the reference's only use)")
        return _LayerNorm64Fn.apply(x, self.weight, self.bias)


class Dropout(nn.Dropout):
    """nn.Dropout with the counter-based Philox masks of the fused kernels (stream = (layer, site)) when called directly."""

    def __init__(self, p=0.5, site=SITE_PROJ):
        super().__init__(p)
        self.site, self.layer = site, 0

    def forward(self, x):
        if not self.training or self.p == 0.0:
            return x
        _require_cuda(x, "Dropout")
        return _DropoutFn.apply(x, self.p, self.layer, self.site, L.new_seed())


# ---------------------------------------------------------------------------------------------------
# Transformer blocks (parameter containers keep the reference key names).  TransformerEncoderBlock.forward is ONE fused
# call; the sub-modules stay callable on their own (the reference's surface) through the same kernels.
# ---------------------------------------------------------------------------------------------------
class MultiHeadAttention(nn.Module):
    """clip_model.py:19-45.  ``mask`` is dead in the reference (``energy.mask_fill`` does not exist, :37)."""

    def __init__(self, emb_size, num_heads, dropout):
        super().__init__()
        self.emb_size, self.num_heads = emb_size, num_heads
        self.keys = nn.Linear(emb_size, emb_size)
        self.queries = nn.Linear(emb_size, emb_size)
        self.values = nn.Linear(emb_size, emb_size)
        self.att_drop = nn.Dropout(dropout)
        self.projection = nn.Linear(emb_size, emb_size)

    def forward(self, x, mask=None):
        if mask is not None:  # same failure as the reference: Tensor has no attribute mask_fill
            raise AttributeError("'Tensor' object has no attribute 'mask_fill'")
        _require_cuda(x, "MultiHeadAttention")
        if self.emb_size != 64 or self.num_heads != 8:
            raise L.EegclipError("the attention kernels are specialised for emb_size=64, 8 heads (the reference's only use)")
        q = _LinearFn.apply(x, self.queries.weight, self.queries.bias)
        k = _LinearFn.apply(x, self.keys.weight, self.keys.bias)
        v = _LinearFn.apply(x, self.values.weight, self.values.bias)
        train = self.training and self.att_drop.p > 0
        o = _AttentionFn.apply(torch.cat([q, k, v], dim=-1), self.att_drop.p, train, getattr(self, "layer", 0),
                               L.new_seed() if train else 0)
        return _LinearFn.apply(o, self.projection.weight, self.projection.bias)


class ResidualAdd(nn.Module):
    """clip_model.py:48-57: x + fn(x) (inside TransformerEncoderBlock the add is fused into the GEMM epilogues)."""

    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, x, **kwargs):
        _require_cuda(x, "ResidualAdd")
        return self.fn(x, **kwargs) + x


class FeedForwardBlock(nn.Sequential):
    """clip_model.py:60-67: Linear -> GELU -> Dropout -> Linear."""

    def __init__(self, emb_size, expansion, drop_p):
        super().__init__(nn.Linear(emb_size, expansion * emb_size), nn.GELU(), nn.Dropout(drop_p),
                         nn.Linear(expansion * emb_size, emb_size))

    def forward(self, x):
        _require_cuda(x, "FeedForwardBlock")
        pre = _LinearFn.apply(x, self[0].weight, self[0].bias)
        train = self.training and self[2].p > 0
        f = _GeluDropFn.apply(pre, self[2].p, train, getattr(self, "layer", 0), SITE_FFN_HID, L.new_seed() if train else 0)
        return _LinearFn.apply(f, self[3].weight, self[3].bias)


_XF_ORDER = ("0.fn.0.weight", "0.fn.0.bias", "0.fn.1.queries.weight", "0.fn.1.queries.bias", "0.fn.1.keys.weight",
             "0.fn.1.keys.bias", "0.fn.1.values.weight", "0.fn.1.values.bias", "0.fn.1.projection.weight",
             "0.fn.1.projection.bias", "1.fn.0.weight", "1.fn.0.bias", "1.fn.1.0.weight", "1.fn.1.0.bias",
             "1.fn.1.3.weight", "1.fn.1.3.bias")


class TransformerEncoderBlock(nn.Sequential):
    """clip_model.py:75-94: x + Drop(MHA(LN(x))); x + Drop(FFN(LN(x))) as one fused call."""

    def __init__(self, emb_size, num_heads=8, drop_p=0.5, forward_expansion=4, forward_drop_p=0.5):
        super().__init__(
            ResidualAdd(nn.Sequential(LayerNorm(emb_size), MultiHeadAttention(emb_size, num_heads, drop_p), Dropout(drop_p, SITE_PROJ))),
            ResidualAdd(nn.Sequential(LayerNorm(emb_size),
                                      FeedForwardBlock(emb_size, expansion=forward_expansion, drop_p=forward_drop_p),
                                      Dropout(drop_p, SITE_FFN_OUT))))
        if emb_size != 64 or num_heads != 8 or forward_expansion != 4:
            raise L.EegclipError("the B200 kernels are specialised for emb_size=64, 8 heads, expansion 4 (the reference's only use)")
        self.drop_p, self.forward_drop_p = drop_p, forward_drop_p

    def abi_params(self):
        named = dict(self.named_parameters())
        return [named[k] for k in _XF_ORDER]

    def forward(self, x, layer=0):
        _require_cuda(x, "TransformerEncoderBlock")
        B, T, _ = x.shape
        d = L.XfBlockDesc(B=B, T=T, layer=layer, train=int(self.training), math=L.default_math(), p_attn=self.drop_p,
                          p_proj=self.drop_p, p_ffn_hid=self.forward_drop_p, p_ffn_out=self.drop_p,
                          seed=L.new_seed() if self.training else 0)
        return _XfBlockFn.apply(x, d, *self.abi_params())


class TransformerEncoder(nn.Sequential):
    """clip_model.py:97-99."""

    def __init__(self, depth, emb_size):
        super().__init__(*[TransformerEncoderBlock(emb_size) for _ in range(depth)])

    def forward(self, x):
        for i, blk in enumerate(self):
            x = blk(x, layer=i)
        return x


# ---------------------------------------------------------------------------------------------------
# Conv front block
# ---------------------------------------------------------------------------------------------------
class BasicBlock(nn.Module):
    """clip_model.py:234-249: Conv1d('same') -> Dropout -> LayerNorm([C,T]) -> GELU on (B,C,T) tensors."""

    def __init__(self, in_channels, out_channels, kernel_size=64, time_dimension=320, dropout_rate=0.2, stride=1,
                 padding='same', dilation=1, activation=nn.LeakyReLU()):
        super().__init__()
        if stride != 1 or dilation != 1 or padding != 'same':
            raise L.EegclipError("BasicBlock kernels cover stride=1, dilation=1, padding='same' (the reference's only use)")
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size, stride, padding, dilation=dilation)
        self.dropout = nn.Dropout(dropout_rate)
        self.normalization = nn.LayerNorm([out_channels, time_dimension])
        self.activation = nn.GELU()  # the ``activation`` argument is ignored by the reference too (:240-241)
        self._act = 0

    def forward_time_major(self, x, skip=None, layer=0):
        """x, skip: (B,T,Cin) -> (B,T,Cout)."""
        _require_cuda(x, "BasicBlock")
        B, T, cin = x.shape
        w = self.conv.weight
        _check_ln_shape(self.normalization, w.shape[0], T, "BasicBlock")
        d = L.ConvBlockDesc(B=B, T=T, Cin=cin, Cout=w.shape[0], taps=w.shape[2], act=self._act, train=int(self.training),
                            math=L.default_math(), p_drop=self.dropout.p, layer=layer,
                            seed=L.new_seed() if self.training else 0)
        return _ConvBlockFn.apply(x, skip, d, w, self.conv.bias, self.normalization.weight, self.normalization.bias)

    def forward(self, x):
        # reference layout is channel-major (B,C,T); the kernels are time-major
        return self.forward_time_major(x.transpose(1, 2)).transpose(1, 2)


# ---------------------------------------------------------------------------------------------------
# EEG towers
# ---------------------------------------------------------------------------------------------------
class _TowerBase(nn.Module):
    _kind = L.TOWER_INTERLEAVED

    def _conv_blocks(self):
        return [getattr(self, f"conv_{i}") for i in range(self._n_conv)]

    def _xf_blocks(self):
        raise NotImplementedError

    def abi_params(self):
        ps = [self.eeg_spatial_mapping.weight, self.eeg_spatial_mapping.bias]
        for c in self._conv_blocks():
            ps += [c.conv.weight, c.conv.bias, c.normalization.weight, c.normalization.bias]
        for blk in self._xf_blocks():
            ps += blk.abi_params()
        return ps + [self.final_layer.weight, self.final_layer.bias]

    def get_output_dim(self, input_window_size):
        return input_window_size * self.output_dim

    def forward(self, x):
        _require_cuda(x, type(self).__name__)
        B, T, cin = x.shape
        convs, xfs = self._conv_blocks(), self._xf_blocks()
        if cin != 64:
            raise L.EegclipError("EEG towers take 64-channel windows (B,T,64)")
        taps = convs[0].conv.weight.shape[2] if convs else 1
        for c in convs:
            _check_ln_shape(c.normalization, 64, T, type(self).__name__)
        blk0 = xfs[0] if xfs else None
        d = L.TowerDesc(kind=self._kind, B=B, T=T, n_conv=len(convs), depth=len(xfs), taps=taps, latent=self.output_dim,
                        train=int(self.training), math=L.default_math(),
                        p_conv=convs[0].dropout.p if convs else 0.0,
                        p_attn=blk0.drop_p if blk0 else 0.0, p_proj=blk0.drop_p if blk0 else 0.0,
                        p_ffn_hid=blk0.forward_drop_p if blk0 else 0.0, p_ffn_out=blk0.drop_p if blk0 else 0.0,
                        seed=L.new_seed() if self.training else 0)
        return _TowerFn.apply(x, d, *self.abi_params())


class EEGConformer(_TowerBase):
    """clip_model.py:327-398: conv stack (input skip except last block) then TransformerEncoder(depth)."""
    _kind = L.TOWER_SEQUENTIAL

    def __init__(self, output_dim=8, conformer_input_dim=64, dropout_rate=0.2, eeg_dim=64, filters=(64,) * 2,
                 kernels=(64,) * 2, dilation_rate=1, input_channels=64, time_dimension=64 * 5, depth=2,
                 normalization_fn='layer_norm', activation_fn='leaky_relu'):
        super().__init__()
        self.spatial_filters, self.output_dim = input_channels, output_dim
        self.eeg_spatial_mapping = nn.Conv1d(eeg_dim, filters[0], kernel_size=1)
        self.n_blocks = self._n_conv = len(filters)
        for i, (f, k) in enumerate(zip(filters, kernels)):
            setattr(self, f"conv_{i}", BasicBlock(f, f, kernel_size=k, dilation=dilation_rate, time_dimension=time_dimension,
                                                  dropout_rate=dropout_rate))
        self.transformerEncoder = TransformerEncoder(depth, conformer_input_dim)
        self.final_layer = nn.Linear(conformer_input_dim, output_dim)

    def _xf_blocks(self):
        return list(self.transformerEncoder)


class EEGConformerInterleaved(_TowerBase):
    """clip_model.py:400-474: depth x [BasicBlock(x + eeg_x) -> 1-layer TransformerEncoder(+ eeg_x unless last)]."""
    _kind = L.TOWER_INTERLEAVED

    def __init__(self, output_dim=8, conformer_input_dim=64, dropout_rate=0.2, eeg_dim=64, filters=(64,) * 1,
                 kernels=(64,) * 1, dilation_rate=1, input_channels=64, time_dimension=64 * 5, depth=4,
                 normalization_fn='layer_norm', activation_fn='leaky_relu'):
        super().__init__()
        self.spatial_filters, self.output_dim = input_channels, output_dim
        self.eeg_spatial_mapping = nn.Conv1d(eeg_dim, filters[0], kernel_size=1)
        self.n_blocks = self._n_conv = depth
        for i in range(depth):
            setattr(self, f"conv_{i}", BasicBlock(filters[0], filters[0], kernel_size=kernels[0], dilation=dilation_rate,
                                                  time_dimension=time_dimension, dropout_rate=dropout_rate))
            setattr(self, f"conformer_{i}", TransformerEncoder(1, conformer_input_dim))
        self.final_layer = nn.Linear(conformer_input_dim, output_dim)

    def _xf_blocks(self):
        return [getattr(self, f"conformer_{i}")[0] for i in range(self._n_conv)]


# ---------------------------------------------------------------------------------------------------
# Speech towers (boundary: SURVEY §8(a14)).  The conv/LN blocks run on the kernels above; the two
# bi-LSTMs of the default tower run on the recurrence kernels of csrc/lstm.cuh (other LSTM shapes raise: no library fallback).
# ---------------------------------------------------------------------------------------------------
class SpeechSmallConv(nn.Module):
    """clip_model.py:204-232: Conv1d(speech_dim->out, k, 'same') -> Dropout -> LayerNorm([out,T]) -> LeakyReLU."""

    def __init__(self, output_dim=64, ks_temporal=20, dropout_rate=0.2, speech_dim=1024, time_dimension=64 * 5):
        super().__init__()
        self.speech_spatial_mapping = nn.Conv1d(speech_dim, output_dim, kernel_size=ks_temporal, padding='same')
        self.dropout = nn.Dropout(dropout_rate)
        self.layernorm = nn.LayerNorm([output_dim, time_dimension])
        self.activation = nn.LeakyReLU()
        self.output_dim = output_dim

    def get_output_dim(self, input_window_size):
        return int(input_window_size * self.output_dim)

    def forward(self, x):
        _require_cuda(x, "SpeechSmallConv")
        B, T, cin = x.shape
        w = self.speech_spatial_mapping.weight
        _check_ln_shape(self.layernorm, w.shape[0], T, "SpeechSmallConv")
        d = L.ConvBlockDesc(B=B, T=T, Cin=cin, Cout=w.shape[0], taps=w.shape[2], act=1, train=int(self.training),
                            math=L.default_math(), p_drop=self.dropout.p, layer=0, seed=L.new_seed() if self.training else 0)
        return _ConvBlockFn.apply(x, None, d, w, self.speech_spatial_mapping.bias, self.layernorm.weight, self.layernorm.bias)


class EEGConvLSTM(nn.Module):
    """clip_model.py:251-325: 1x1 conv -> BasicBlocks (input skip except last) -> two bi-LSTMs."""

    def __init__(self, units_lstm=128, output_dim=64, dropout_rate=0.2, eeg_dim=64, filters=(256, 256, 256, 128, 128),
                 kernels=(64,) * 5, dilation_rate=1, input_channels=64, time_dimension=64 * 5,
                 normalization_fn='layer_norm', activation_fn='leaky_relu'):
        super().__init__()
        self.speech_lstm1 = nn.LSTM(filters[-1], units_lstm, batch_first=True, bidirectional=True)
        self.speech_lstm2 = nn.LSTM(units_lstm * 2, int(output_dim / 2), batch_first=True, bidirectional=True)
        self.spatial_filters, self.output_dim = input_channels, output_dim
        self.eeg_spatial_mapping = nn.Conv1d(eeg_dim, filters[0], kernel_size=1)
        self.n_blocks = len(filters)
        for i, (f, k) in enumerate(zip(filters, kernels)):
            setattr(self, f"conv_{i}", BasicBlock(f, f, kernel_size=k, dilation=dilation_rate, time_dimension=time_dimension,
                                                  dropout_rate=dropout_rate))

    def get_output_dim(self, input_window_size):
        return input_window_size * self.output_dim

    def forward(self, x):
        _require_cuda(x, "EEGConvLSTM")
        x = _LinearFn.apply(x, self.eeg_spatial_mapping.weight, self.eeg_spatial_mapping.bias)  # (B,T,f0), time-major
        eeg = x
        for i in range(self.n_blocks):
            blk = getattr(self, f"conv_{i}")
            x = blk.forward_time_major(x, None if i == self.n_blocks - 1 else eeg, layer=i)
        x = _bilstm(self.speech_lstm1, x)
        x = _bilstm(self.speech_lstm2, x)
        return x


# ---------------------------------------------------------------------------------------------------
# Two-stream towers.  The two encoders of a CLIP wrapper are independent until the head, and the speech tower is a chain of
# latency-bound launches (two bi-LSTM recurrences on 128 SMs, skinny projections) while the EEG tower's 1.73-wave convolutions
# leave 40 SMs idle in their second round: the speech tower therefore runs on a side stream next to the EEG tower, in the
# forward AND -- autograd runs a node's backward on the stream of its forward -- in the backward.  Ordering: the side stream
# waits for the caller's stream before it starts (inputs, parameter updates of the previous step), the caller's stream waits
# for the side stream before the head, and once more when the backward pass has been enqueued (a final engine callback), so
# optimizer.step(), gradient all-reduces and reads of .grad on the caller's stream see every gradient.
# EEGCLIP_TWO_STREAMS=0 runs both towers on the caller's stream (identical results: the kernels and their order per tower are the same).
# ---------------------------------------------------------------------------------------------------
import os as _os

_SIDE_STREAMS = {}


def two_streams_enabled():
    return _os.environ.get("EEGCLIP_TWO_STREAMS", "1") != "0"


def _side_stream(device):
    s = _SIDE_STREAMS.get(device.index)
    if s is None:
        s = _SIDE_STREAMS[device.index] = torch.cuda.Stream(device=device)
    return s


class _RejoinFn(torch.autograd.Function):
    """Identity on the side-stream tower's output, applied on the caller's stream after it has waited for the side stream; its
    backward queues the end-of-backward wait of the caller's stream on the side stream."""

    @staticmethod
    def forward(ctx, x, main, side):
        ctx.main, ctx.side = main, side
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        main, side = ctx.main, ctx.side
        torch.autograd.Variable._execution_engine.queue_callback(lambda: main.wait_stream(side))
        return g, None, None


def run_towers(eegModel, speechModel, eeg, speech, speech_first=True):
    """(eegModel(eeg), speechModel(speech)) with the speech tower on a side stream (see above).  ``speech_first`` is the order in
    which the two towers are called on the host -- the order of their dropout-seed draws and, reversed, of their backward nodes: the
    CLI default wrapper calls the speech tower first (data-parallel bucket overlap), the other wrappers the EEG tower, as the
    reference does."""
    if not (two_streams_enabled() and eeg.is_cuda and speech.is_cuda and eeg.device == speech.device):
        if speech_first:
            sf = speechModel(speech)
            return eegModel(eeg), sf
        ef = eegModel(eeg)
        return ef, speechModel(speech)
    main, side = torch.cuda.current_stream(eeg.device), _side_stream(eeg.device)
    side.wait_stream(main)                      # (before the EEG tower is enqueued: the side stream must not wait for it)
    if not speech_first:
        ef = eegModel(eeg)
    with torch.cuda.stream(side):
        sf = speechModel(speech)
    speech.record_stream(side)
    if speech_first:
        ef = eegModel(eeg)
    main.wait_stream(side)
    sf.record_stream(main)
    if sf.requires_grad:
        sf = _RejoinFn.apply(sf, main, side)
    return ef, sf


# ---------------------------------------------------------------------------------------------------
# Loss wrappers
# ---------------------------------------------------------------------------------------------------
class CLIP(nn.Module):
    """clip_model.py:657-693: symmetric InfoNCE with a learnable log-scale."""

    def __init__(self, eegModel, speechModel, temperature=1.):
        super().__init__()
        self.eegModel, self.speechModel = eegModel, speechModel
        self.temperature = nn.Parameter(torch.tensor(temperature))
        self.shard_group = None  # set to a torch.distributed group for sharded InfoNCE

    def forward(self, eeg, speech):
        ef, sf = run_towers(self.eegModel, self.speechModel, eeg, speech)
        E, S = torch.flatten(ef, start_dim=1), torch.flatten(sf, start_dim=1)
        return infonce_loss(E, S, self.temperature, group=self.shard_group)


class memoryBank(nn.Module):
    """clip_model.py:697-745: EMA bank indexed by segment id; updated in train *and* eval mode."""

    def __init__(self, bank_size, device, dim, momentum=0.90):
        super().__init__()
        self.bank_size, self.dim, self.momentum = bank_size, dim, momentum
        self.register_buffer("memory", torch.rand(bank_size + 1, dim).to(device))

    def forward(self, idx, data):
        _require_cuda(data, "memoryBank")
        d = L.f32c(data)
        idx = idx.view(-1).to(torch.int64).contiguous()
        if idx.device.type == "cpu":
            if idx.numel() and (int(idx.min()) < 0 or int(idx.max()) >= self.memory.shape[0]):
                raise IndexError(f"memoryBank: index out of range for a bank of {self.memory.shape[0]} rows")   # as index_select (:738)
            idx = idx.to(d.device)
        old = torch.empty_like(d)
        # ids already on the device are range-checked by the kernel (no host sync): an out-of-range row writes nothing and reads NaN
        L.call("eegclip_membank_update", L.ptr(self.memory), self.memory.shape[0], L.ptr(idx), L.ptr(d), L.ptr(old), d.shape[0],
               d.shape[1], float(self.momentum), float(1 - self.momentum), L.stream())
        return old


class _ZeroGradLink(torch.autograd.Function):
    """value passes through; ``anchors`` receive exact-zero gradients (the reference's lambda_average == 0 case:
    temperature_eeg.grad is tensor(0.), not None, so AdamW still decays it)."""

    @staticmethod
    def forward(ctx, value, *anchors):
        ctx.shapes = [a.shape for a in anchors]
        ctx.dev = value.device
        return value.clone()

    @staticmethod
    def backward(ctx, g):
        return (None, *[torch.zeros(s, device=ctx.dev) for s in ctx.shapes])


class CLIPSimNoLatentProj(nn.Module):
    """clip_model.py:868-944 (CLI default): CLIP loss + memory-bank cross-entropy, returns three 0-dim losses."""

    def __init__(self, eegModel, speechModel, eegMemoryBank, temperature=1., window_length=192, lambda_clip=1,
                 lambda_average=1):
        super().__init__()
        self.eegModel, self.speechModel, self.eegMemoryBank = eegModel, speechModel, eegMemoryBank
        self.window_length, self.lambda_clip, self.lambda_average = window_length, lambda_clip, lambda_average
        self.temperature = nn.Parameter(torch.tensor(temperature))
        self.temperature_eeg = nn.Parameter(torch.tensor(temperature))
        self.shard_group = None

    def forward(self, eeg, speech, ids):
        # the speech tower runs first so that autograd runs the (three times longer) EEG tower backward first: under data
        # parallelism its 14.5 MB of gradients are all-reduced while the speech tower's backward is still running
        ef, sf = run_towers(self.eegModel, self.speechModel, eeg, speech)
        if sf.shape[1] > sf.shape[2]:
            sf = sf.transpose(1, 2)
        if ef.shape[1] > ef.shape[2]:
            ef = ef.transpose(1, 2)
        E_raw, S_raw = torch.flatten(ef, start_dim=1), torch.flatten(sf, start_dim=1)
        loss_ce, En, E_all = infonce_loss(E_raw, S_raw, self.temperature, group=self.shard_group, return_normalized="all")
        world, rank = _world(self.shard_group)
        if world > 1:
            # data-parallel: every rank applies the SAME bank update (ids of all ranks, gathered embeddings), so the banks stay
            # identical across ranks and equal to a single process at the global batch; this rank keeps its rows of the old values
            if self.lambda_average != 0:
                raise L.EegclipError("CLIPSimNoLatentProj: the memory-bank loss term (lambda_average != 0) is not sharded; train it "
                                     "on one rank or with lambda_sim_loss = 0 (the reference default)")
            ids_all = _all_gather_rows(ids.view(-1).to(device=En.device, dtype=torch.int64).contiguous(), self.shard_group, world)
            b = En.shape[0]
            avg = self.eegMemoryBank(ids_all, E_all.detach())[rank * b:(rank + 1) * b]
        else:
            avg = self.eegMemoryBank(ids, En.detach())
        if self.lambda_average == 0:
            with torch.no_grad():
                avg_val = ce_rows_loss(avg, En.detach(), self.temperature_eeg.detach())
            avg_loss = _ZeroGradLink.apply(avg_val, self.temperature_eeg)
        else:
            # gradient reaches the EEG tower through the normalised embeddings here (clip_model.py:934)
            avg_loss = ce_rows_loss(avg, l2_normalize(E_raw), self.temperature_eeg)
        loss_total = self.lambda_clip * loss_ce + self.lambda_average * avg_loss
        return loss_ce.mean(), avg_loss.mean(), loss_total.mean()


# ---------------------------------------------------------------------------------------------------
# Non-default loss wrappers (train_clip_final.py:379-396: --model_arch clip_sim / clip_mp / clip_kld / no_contrastive_learning).
# All of them sit on the same kernels: towers, _LinearFn (latent projections and the S.E^T similarity GEMM with its two
# gradient GEMMs), l2-normalisation, and -- where the logits are square -- the fused symmetric-InfoNCE head.  What is left in
# torch is O(B x n.B) elementwise / log-sum-exp arithmetic on the similarity matrix and O(B x latent) regularisers.
# ---------------------------------------------------------------------------------------------------
def _flat(x):
    return torch.flatten(x, start_dim=1)


def _single_rank_only(module):
    """The non-default wrappers carry batch-mean regularisers and per-rank state that the SUM all-reduce of the sharded scheme would
    mis-scale: data parallelism is implemented for CLIP and CLIPSimNoLatentProj (the CLI default) only."""
    if _world(getattr(module, "shard_group", None))[0] > 1:
        raise L.EegclipError(f"{type(module).__name__}: sharded (data-parallel) training is implemented for CLIP / CLIPSimNoLatentProj only")


def _proj(lin, x):
    return _LinearFn.apply(x, lin.weight, lin.bias)


def _similarity(Sn, En, log_scale):
    """(Sn . En^T) * exp(log_scale) with autograd to both operands (tcgen05 token-GEMM kernels)."""
    return _LinearFn.apply(Sn, En, None) * torch.exp(log_scale)


def _ce_columns(logits):
    """F.cross_entropy(logits.T, targets) with targets = arange(rows) tiled over the columns (clip_model.py:1044-1049)."""
    rows = logits.shape[0]
    col = torch.arange(logits.shape[1], device=logits.device)
    return (torch.logsumexp(logits, dim=0) - logits[col % rows, col]).mean()


class CLIPSim(nn.Module):
    """clip_model.py:747-810: latent projections (no bias) -> symmetric InfoNCE + MSE(normalize(bank average), EEG embedding)."""

    def __init__(self, eegModel, speechModel, eegMemoryBank, temperature=1., latent_dim=16, window_length=192, lambda_clip=1,
                 lambda_average=1):
        super().__init__()
        self.eegModel, self.speechModel, self.eegMemoryBank = eegModel, speechModel, eegMemoryBank
        self.latent_dim, self.window_length = latent_dim, window_length
        self.lambda_clip, self.lambda_average = lambda_clip, lambda_average
        d_in = self.eegModel.get_output_dim(input_window_size=self.window_length)
        self.latent_projection_eeg = nn.Linear(d_in, latent_dim, bias=False)
        self.latent_projection_speech = nn.Linear(d_in, latent_dim, bias=False)
        self.temperature = nn.Parameter(torch.tensor(temperature))
        self.shard_group = None

    def forward(self, eeg, speech, ids):
        _single_rank_only(self)
        ef, sf = run_towers(self.eegModel, self.speechModel, eeg, speech, speech_first=False)
        E = _proj(self.latent_projection_eeg, _flat(ef))
        S = _proj(self.latent_projection_speech, _flat(sf))
        loss_ce, En = infonce_loss(E, S, self.temperature, group=self.shard_group, return_normalized=True)
        avg = l2_normalize(self.eegMemoryBank(ids, En))
        avg_loss = ((avg - l2_normalize(E)) ** 2).mean()          # the gradient reaches the EEG tower through the normalisation
        loss_total = self.lambda_clip * loss_ce + self.lambda_average * avg_loss
        return loss_ce.mean(), avg_loss.mean(), loss_total.mean()


class CLIPNoContrastiveLearning(nn.Module):
    """clip_model.py:948-995: binary cross-entropy on the diagonal (match) vs the first super-diagonal (mismatch)."""

    def __init__(self, eegModel, speechModel, window_length=192):
        super().__init__()
        self.eegModel, self.speechModel, self.window_length = eegModel, speechModel, window_length

    def forward(self, eeg, speech, ids):
        ef, sf = run_towers(self.eegModel, self.speechModel, eeg, speech, speech_first=False)
        if sf.shape[1] > sf.shape[2]:
            sf = sf.transpose(1, 2)
        if ef.shape[1] > ef.shape[2]:
            ef = ef.transpose(1, 2)
        En, Sn = l2_normalize(_flat(ef)), l2_normalize(_flat(sf))
        logits = _LinearFn.apply(Sn, En, None)
        match, mismatch = torch.diagonal(logits)[:-1], torch.diagonal(logits, offset=1)
        loss = (F.softplus(-match).mean() + F.softplus(mismatch).mean()) / 2     # BCE-with-logits, targets 1 / 0
        return loss.mean(), loss.mean(), loss.mean()


def log_softmax_mp(x):
    """clip_model.py:1484-1487 on x (B, n, B): log sum_n exp x[i,n,j] - log sum_{n,j} exp x[i,n,j]."""
    return torch.logsumexp(x, dim=-2) - torch.logsumexp(x.flatten(-2), dim=-1, keepdim=True)


def multiple_postives_loss(preds, targets, reduction='mean'):
    """clip_model.py:1490-1494 (name as in the reference)."""
    return F.nll_loss(log_softmax_mp(preds), targets, reduction=reduction)


def simloss(x, target):
    """clip_model.py:1478-1480."""
    return F.nll_loss(x.sum(-2), target)


class CLIPSimMultiplePositives(nn.Module):
    """clip_model.py:1000-1078: the EEG batch holds n windows per speech segment; logits (B, n*B)."""

    def __init__(self, eegModel, speechModel, temperature=1., window_length=192, lambda_clip=1, lambda_average=1):
        super().__init__()
        self.eegModel, self.speechModel, self.window_length = eegModel, speechModel, window_length
        self.lambda_clip, self.lambda_average = lambda_clip, lambda_average
        self.temperature = nn.Parameter(torch.tensor(temperature))
        self.temperature_eeg = nn.Parameter(torch.tensor(temperature))

    def _logits(self, eeg, speech):
        ef, sf = run_towers(self.eegModel, self.speechModel, eeg, speech, speech_first=False)
        En, Sn = l2_normalize(_flat(ef)), l2_normalize(_flat(sf))
        return _similarity(Sn, En, self.temperature)

    def forward(self, eeg, speech, ids):
        logits = self._logits(eeg, speech)
        B = logits.shape[0]
        eeg_loss = _ce_columns(logits)
        grouped = logits.reshape(B, -1, B)
        targets = torch.arange(B, device=logits.device)
        speech_loss = multiple_postives_loss(grouped, targets)
        sim_loss = simloss(grouped, targets)
        loss_ce = (speech_loss + eeg_loss) / 2.0
        loss_total = self.lambda_clip * loss_ce + self.lambda_average * sim_loss
        return loss_ce.mean(), sim_loss.mean(), loss_total.mean()


class CLIPSimMultiplePositivesAdapted(CLIPSimMultiplePositives):
    """clip_model.py:1081-1168: the n logits of a speech/EEG-group pair are summed before the row cross-entropy."""

    def forward(self, eeg, speech, ids):
        logits = self._logits(eeg, speech)
        B = logits.shape[0]
        eeg_loss = _ce_columns(logits)
        summed = logits.reshape(B, -1, B).sum(dim=1)
        speech_loss = (torch.logsumexp(summed, dim=1) - torch.diagonal(summed)).mean()
        loss_ce = (speech_loss + eeg_loss) / 2.0
        loss_total = self.lambda_clip * loss_ce
        return loss_ce.mean(), loss_ce.mean(), loss_total.mean()


def log_gauss(x, mu, logvar):
    """clip_model.py:1499-1501: log N(x; mu, exp(logvar))."""
    return -0.5 * (math.log(2 * math.pi) + logvar + (x - mu) ** 2 / torch.exp(logvar))


def kld(p_mu, p_logvar, q_mu, q_logvar):
    """clip_model.py:1503-1504: KL(N(p_mu, e^p_logvar) || N(q_mu, e^q_logvar)) per element."""
    return -0.5 * (1 + p_logvar - q_logvar - ((p_mu - q_mu) ** 2 + torch.exp(p_logvar)) / torch.exp(q_logvar))


def _kld_terms(mu2, z_mu, z_logvar):
    """Variational lower bound shared by both KLD wrappers (clip_model.py:1231-1239, 1407-1415)."""
    dev = z_mu.device
    q_logvar = torch.tensor([math.log(0.5 ** 2)], dtype=torch.float32, device=dev)
    zero = torch.zeros(1, dtype=torch.float32, device=dev)
    log_pmu2 = log_gauss(mu2, zero, zero).mean(dim=1)
    kld_z2 = kld(z_mu, z_logvar, mu2, q_logvar).mean(dim=1)
    return log_pmu2, kld_z2, (-log_pmu2 + kld_z2).mean(dim=0)


class _KLDBase(nn.Module):
    def reparameterize(self, mu, logvar):
        if self.training:
            return torch.randn_like(logvar).mul(torch.exp(0.5 * logvar)).add_(mu)
        return mu


class CLIPKLDNoLatentProj(_KLDBase):
    """clip_model.py:1174-1279: InfoNCE on the flattened tower outputs + KL of a per-segment latent (embedding table prior)."""

    def __init__(self, eegModel, speechModel, latent_dimension, number_of_classes, latent_dimension2=64, temperature=1.,
                 window_length=192, lambda_clip=1, lambda_lower_bound=1, lambda_discriminative=1):
        super().__init__()
        self.eegModel, self.speechModel, self.window_length = eegModel, speechModel, window_length
        self.lambda_clip, self.lambda_lower_bound, self.lambda_discriminative = lambda_clip, lambda_lower_bound, lambda_discriminative
        self.number_of_classes, self.latent_dimension, self.latent_dimension2 = number_of_classes, latent_dimension, latent_dimension2
        self.temperature = nn.Parameter(torch.tensor(temperature))
        self.temperature_eeg = nn.Parameter(torch.tensor(temperature))
        self.mu_eeg_lookup = nn.Embedding(number_of_classes + 1, latent_dimension2)
        self.eeg_mu_linear = nn.Linear(latent_dimension, latent_dimension2)
        self.eeg_logvar_linear = nn.Linear(latent_dimension, latent_dimension2)
        self.shard_group = None

    def encode(self, eeg, speech, ids):
        ef, sf = run_towers(self.eegModel, self.speechModel, eeg, speech, speech_first=False)
        E, S = _flat(ef), _flat(sf)
        mu2 = self.mu_eeg_lookup(ids)
        z_mu, z_logvar = _proj(self.eeg_mu_linear, E), _proj(self.eeg_logvar_linear, E)
        return mu2, z_mu, z_logvar, self.reparameterize(z_mu, z_logvar), S, E

    def forward(self, eeg, speech, ids):
        _single_rank_only(self)
        mu2, z_mu, z_logvar, _, S, E = self.encode(eeg, speech, ids)
        log_pmu2, kld_z2, lower_bound = _kld_terms(mu2, z_mu, z_logvar)
        loss_ce = infonce_loss(E, S, self.temperature, group=self.shard_group)
        loss_total = self.lambda_clip * loss_ce + self.lambda_lower_bound * lower_bound
        return loss_total.mean(), loss_ce.mean(), log_pmu2.mean(), kld_z2.mean()


class ProjectionHeadLinear(nn.Module):
    """clip_model.py:1303-1320: Linear -> LeakyReLU -> Linear."""

    def __init__(self, embedding_dim, projection_dim=512):
        super().__init__()
        self.projection = nn.Linear(embedding_dim, projection_dim * 2)
        self.relu = nn.LeakyReLU()
        self.last_linear = nn.Linear(projection_dim * 2, projection_dim)

    def forward(self, x):
        _require_cuda(x, "ProjectionHeadLinear")
        return _proj(self.last_linear, F.leaky_relu(_proj(self.projection, x), 0.01))


class CLIPKLDWithLatentProj(_KLDBase):
    """clip_model.py:1325-1450 with the default linear projection heads (the 'non-linear' head is outside the kernel set)."""

    def __init__(self, eegModel, speechModel, latent_dimension, number_of_classes, temperature=1., window_length=192,
                 lambda_clip=1, lambda_lower_bound=1, lambda_discriminative=1, projection_head='linear'):
        super().__init__()
        if projection_head != 'linear':
            raise L.EegclipError("CLIPKLDWithLatentProj: only projection_head='linear' (the reference default) is on the B200 path")
        self.eegModel, self.speechModel, self.window_length = eegModel, speechModel, window_length
        self.lambda_clip, self.lambda_lower_bound, self.lambda_discriminative = lambda_clip, lambda_lower_bound, lambda_discriminative
        self.number_of_classes, self.latent_dimension = number_of_classes, latent_dimension
        self.temperature = nn.Parameter(torch.tensor(temperature))
        self.temperature_eeg = nn.Parameter(torch.tensor(temperature))
        self.mu_eeg_lookup = nn.Embedding(number_of_classes + 1, latent_dimension)
        self.eeg_mu_linear = ProjectionHeadLinear(self.eegModel.get_output_dim(window_length), latent_dimension)
        self.eeg_logvar_linear = ProjectionHeadLinear(self.eegModel.get_output_dim(window_length), latent_dimension)
        self.speech_latent_projection = ProjectionHeadLinear(self.speechModel.get_output_dim(window_length), latent_dimension)
        self.shard_group = None

    def encode(self, eeg, speech, ids):
        ef, sf = run_towers(self.eegModel, self.speechModel, eeg, speech, speech_first=False)
        E, S = _flat(ef), _flat(sf)
        z_logvar, z_mu, Sp = self.eeg_logvar_linear(E), self.eeg_mu_linear(E), self.speech_latent_projection(S)
        return self.mu_eeg_lookup(ids), z_mu, z_logvar, self.reparameterize(z_mu, z_logvar), Sp, z_mu

    def forward(self, eeg, speech, ids):
        _single_rank_only(self)
        mu2, z_mu, z_logvar, _, Sp, Ep = self.encode(eeg, speech, ids)
        log_pmu2, kld_z2, lower_bound = _kld_terms(mu2, z_mu, z_logvar)
        loss_ce = infonce_loss(Ep, Sp, self.temperature, group=self.shard_group)   # normalises both sides, as :1383-1384
        loss_total = self.lambda_clip * loss_ce + self.lambda_lower_bound * lower_bound
        return loss_total.mean(), loss_ce.mean(), log_pmu2.mean(), kld_z2.mean()
