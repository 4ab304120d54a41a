"""TEST INFRASTRUCTURE ONLY -- deterministic synthetic weights / inputs.

Everything is drawn from numpy's frozen ``RandomState`` (MT19937) so the same
seed gives the same bits in the build container (where the real reference is
run to make ``tests/golden``) and on the GPU box (where the CUDA path and the
oracle are compared).  Shapes and key names are the reference's ``state_dict``
schema (observed by instantiating clip_model.py:400-474,
251-325, 204-232 and vlaai.py:74-134).
"""
import numpy as np
import torch


def _u(rs, shape, bound):
    return torch.from_numpy(rs.uniform(-bound, bound, size=shape).astype(np.float32))


def transformer_block_shapes(prefix, emb=64, expansion=4):
    """Keys of one TransformerEncoderBlock (clip_model.py:75-94) under ``prefix``."""
    s = {}
    s[prefix + "0.fn.0.weight"] = (emb,)
    s[prefix + "0.fn.0.bias"] = (emb,)
    for n in ("keys", "queries", "values", "projection"):
        s[prefix + f"0.fn.1.{n}.weight"] = (emb, emb)
        s[prefix + f"0.fn.1.{n}.bias"] = (emb,)
    s[prefix + "1.fn.0.weight"] = (emb,)
    s[prefix + "1.fn.0.bias"] = (emb,)
    s[prefix + "1.fn.1.0.weight"] = (expansion * emb, emb)
    s[prefix + "1.fn.1.0.bias"] = (expansion * emb,)
    s[prefix + "1.fn.1.3.weight"] = (emb, expansion * emb)
    s[prefix + "1.fn.1.3.bias"] = (emb,)
    return s


def basic_block_shapes(prefix, cin, cout, k, T):
    return {
        prefix + "conv.weight": (cout, cin, k),
        prefix + "conv.bias": (cout,),
        prefix + "normalization.weight": (cout, T),
        prefix + "normalization.bias": (cout, T),
    }


def interleaved_shapes(depth, T, latent=8, k=64, emb=64):
    """EEGConformerInterleaved state_dict schema (clip_model.py:400-439)."""
    s = {"eeg_spatial_mapping.weight": (emb, 64, 1), "eeg_spatial_mapping.bias": (emb,)}
    for i in range(depth):
        s.update(basic_block_shapes(f"conv_{i}.", emb, emb, k, T))
        s.update(transformer_block_shapes(f"conformer_{i}.0.", emb))
    s["final_layer.weight"] = (latent, emb)
    s["final_layer.bias"] = (latent,)
    return s


def conformer_shapes(n_conv, depth, T, latent=8, k=64, emb=64):
    """EEGConformer (sequential variant) schema (clip_model.py:327-370)."""
    s = {"eeg_spatial_mapping.weight": (emb, 64, 1), "eeg_spatial_mapping.bias": (emb,)}
    for i in range(n_conv):
        s.update(basic_block_shapes(f"conv_{i}.", emb, emb, k, T))
    for i in range(depth):
        s.update(transformer_block_shapes(f"transformerEncoder.{i}.", emb))
    s["final_layer.weight"] = (latent, emb)
    s["final_layer.bias"] = (latent,)
    return s


def small_conv_shapes(T, speech_dim=1024, out=8, k=16):
    """SpeechSmallConv schema (clip_model.py:204-217)."""
    return {
        "speech_spatial_mapping.weight": (out, speech_dim, k),
        "speech_spatial_mapping.bias": (out,),
        "layernorm.weight": (out, T),
        "layernorm.bias": (out, T),
    }


def conv_lstm_shapes(T, in_dim=1024, filters=(64,), kernels=(32,), units=128, out=8):
    """EEGConvLSTM schema (clip_model.py:251-283)."""
    s = {"eeg_spatial_mapping.weight": (filters[0], in_dim, 1), "eeg_spatial_mapping.bias": (filters[0],)}
    for i, (f, k) in enumerate(zip(filters, kernels)):
        s.update(basic_block_shapes(f"conv_{i}.", f, f, k, T))
    for name, inp, hid in (("speech_lstm1", filters[-1], units), ("speech_lstm2", 2 * units, out // 2)):
        for suf in ("", "_reverse"):
            s[f"{name}.weight_ih_l0{suf}"] = (4 * hid, inp)
            s[f"{name}.weight_hh_l0{suf}"] = (4 * hid, hid)
            s[f"{name}.bias_ih_l0{suf}"] = (4 * hid,)
            s[f"{name}.bias_hh_l0{suf}"] = (4 * hid,)
    return s


def vlaai_shapes(T=320):
    """VLAAI schema (vlaai.py:74-104)."""
    s = {}
    pre = "sequentialConvStack.0."
    s[pre + "eeg.weight"] = (64, 64, 1)
    s[pre + "eeg.bias"] = (64,)
    cin = 64
    for j, f in enumerate((256, 256, 256, 128, 128)):
        s[pre + f"conv_layers.{3 * j}.weight"] = (f, cin, 64)
        s[pre + f"conv_layers.{3 * j}.bias"] = (f,)
        s[pre + f"conv_layers.{3 * j + 1}.weight"] = (f, T)
        s[pre + f"conv_layers.{3 * j + 1}.bias"] = (f, T)
        cin = f
    s["sequentialConvStack.1.weight"] = (64, 128, 1)
    s["sequentialConvStack.1.bias"] = (64,)
    s["sequentialConvStack.2.conv1d.weight"] = (64, 64, 64)
    s["sequentialConvStack.2.conv1d.bias"] = (64,)
    s["sequentialConvStack.2.normalization_fn.weight"] = (64, T)
    s["sequentialConvStack.2.normalization_fn.bias"] = (64, T)
    s["eeg.weight"] = (64, 64, 1)
    s["eeg.bias"] = (64,)
    s["final_linear.weight"] = (64, 64, 1)
    s["final_linear.bias"] = (64,)
    return s


def make_state_dict(shapes, seed, prefix=""):
    """Fill a schema with PyTorch-default-like magnitudes (uniform +-1/sqrt(fan_in)).

    LayerNorm affines are drawn around (1, 0) instead of being exactly (1, 0) so
    that the (C,T)-shaped affine layout is actually exercised by parity tests.
    """
    rs = np.random.RandomState(seed)
    out = {}
    for key in sorted(shapes):
        shp = shapes[key]
        is_norm = ("normalization" in key or "layernorm" in key or ".fn.0." in key
                   or (key.startswith("sequentialConvStack.0.conv_layers.") and int(key.split(".")[3]) % 3 == 1))
        if is_norm:
            if key.endswith("weight"):
                t = 1.0 + _u(rs, shp, 0.25)
            else:
                t = _u(rs, shp, 0.25)
        elif len(shp) == 1:
            t = _u(rs, shp, 0.1)
        else:
            fan_in = int(np.prod(shp[1:]))
            t = _u(rs, shp, 1.0 / np.sqrt(fan_in))
        out[prefix + key] = t
    return out


def randn(seed, *shape):
    rs = np.random.RandomState(seed)
    return torch.from_numpy(rs.standard_normal(size=shape).astype(np.float32))


def grad_digest(t, n=16):
    """Small fingerprint of a tensor for golden files: norm, sum, abs-sum, n strided samples."""
    f = t.detach().double().reshape(-1)
    step = max(1, f.numel() // n)
    return {
        "shape": list(t.shape),
        "norm": float(f.norm()),
        "sum": float(f.sum()),
        "asum": float(f.abs().sum()),
        "samples": [float(x) for x in f[::step][:n]],
    }
