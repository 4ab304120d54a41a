"""ctypes binding of libeegclip_b200.so (C ABI declared in include/eegclip.h).

There is no CPU fallback and no alternative backend: if the shared library is missing or a tensor
is not a contiguous fp32 CUDA tensor the call raises.
"""
import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libeegclip_b200.so")

MATH_FP32, MATH_BF16X3, MATH_BF16 = 0, 1, 2
TOWER_INTERLEAVED, TOWER_SEQUENTIAL = 0, 1
_MATH_NAMES = {"fp32": MATH_FP32, "bf16x3": MATH_BF16X3, "bf16": MATH_BF16}

# arithmetic of the tensor-core kernels; "bf16x3" (split-bf16, fp32 accumulate) is the parity mode
_default_math = _MATH_NAMES[os.environ.get("EEGCLIP_MATH", "bf16x3").lower()]


def set_default_math(name):
    global _default_math
    _default_math = _MATH_NAMES[name.lower()] if isinstance(name, str) else int(name)


def default_math():
    return _default_math


class EegclipError(RuntimeError):
    pass


class TowerDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("B", C.c_int32), ("T", C.c_int32), ("n_conv", C.c_int32), ("depth", C.c_int32),
        ("taps", C.c_int32), ("latent", C.c_int32), ("train", C.c_int32), ("math", C.c_int32), ("reserved", C.c_int32),
        ("p_conv", C.c_float), ("p_attn", C.c_float), ("p_proj", C.c_float), ("p_ffn_hid", C.c_float),
        ("p_ffn_out", C.c_float), ("reserved_f", C.c_float), ("seed", C.c_uint64),
    ]


class ConvBlockDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("T", C.c_int32), ("Cin", C.c_int32), ("Cout", C.c_int32), ("taps", C.c_int32),
        ("act", C.c_int32), ("train", C.c_int32), ("math", C.c_int32), ("p_drop", C.c_float), ("layer", C.c_int32),
        ("seed", C.c_uint64),
    ]


class BiLstmDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("T", C.c_int32), ("In", C.c_int32), ("H", C.c_int32), ("math", C.c_int32), ("reserved", C.c_int32)]


class XfBlockDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("T", C.c_int32), ("layer", C.c_int32), ("train", C.c_int32), ("math", C.c_int32),
        ("reserved", C.c_int32), ("p_attn", C.c_float), ("p_proj", C.c_float), ("p_ffn_hid", C.c_float),
        ("p_ffn_out", C.c_float), ("seed", C.c_uint64),
    ]


_vp, _i32, _i64, _f32, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t
_psz = C.POINTER(C.c_size_t)

# symbol -> (restype, argtypes); must list every function declared in include/eegclip.h
SIGNATURES = {
    "eegclip_abi_version": (C.c_int, []),
    "eegclip_build_info": (C.c_char_p, []),
    "eegclip_launch_count": (C.c_longlong, []),
    "eegclip_tune_set": (C.c_int, [_i32, _i32]),
    "eegclip_debug_buffer": (C.c_int, [_vp]),
    "eegclip_profile_begin": (C.c_int, []),
    "eegclip_profile_end": (C.c_int, [_vp, _vp, _i32]),
    "eegclip_tower_workspace": (C.c_int, [C.POINTER(TowerDesc), _psz, _psz]),
    "eegclip_tower_forward": (C.c_int, [C.POINTER(TowerDesc), _vp, _vp, _vp, _vp, _vp, _vp]),
    "eegclip_tower_backward": (C.c_int, [C.POINTER(TowerDesc), _vp, _vp, _vp, _sz, _vp, _vp, _vp, _vp, _vp, _vp]),
    "eegclip_xfblock_workspace": (C.c_int, [C.POINTER(XfBlockDesc), _psz, _psz]),
    "eegclip_xfblock_forward": (C.c_int, [C.POINTER(XfBlockDesc), _vp, _vp, _vp, _vp, _vp, _vp]),
    "eegclip_xfblock_backward": (C.c_int, [C.POINTER(XfBlockDesc), _vp, _vp, _vp, _sz, _vp, _vp, _vp, _vp, _vp, _vp]),
    "eegclip_convblock_workspace": (C.c_int, [C.POINTER(ConvBlockDesc), _psz, _psz]),
    "eegclip_convblock_forward": (C.c_int, [C.POINTER(ConvBlockDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "eegclip_convblock_backward": (C.c_int, [C.POINTER(ConvBlockDesc)] + [_vp] * 14),
    "eegclip_linear_workspace": (C.c_int, [_i64, _i32, _i32, _psz]),
    "eegclip_linear_forward": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    "eegclip_linear_backward": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    "eegclip_bilstm_supported": (C.c_int, [C.POINTER(BiLstmDesc)]),
    "eegclip_bilstm_workspace": (C.c_int, [C.POINTER(BiLstmDesc), _psz, _psz]),
    "eegclip_bilstm_forward": (C.c_int, [C.POINTER(BiLstmDesc), _vp, _vp, _vp, _vp, _vp, _vp]),
    "eegclip_bilstm_backward": (C.c_int, [C.POINTER(BiLstmDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "eegclip_l2norm_forward": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _vp]),
    "eegclip_l2norm_backward": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _vp]),
    "eegclip_infonce_workspace": (C.c_int, [_i32, _i32, _i32, _psz]),
    "eegclip_infonce_lse": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _i32, _i32, _vp, _vp]),
    "eegclip_infonce_loss": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _vp, _vp]),
    "eegclip_infonce_backward": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp]),
    "eegclip_membank_update": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _i32, _i32, _f32, _f32, _vp]),
    "eegclip_adamw_step": (C.c_int, [_vp, _i32, _i64, _f32, _f32, _f32, _f32, _f32, _i32, _i64, _vp]),
    "eegclip_mm_rowdots": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "eegclip_mm_bank_workspace": (C.c_int, [_i32, _i32, _i32, _psz]),
    "eegclip_mm_bank_logits": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "eegclip_attention_forward": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _f32, _i32, _i32, C.c_uint64, _i32, _vp]),
    "eegclip_attention_backward": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _f32, _i32, _i32, C.c_uint64, _i32, _vp]),
    "eegclip_layernorm_forward": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "eegclip_layernorm_backward": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "eegclip_dropout": (C.c_int, [_vp, _vp, _i64, _f32, _i32, _i32, _i32, C.c_uint64, _vp]),
    "eegclip_gelu_dropout_forward": (C.c_int, [_vp, _vp, _i64, _f32, _i32, _i32, _i32, C.c_uint64, _vp]),
    "eegclip_gelu_dropout_backward": (C.c_int, [_vp, _vp, _vp, _i64, _f32, _i32, _i32, _i32, C.c_uint64, _vp]),
    "eegclip_mvn_workspace": (C.c_int, [_i32, _psz]),
    "eegclip_mvn_normalize": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "eegclip_row_topk": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _i64, _vp, _vp, _vp]),
    "eegclip_conv_small_workspace": (C.c_int, [_i32, _i32, _i32, _i32, _psz]),
    "eegclip_conv_small_forward": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "eegclip_conv_small_backward": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "eegclip_pearson_forward": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "eegclip_pearson_backward": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp]),
}

_ERR = {-1: "invalid argument", -2: "CUDA error", -3: "unsupported shape/configuration"}
_lib = None


def load():
    """Load the shared library (once). Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EegclipError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.eegclip_abi_version() != 4:
        raise EegclipError("libeegclip_b200.so ABI version mismatch")
    # development knobs from the environment: EEGCLIP_TUNE="7=1,6=1" -> eegclip_tune_set(7, 1), eegclip_tune_set(6, 1)
    for kv in filter(None, os.environ.get("EEGCLIP_TUNE", "").split(",")):
        k, v = kv.split("=")
        lib.eegclip_tune_set(int(k), int(v))
    _lib = lib
    return lib


def call(name, *args):
    rc = getattr(load(), name)(*args)
    if rc != 0:
        detail = ""
        if rc == -2 and torch.cuda.is_available():
            try:
                torch.cuda.synchronize()
            except Exception as e:  # surfaces the sticky CUDA error text
                detail = f" ({e})"
        raise EegclipError(f"{name} failed: {_ERR.get(rc, rc)}{detail}")


def ptr(t):
    """Device pointer of a contiguous fp32/int64 CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise EegclipError("eegclip_b200 kernels need CUDA tensors: there is no CPU fallback for this path")
    if not t.is_contiguous():
        raise EegclipError("non-contiguous tensor passed to an eegclip_b200 kernel")
    return t.data_ptr()


def f32c(t):
    """Detach, cast to fp32 and make contiguous (no copy when already so)."""
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def stream():
    return torch.cuda.current_stream().cuda_stream


def ptr_table(tensors):
    """Host int64 tensor holding the device pointers of ``tensors`` (kept alive by the caller)."""
    return torch.tensor([ptr(t) for t in tensors], dtype=torch.int64)


def new_seed():
    """Philox key for one forward/backward pair, drawn from torch's CPU generator (honours manual_seed)."""
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


# ---------------------------------------------------------------------------------------------------
# Gradient sinks: the optimizer (optim.AdamW) owns one flat gradient arena and registers, per parameter, the arena view that
# is its ``.grad``.  A backward of this package then writes the parameter gradients straight into those views and returns
# None for them, so autograd launches no ``grad += new`` kernel per parameter (229 launches per step in the default model).
# A view is handed out at most once per zero_grad() epoch; any second backward falls back to the ordinary accumulate path
# (autograd then adds into the same view in place, which is the correct accumulation).
# ---------------------------------------------------------------------------------------------------
import weakref

_SINKS = {}


def register_grad_sinks(params, views, arena):
    for p_, v in zip(params, views):
        _SINKS[p_.data_ptr()] = (weakref.ref(p_), v, arena)


def claim_grad_sinks(ps):
    """Arena views to write the gradients of ``ps`` into (in order), or None when any of them cannot be claimed.

    A view can be claimed once per zero_grad() epoch, and only while the parameter's ``.grad`` is None (set_to_none) or is
    that very view; the claim sets ``.grad`` to the view, so after backward the parameter looks exactly as if autograd
    had accumulated into it.  Parameters that are never claimed (nor reached by autograd) keep ``grad is None`` and the
    optimizer skips them, as torch.optim does."""
    found = []
    for p_ in ps:
        e = _SINKS.get(p_.data_ptr())
        if e is None:
            return None
        ref, v, arena = e
        owner = ref()
        if owner is None or owner.data_ptr() != p_.data_ptr() or v.shape != p_.shape or p_.data_ptr() in arena["written"]:
            return None
        if owner.grad is not None and owner.grad.data_ptr() != v.data_ptr():
            return None                                  # a foreign gradient is already accumulated there
        found.append((owner, v, arena))
    for p_, (owner, v, arena) in zip(ps, found):
        arena["written"].add(p_.data_ptr())
        owner.grad = v
    return [v for _, v, _ in found]


# listeners told that a backward of this package has just written a set of gradient sinks (data-parallel bucketed all-reduce)
_SINK_LISTENERS = []


def fire_sinks_written(views):
    for fn in list(_SINK_LISTENERS):
        fn(views)
