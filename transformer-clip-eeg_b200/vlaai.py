"""Drop-in mirror of the reference ``vlaai.py`` (VLAAI envelope-reconstruction baseline) on B200 kernels.

Same classes / ctor signatures / state_dict keys as /root/reference/vlaai.py:5-134.  Each
Conv1d(k=64,'same') -> LayerNorm([f,T]) -> LeakyReLU triple is one fused conv-block call
(eegclip_convblock_*), each 1x1 Conv1d one per-token linear (eegclip_linear_*); activations stay
time-major (B,T,C) throughout, the reference's (B,C,T) only appears at the module boundary.
"""
import torch
import torch.nn as nn

from . import _lib as L
from .clip_model import _ConvBlockFn, _LinearFn, _check_ln_shape, _require_cuda


BRANCH_TAP = None   # tests: a list that receives, per conv block call, the LeakyReLU branch mask (out > 0) in (B,C,T) layout


def _conv_ln_act(x, skip, conv, norm, training):
    B, T, cin = x.shape
    w = conv.weight
    _check_ln_shape(norm, w.shape[0], T, "VLAAI conv block")
    d = L.ConvBlockDesc(B=B, T=T, Cin=cin, Cout=w.shape[0], taps=w.shape[2], act=1, train=0, math=L.default_math(),
                        p_drop=0.0, layer=0, seed=0)
    out = _ConvBlockFn.apply(x, skip, d, w, conv.bias, norm.weight, norm.bias)
    if BRANCH_TAP is not None:
        BRANCH_TAP.append((out.detach() > 0).transpose(1, 2).cpu())
    return out


class Extractor(nn.Module):
    """vlaai.py:5-46."""

    def __init__(self, filters=(256, 256, 256, 128, 128), kernels=(64,) * 5, dilation_rate=1, input_channels=64,
                 time_dimension=64 * 5, normalization_fn='layer_norm', activation_fn='leaky_relu'):
        super().__init__()
        if len(filters) != len(kernels):
            raise ValueError("'filters' and 'kernels' must have the same length")
        if normalization_fn != 'layer_norm' or activation_fn != 'leaky_relu' or dilation_rate != 1:
            raise L.EegclipError("Extractor kernels cover layer_norm + leaky_relu, dilation 1 (the reference defaults)")
        self.eeg = nn.Conv1d(input_channels, input_channels, kernel_size=1)
        layers = []
        for f, k in zip(filters, kernels):
            layers += [nn.Conv1d(input_channels, f, kernel_size=k, padding='same', dilation=dilation_rate),
                       nn.LayerNorm([f, time_dimension]), nn.LeakyReLU()]
            input_channels = f
        self.conv_layers = nn.Sequential(*layers)

    def forward_time_major(self, x, skip=None):
        if skip is not None:
            x = x + skip
        x = _LinearFn.apply(x, self.eeg.weight, self.eeg.bias)
        for j in range(0, len(self.conv_layers), 3):
            x = _conv_ln_act(x, None, self.conv_layers[j], self.conv_layers[j + 1], self.training)
        return x

    def forward(self, x):
        _require_cuda(x, "Extractor")
        return self.forward_time_major(x.transpose(1, 2).contiguous()).transpose(1, 2)


class OutputContext(nn.Module):
    """vlaai.py:48-72."""

    def __init__(self, filter_=64, kernel=64, input_channels=64, time_dimension=64 * 5, normalization_fn='layer_norm',
                 activation_fn='leaky_relu'):
        super().__init__()
        self.conv1d = nn.Conv1d(input_channels, filter_, kernel_size=kernel, padding='same')
        self.normalization_fn = nn.LayerNorm([filter_, time_dimension])
        self.activation_fn = nn.LeakyReLU()

    def forward_time_major(self, x):
        return _conv_ln_act(x, None, self.conv1d, self.normalization_fn, self.training)

    def forward(self, x):
        _require_cuda(x, "OutputContext")
        return self.forward_time_major(x.transpose(1, 2).contiguous()).transpose(1, 2)


class VLAAI(nn.Module):
    """vlaai.py:74-134: the shared stack applied nb_blocks times (input skip on the inner iterations)."""

    def __init__(self, nb_blocks=4, extractor_model=None, output_context_model=None, use_skip=True, input_channels=64,
                 output_dim=64):
        super().__init__()
        extractor_model = extractor_model if extractor_model is not None else Extractor()
        output_context_model = output_context_model if output_context_model is not None else OutputContext()
        linear_recombination = nn.Conv1d(128, input_channels, kernel_size=1, padding='same')
        self.eeg = nn.Conv1d(input_channels, input_channels, kernel_size=1)
        self.use_skip = bool(use_skip)
        self.sequentialConvStack = nn.Sequential(extractor_model, linear_recombination, output_context_model)
        self.output_dim, self.nb_blocks = output_dim, nb_blocks
        self.final_linear = nn.Conv1d(input_channels, output_dim, kernel_size=1, padding='same')

    def get_output_dim(self, input_window_size):
        return input_window_size * self.output_dim

    def _stack(self, x, skip):
        ext, rec, ctx = self.sequentialConvStack
        x = ext.forward_time_major(x, skip)
        x = _LinearFn.apply(x, rec.weight, rec.bias)
        return ctx.forward_time_major(x)

    def forward(self, x):
        _require_cuda(x, "VLAAI")
        eeg = x.contiguous()                      # (B,T,64) time-major == reference x.transpose(1,2)
        skip = eeg if self.use_skip else None     # use_skip=False adds zeros in the reference (vlaai.py:118)
        x = _LinearFn.apply(eeg, self.eeg.weight, self.eeg.bias)
        for idx in range(self.nb_blocks):
            inner = not (idx == 0 or idx == self.nb_blocks - 1)
            x = self._stack(x, skip if inner else None)
        x = _LinearFn.apply(x, self.final_linear.weight, self.final_linear.bias)
        return x.transpose(1, 2)                  # reference returns (B, output_dim, T)
