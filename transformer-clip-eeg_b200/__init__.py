"""B200-native EEG-CLIP hot path (package directory ``transformer-clip-eeg_b200``).

The directory name is not a Python identifier; import it as ``transformer_clip_eeg_b200`` through the
repo-root shim, or put this directory's parent on ``sys.path`` and call :func:`install` to expose the
modules under the reference's own names (``clip_model``, ``vlaai``, ``train_clip_helper_functions``),
which is the drop-in: unmodified reference scripts doing ``from clip_model import *`` then run on the
B200 kernels.
"""
import sys

from . import _lib  # noqa: F401  (ctypes binding; raises if the shared library is missing when first used)
from ._lib import EegclipError, set_default_math, default_math  # noqa: F401


def build(force=False, verbose=False):
    from . import _build
    return _build.build(force=force, verbose=verbose)


def install():
    """Register this package's modules under the reference's top-level module names."""
    from . import clip_model, vlaai, train_clip_helper_functions
    sys.modules["clip_model"] = clip_model
    sys.modules["vlaai"] = vlaai
    sys.modules["train_clip_helper_functions"] = train_clip_helper_functions
    return clip_model
