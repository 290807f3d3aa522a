"""Arithmetic mode of the tensor-core contractions (include/sgv_b200_conv.h).

  'tf32'    (default) TF32 x TF32 products, fp32 accumulation: <= 1e-3 of fp32 per contraction (north_star tolerance).
  'tf32x3'  fp32-grade: operands split into hi + lo TF32 parts inside the kernels, hi*hi + lo*hi + hi*lo accumulated in fp32
            (~1e-6 of fp32).  This is the mode that corresponds to the reference's `torch.backends.cudnn.allow_tf32 = False`
            (src/training/training_loop.py:141-142); it costs about three times the tensor-core work.

The mode is read when a forward op runs and recorded in its autograd context, so a backward pass always uses the mode of its
forward pass, whatever thread it runs on.  `SGV_PRECISION` sets the process default.
"""
import contextlib
import os

MODES = ('tf32', 'tf32x3')
_mode = os.environ.get('SGV_PRECISION', 'tf32')
assert _mode in MODES, f'SGV_PRECISION must be one of {MODES}'


def get():
    return _mode


def is_x3(mode=None):
    return (mode if mode is not None else _mode) == 'tf32x3'


def set_precision(mode):
    global _mode
    assert mode in MODES, f'precision must be one of {MODES}'
    _mode = mode


@contextlib.contextmanager
def precision(mode):
    prev = get()
    set_precision(mode)
    try:
        yield
    finally:
        set_precision(prev)
