"""The small dense layers of the path on the exact-fp32 kernels of libsgv_b200 (csrc/dense_f32.cu, include/sgv_b200_aux.h):

    linear            FullyConnectedLayer / EqualizedLinear (src/training/layers.py:108-138): mapping networks, discriminator dense layers,
                      time-encoder heads — bias, leaky ReLU and the gains run in the GEMM's epilogue, the activation gradient in the operand
                      load of the two gradient kernels
    stacked_affine    ALL style affines of a synthesis network in one launch (networks.py:124-126,159-160 evaluated for every layer up front):
                      column group g reads ws[:, g, :]
    conv1d_slabs      EqualizedConv1d (layers.py:331-373) as a GEMM over windows, evaluated only at the output positions the caller needs
                      (the motion encoder reads 2 of the 66 trajectory positions per frame, motion.py:105-115)

The reference runs these as torch.addmm / matmul (cuBLAS SIMT sgemm), cuDNN conv1d (FFT / TF32 engines) and separate bias_act passes.  Plain
fp32 FMAs with round-to-nearest accumulation on purpose: M = batch rows fill a quarter of one 128-row tensor-core tile, and the motion codes are
multiplied by phase scales up to 64 before sin / cos (motion.py:198-212) — the tcgen05 accumulator truncates, which showed up as 2e-3 in motion_v
(profiles/dense_precision_r2.txt).  First-order autograd nodes; CUDA / float32 only — anything else takes the PyTorch formulation in the calling
module, like the reference's ops do for CPU tensors.
"""
import ctypes

import torch

from . import _lib
from . import conv as _conv

_ACT = {'linear': 1, 'lrelu': 3}


def supported(x, weight, act='linear'):
    """CUDA fp32, reduction length a multiple of 4 (16-byte vector loads), linear or leaky-ReLU epilogue."""
    return (x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and act in _ACT and x.ndim == 2
            and weight.shape[-1] % 4 == 0 and x.shape[1] == weight.shape[-1])


def _ptr(t):
    return t.data_ptr() if t is not None else None


def _rows(x):
    """[M, K] with unit inner stride and a 16-byte addressable row stride (views like y[:, 0] of a [R, 2, C] tensor qualify); else a copy."""
    if x.stride(1) != 1 or x.stride(0) % 4 != 0 or x.data_ptr() % 16 != 0:
        x = x.contiguous()
    return x


def _params(m, n, k, w, w_gain, b_gain, act, gain, groups=None):
    p = _lib.DenseParams()
    p.m, p.n, p.k = int(m), int(n), int(k)
    p.w = w.data_ptr()
    p.w_gain, p.b_gain, p.act, p.alpha, p.gain = float(w_gain), float(b_gain), _ACT[act], 0.2, float(gain)
    if groups is not None:
        p.groups, p.group_col, p.group_off = int(groups[0]), groups[1].data_ptr(), groups[2].data_ptr()
    return p


def _call(name, p, device):
    with torch.cuda.device(device):
        _lib.check(getattr(_lib.lib(), name)(ctypes.byref(p), _conv._stream(device)), name)


def _forward(a, row_off, lda, m, w, bias, w_gain, b_gain, act, gain, groups=None):
    n, k = w.shape
    y = torch.empty([m, n], dtype=torch.float32, device=w.device)
    p = _params(m, n, k, w, w_gain, b_gain, act, gain, groups)
    p.a, p.a_row_off, p.lda, p.bias, p.y, p.ldy = a.data_ptr(), _ptr(row_off), int(lda), _ptr(bias), y.data_ptr(), n
    _call('sgv_dense_f32_fwd', p, w.device)
    return y


def _dgrad(dy, y, w, m, da, ldda, w_gain, act, gain, groups=None):
    """da (zero-filled by the caller, row stride ldda) += w_gain * dz @ w."""
    n, k = w.shape
    p = _params(m, n, k, w, w_gain, 1.0, act, gain, groups)
    p.dy, p.lddy, p.y, p.ldy, p.da, p.ldda = dy.data_ptr(), dy.stride(0), _ptr(y), n, da.data_ptr(), int(ldda)
    _call('sgv_dense_f32_dgrad', p, w.device)
    return da


def _wgrad(dy, y, a, row_off, lda, m, w, w_gain, b_gain, act, gain, want_db, groups=None):
    n, k = w.shape
    dw = torch.empty([n, k], dtype=torch.float32, device=w.device)
    db = torch.empty([n], dtype=torch.float32, device=w.device) if want_db else None
    p = _params(m, n, k, w, w_gain, b_gain, act, gain, groups)
    p.dy, p.lddy, p.y, p.ldy = dy.data_ptr(), dy.stride(0), _ptr(y), n
    p.a, p.a_row_off, p.lda, p.dw, p.db = a.data_ptr(), _ptr(row_off), int(lda), dw.data_ptr(), _ptr(db)
    _call('sgv_dense_f32_wgrad', p, w.device)
    return dw, db


def _grad_rows(dy):
    if dy.stride(1) != 1:
        dy = dy.contiguous()
    return dy


class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, w_gain, b_gain, act, gain):
        x = _rows(x)
        y = _forward(x, None, x.stride(0), x.shape[0], weight, bias, w_gain, b_gain, act, gain)
        ctx.save_for_backward(x, weight, y if act != 'linear' else x.new_empty(0))
        ctx.cfg = (w_gain, b_gain, act, gain, bias is not None)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        w_gain, b_gain, act, gain, has_b = ctx.cfg
        dy = _grad_rows(dy)
        y = y if act != 'linear' else None
        M, K = x.shape
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = _dgrad(dy, y, weight, M, torch.zeros([M, K], dtype=torch.float32, device=x.device), K, w_gain, act, gain)
        if ctx.needs_input_grad[1] or (has_b and ctx.needs_input_grad[2]):
            dw, db = _wgrad(dy, y, x, None, x.stride(0), M, weight, w_gain, b_gain, act, gain, has_b and ctx.needs_input_grad[2])
        return dx, dw, db, None, None, None, None


def linear(x, weight, bias=None, weight_gain=1.0, bias_gain=1.0, act='linear', gain=1.0):
    """act(x @ (weight * weight_gain).T + bias * bias_gain) * gain for x [M, K], weight [O, K]; act in {'linear', 'lrelu'} (bias_act's
    default sqrt(2) for lrelu is the caller's `gain`).  Caller checks supported()."""
    assert act in _ACT and x.ndim == 2
    return _Linear.apply(x, weight.contiguous(), bias, float(weight_gain), float(bias_gain), act, float(gain))


class _StackedAffine(torch.autograd.Function):
    """styles[m, col] = w_gain * sum_k ws[m, g(col), k] * wcat[col, k] + bcat[col]   (g(col) from the group tables)."""

    @staticmethod
    def forward(ctx, ws, wcat, bcat, groups, w_gain):
        ws = ws.contiguous()
        M, G, K = ws.shape
        y = _forward(ws, None, G * K, M, wcat, bcat, w_gain, 1.0, 'linear', 1.0, groups)
        ctx.save_for_backward(ws, wcat)
        ctx.cfg = (groups, w_gain)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        ws, wcat = ctx.saved_tensors
        groups, w_gain = ctx.cfg
        dy = _grad_rows(dy)
        M, G, K = ws.shape
        dws = dw = db = None
        if ctx.needs_input_grad[0]:
            dws = _dgrad(dy, None, wcat, M, torch.zeros([M, G, K], dtype=torch.float32, device=ws.device), G * K, w_gain, 'linear', 1.0, groups)
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            dw, db = _wgrad(dy, None, ws, None, G * K, M, wcat, w_gain, 1.0, 'linear', 1.0, ctx.needs_input_grad[2], groups)
        return dws, dw, db, None, None


def make_groups(col_begin, w_index, k, device):
    """Group tables for stacked_affine: columns [col_begin[i], col_begin[i+1]) read ws[:, w_index[i], :]; col_begin ascending multiples of 8."""
    assert len(col_begin) == len(w_index) + 1 and all(c % 8 == 0 for c in col_begin)
    col = torch.tensor(col_begin, dtype=torch.int32, device=device)
    off = torch.tensor([int(i) * k for i in w_index], dtype=torch.int64, device=device)
    return (len(w_index), col, off)


def stacked_affine(ws, wcat, bcat, groups, weight_gain):
    """ws [M, num_ws, K], wcat [N, K], bcat [N] -> [M, N]; `groups` from make_groups."""
    return _StackedAffine.apply(ws, wcat, bcat, groups, float(weight_gain))


class _Conv1dSlabs(torch.autograd.Function):
    """y[g, p, o] = act(w_gain * sum_{j, c} src(g, p + j, c) * w[o, j * C + c] + b_gain * b[o]) for p < P: a valid conv1d evaluated on slabs of
    P + k - 1 consecutive positions.  src is either a [B, L, C] sequence with per-slab element offsets `base` (no gradient to src), or — with
    base None — the slabs themselves, [G, P + k - 1, C] (gradient by overlap-add)."""

    @staticmethod
    def forward(ctx, src, base, P, weight_r, bias, w_gain, b_gain, act, gain):
        src = src.contiguous()
        C = src.shape[-1]
        O, KC = weight_r.shape
        k = KC // C
        if base is None:
            G = src.shape[0]
            assert src.shape[1] == P + k - 1
            base = torch.arange(G, device=src.device, dtype=torch.int64) * ((P + k - 1) * C)
            ctx.slab_src = True
        else:
            G = base.numel()
            ctx.slab_src = False
        row_off = (base.reshape(G, 1) + torch.arange(P, device=src.device, dtype=torch.int64).reshape(1, P) * C).reshape(-1).contiguous()
        y = _forward(src, row_off, 0, G * P, weight_r, bias, w_gain, b_gain, act, gain)
        ctx.save_for_backward(src, row_off, weight_r, y if act != 'linear' else src.new_empty(0))
        ctx.cfg = (G, P, k, C, w_gain, b_gain, act, gain, bias is not None)
        return y.reshape(G, P, O)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        src, row_off, weight_r, y = ctx.saved_tensors
        G, P, k, C, w_gain, b_gain, act, gain, has_b = ctx.cfg
        O = weight_r.shape[0]
        dy = dy.reshape(G * P, O).contiguous()
        y = y if act != 'linear' else None
        dsrc = dw = db = None
        if ctx.needs_input_grad[0]:
            assert ctx.slab_src, 'no gradient path to a gathered source sequence'
            da = _dgrad(dy, y, weight_r, G * P, torch.zeros([G * P, k * C], dtype=torch.float32, device=src.device), k * C, w_gain, act, gain)
            da = da.reshape(G, P, k, C)
            dsrc = torch.zeros([G, P + k - 1, C], dtype=torch.float32, device=src.device)
            if P <= k:
                for p in range(P):
                    dsrc[:, p:p + k] += da[:, p]
            else:
                for j in range(k):
                    dsrc[:, j:j + P] += da[:, :, j]
        if ctx.needs_input_grad[3] or (has_b and ctx.needs_input_grad[4]):
            dw, db = _wgrad(dy, y, src, row_off, 0, G * P, weight_r, w_gain, b_gain, act, gain, has_b and ctx.needs_input_grad[4])
        return dsrc, None, None, dw, db, None, None, None, None


def conv1d_slabs(src, base, P, weight, bias=None, weight_gain=1.0, bias_gain=1.0, act='linear', gain=1.0):
    """weight [O, C, k] (the conv1d parameter layout); see _Conv1dSlabs.  Returns [G, P, O]."""
    O, C, k = weight.shape
    weight_r = weight.permute(0, 2, 1).reshape(O, k * C)          # tap-major rows match the contiguous [k, C] window of the sequence
    return _Conv1dSlabs.apply(src, base, int(P), weight_r, bias, float(weight_gain), float(bias_gain), act, float(gain))
