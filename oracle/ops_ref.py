"""ORACLE (test infrastructure only — see oracle/__init__.py).

CPU restatement of the reference's `src/torch_utils/ops` layer for the hot path:

  upfirdn2d_ref            src/torch_utils/ops/upfirdn2d.cu:29-92 + upfirdn2d.cpp:32-33 (via oracle/c)
  upfirdn2d_ref_torch      src/torch_utils/ops/upfirdn2d.py:169-208 (_upfirdn2d_ref: zero-insert, pad/crop, conv, decimate)
  bias_act_ref             src/torch_utils/ops/bias_act.cu:23-147 (via oracle/c)  /  bias_act.py:94-123
  conv2d_resample_ref      src/torch_utils/ops/conv2d_resample.py:59-154 (padding arithmetic 95-104, fast paths 106-154)
  modulated_conv2d_ref     src/training/networks.py:30-86 (non-fused "train" path 64-74, fused path 76-86)

Dense contractions use torch.nn.functional.conv2d / conv_transpose2d on CPU in fp32 or fp64: that
is third-party arithmetic for the reference as well (cuDNN / oneDNN via torch, pinned only by
environment.yaml:8-10 `pytorch 1.7.1`), fully specified by torch's documented definition.
"""
import ctypes
import numpy as np
import torch
import torch.nn.functional as F

from . import build as _build

_lib = None


class _Geom(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ('up_x', 'up_y', 'down_x', 'down_y', 'pad_x0', 'pad_y0', 'flip',
                                             'in_w', 'in_h', 'in_c', 'in_n')] + \
               [(n, ctypes.c_int64) for n in ('in_sx', 'in_sy', 'in_sc', 'in_sn')] + \
               [('f_w', ctypes.c_int), ('f_h', ctypes.c_int), ('f_sx', ctypes.c_int64), ('f_sy', ctypes.c_int64),
                ('out_w', ctypes.c_int), ('out_h', ctypes.c_int)] + \
               [(n, ctypes.c_int64) for n in ('out_sx', 'out_sy', 'out_sc', 'out_sn')]


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
        _lib.oracle_upfirdn2d_out_size.restype = ctypes.c_int
    return _lib


# ----------------------------------------------------------------------------------------------
# argument parsing (upfirdn2d.py:37-68)

def parse_scaling(s):
    if isinstance(s, int):
        s = [s, s]
    sx, sy = s
    assert sx >= 1 and sy >= 1
    return int(sx), int(sy)


def parse_padding(p):
    if isinstance(p, int):
        p = [p, p]
    p = list(p)
    if len(p) == 2:
        p = [p[0], p[0], p[1], p[1]]
    px0, px1, py0, py1 = p
    return int(px0), int(px1), int(py0), int(py1)


def filter_size(f):
    if f is None:
        return 1, 1
    return int(f.shape[-1]), int(f.shape[0])


def setup_filter(f, normalize=True, flip_filter=False, gain=1, separable=None):
    """upfirdn2d.py:72-116."""
    if f is None:
        f = 1
    f = torch.as_tensor(f, dtype=torch.float32)
    if f.ndim == 0:
        f = f[None]
    if separable is None:
        separable = (f.ndim == 1 and f.numel() >= 8)
    if f.ndim == 1 and not separable:
        f = torch.outer(f, f)
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f.flip(list(range(f.ndim)))
    return f * (gain ** (f.ndim / 2))


# ----------------------------------------------------------------------------------------------
# upfirdn2d

def _upfirdn2d_pass(x, f2d, upx, upy, downx, downy, px0, px1, py0, py1, flip, gain):
    """One plugin call (upfirdn2d.cpp:16-94) on a torch CPU tensor, any strides, f32/f64."""
    assert x.ndim == 4 and f2d.ndim == 2 and f2d.dtype == torch.float32
    L = lib()
    N, C, H, W = x.shape
    fh, fw = f2d.shape
    ow = L.oracle_upfirdn2d_out_size(W, upx, px0, px1, fw, downx)
    oh = L.oracle_upfirdn2d_out_size(H, upy, py0, py1, fh, downy)
    assert ow >= 1 and oh >= 1
    cl = x.ndim == 4 and x.stride(1) == 1 and C > 1
    y = torch.empty([N, C, oh, ow], dtype=x.dtype).contiguous(memory_format=torch.channels_last if cl else torch.contiguous_format)
    f2d = f2d.contiguous()
    g = _Geom(upx, upy, downx, downy, px0, py0, int(bool(flip)), W, H, C, N,
              x.stride(3), x.stride(2), x.stride(1), x.stride(0),
              fw, fh, f2d.stride(1), f2d.stride(0), ow, oh,
              y.stride(3), y.stride(2), y.stride(1), y.stride(0))
    fn = {torch.float32: L.oracle_upfirdn2d_f32, torch.float64: L.oracle_upfirdn2d_f64}[x.dtype]
    rc = fn(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(f2d.data_ptr()), ctypes.c_void_p(y.data_ptr()),
            ctypes.byref(g), ctypes.c_double(float(gain)))
    assert rc == 0
    return y


def upfirdn2d_ref(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1):
    """Scalar port of the CUDA kernel; separable filters take two passes like upfirdn2d.py:236-240."""
    upx, upy = parse_scaling(up)
    downx, downy = parse_scaling(down)
    px0, px1, py0, py1 = parse_padding(padding)
    if f is None:
        f = torch.ones([1, 1], dtype=torch.float32)
    x = x.detach()
    if f.ndim == 2:
        return _upfirdn2d_pass(x, f, upx, upy, downx, downy, px0, px1, py0, py1, flip_filter, gain)
    y = _upfirdn2d_pass(x, f.unsqueeze(0), upx, 1, downx, 1, px0, px1, 0, 0, flip_filter, np.sqrt(gain))
    return _upfirdn2d_pass(y, f.unsqueeze(1), 1, upy, 1, downy, 0, 0, py0, py1, flip_filter, np.sqrt(gain))


def upfirdn2d_ref_torch(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1):
    """Differentiable restatement with standard torch ops (upfirdn2d.py:169-208)."""
    upx, upy = parse_scaling(up)
    downx, downy = parse_scaling(down)
    px0, px1, py0, py1 = parse_padding(padding)
    if f is None:
        f = torch.ones([1, 1], dtype=torch.float32)
    N, C, H, W = x.shape
    # zero insertion
    z = x.new_zeros([N, C, H, upy, W, upx])
    z[:, :, :, 0, :, 0] = x
    z = z.reshape(N, C, H * upy, W * upx)
    # pad / crop
    z = F.pad(z, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    z = z[:, :, max(-py0, 0): z.shape[2] - max(-py1, 0), max(-px0, 0): z.shape[3] - max(-px1, 0)]
    k = (f * (gain ** (f.ndim / 2))).to(x.dtype)
    if not flip_filter:
        k = k.flip(list(range(k.ndim)))
    if k.ndim == 2:
        z = F.conv2d(z, k[None, None].repeat(C, 1, 1, 1), groups=C)
    else:
        z = F.conv2d(z, k[None, None, None, :].repeat(C, 1, 1, 1), groups=C)
        z = F.conv2d(z, k[None, None, :, None].repeat(C, 1, 1, 1), groups=C)
    return z[:, :, ::downy, ::downx]


def upfirdn2d_backward_padding(x_shape, dy_shape, f, up, down, padding):
    """Padding of the gradient pass (upfirdn2d.py:246-256)."""
    upx, upy = parse_scaling(up)
    downx, downy = parse_scaling(down)
    px0, _, py0, _ = parse_padding(padding)
    _, _, ih, iw = x_shape
    _, _, oh, ow = dy_shape
    fw, fh = filter_size(f)
    return [fw - px0 - 1, iw * upx - ow * downx + px0 - upx + 1,
            fh - py0 - 1, ih * upy - oh * downy + py0 - upy + 1]


# ----------------------------------------------------------------------------------------------
# bias_act

ACTS = {  # name: (cuda_idx, def_alpha, def_gain, ref, has_2nd_grad)   bias_act.py:23-33
    'linear': (1, 0.0, 1.0, '', False),
    'relu': (2, 0.0, float(np.sqrt(2)), 'y', False),
    'lrelu': (3, 0.2, float(np.sqrt(2)), 'y', False),
    'tanh': (4, 0.0, 1.0, 'y', True),
    'sigmoid': (5, 0.0, 1.0, 'y', True),
    'elu': (6, 0.0, 1.0, 'y', True),
    'selu': (7, 0.0, 1.0, 'y', True),
    'softplus': (8, 0.0, 1.0, 'y', True),
    'swish': (9, 0.0, float(np.sqrt(2)), 'x', True),
}


def bias_act_kernel_ref(x, b=None, xref=None, yref=None, dy=None, grad=0, dim=1, act='linear', alpha=None, gain=None, clamp=None):
    """One plugin call (bias_act.cpp:32-90) on a dense torch CPU tensor (f32/f64)."""
    idx, def_alpha, def_gain, _, _ = ACTS[act]
    alpha = float(def_alpha if alpha is None else alpha)
    gain = float(def_gain if gain is None else gain)
    clamp = float(-1 if clamp is None else clamp)
    assert x.is_contiguous() or x.is_contiguous(memory_format=torch.channels_last)
    y = torch.empty_like(x)
    assert y.stride() == x.stride()
    L = lib()
    fn = {torch.float32: L.oracle_bias_act_f32, torch.float64: L.oracle_bias_act_f64}[x.dtype]

    def ptr(t):
        if t is None:
            return ctypes.c_void_p(0)
        assert t.dtype == x.dtype
        if t.ndim == x.ndim:
            assert t.stride() == x.stride() or t.numel() <= 1
        return ctypes.c_void_p(t.data_ptr())
    if b is not None:
        b = b.contiguous()
    rc = fn(ptr(x), ptr(b), ptr(xref), ptr(yref), ptr(dy), ptr(y), int(grad), idx,
            ctypes.c_double(alpha), ctypes.c_double(gain), ctypes.c_double(clamp),
            ctypes.c_int64(x.numel()), int(b.numel()) if b is not None else 0,
            ctypes.c_int64(x.stride(dim) if b is not None else 1))
    assert rc == 0
    return y


_ACT_FUNCS = {
    'linear': lambda x, alpha: x,
    'relu': lambda x, alpha: F.relu(x),
    'lrelu': lambda x, alpha: F.leaky_relu(x, alpha),
    'tanh': lambda x, alpha: torch.tanh(x),
    'sigmoid': lambda x, alpha: torch.sigmoid(x),
    'elu': lambda x, alpha: F.elu(x),
    'selu': lambda x, alpha: F.selu(x),
    'softplus': lambda x, alpha: F.softplus(x),
    'swish': lambda x, alpha: torch.sigmoid(x) * x,
}


def bias_act_ref_torch(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None):
    """Differentiable restatement with torch ops (bias_act.py:94-123)."""
    _, def_alpha, def_gain, _, _ = ACTS[act]
    alpha = float(def_alpha if alpha is None else alpha)
    gain = float(def_gain if gain is None else gain)
    clamp = float(-1 if clamp is None else clamp)
    if b is not None:
        x = x + b.reshape([-1 if i == dim else 1 for i in range(x.ndim)])
    x = _ACT_FUNCS[act](x, alpha)
    if gain != 1:
        x = x * gain
    if clamp >= 0:
        x = x.clamp(-clamp, clamp)
    return x


# ----------------------------------------------------------------------------------------------
# conv2d_resample (conv2d_resample.py:59-154) — differentiable, torch CPU ops

def _conv(x, w, stride=1, padding=0, groups=1, transpose=False, flip_weight=True):
    if not flip_weight:   # F.conv2d is a correlation; "flip_weight=False" means true convolution (conv2d_resample.py:35-36)
        w = w.flip([2, 3])
    op = F.conv_transpose2d if transpose else F.conv2d
    return op(x, w, stride=stride, padding=padding, groups=groups)


def conv2d_resample_ref(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False,
                        upfirdn=upfirdn2d_ref_torch):
    out_channels, in_per_group, kh, kw = w.shape
    fw, fh = filter_size(f)
    px0, px1, py0, py1 = parse_padding(padding)
    if up > 1:      # conv2d_resample.py:95-99
        px0 += (fw + up - 1) // 2
        px1 += (fw - up) // 2
        py0 += (fh + up - 1) // 2
        py1 += (fh - up) // 2
    if down > 1:    # conv2d_resample.py:100-104
        px0 += (fw - down + 1) // 2
        px1 += (fw - down) // 2
        py0 += (fh - down + 1) // 2
        py1 += (fh - down) // 2
    if kw == 1 and kh == 1 and down > 1 and up == 1:         # :107-110
        x = upfirdn(x, f, down=down, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return _conv(x, w, groups=groups, flip_weight=flip_weight)
    if kw == 1 and kh == 1 and up > 1 and down == 1:         # :113-116
        x = _conv(x, w, groups=groups, flip_weight=flip_weight)
        return upfirdn(x, f, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    if down > 1 and up == 1:                                  # :119-122
        x = upfirdn(x, f, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return _conv(x, w, stride=down, groups=groups, flip_weight=flip_weight)
    if up > 1:                                                # :125-142
        if groups == 1:
            w = w.transpose(0, 1)
        else:
            w = w.reshape(groups, out_channels // groups, in_per_group, kh, kw).transpose(1, 2)
            w = w.reshape(groups * in_per_group, out_channels // groups, kh, kw)
        px0 -= kw - 1
        px1 -= kw - up
        py0 -= kh - 1
        py1 -= kh - up
        pxt = max(min(-px0, -px1), 0)
        pyt = max(min(-py0, -py1), 0)
        x = _conv(x, w, stride=up, padding=[pyt, pxt], groups=groups, transpose=True, flip_weight=(not flip_weight))
        x = upfirdn(x, f, padding=[px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt], gain=up ** 2, flip_filter=flip_filter)
        if down > 1:
            x = upfirdn(x, f, down=down, flip_filter=flip_filter)
        return x
    if up == 1 and down == 1 and px0 == px1 and py0 == py1 and px0 >= 0 and py0 >= 0:   # :145-147
        return _conv(x, w, padding=[py0, px0], groups=groups, flip_weight=flip_weight)
    x = upfirdn(x, (f if up > 1 else None), up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    x = _conv(x, w, groups=groups, flip_weight=flip_weight)
    if down > 1:
        x = upfirdn(x, f, down=down, flip_filter=flip_filter)
    return x


# ----------------------------------------------------------------------------------------------
# modulated_conv2d (networks.py:30-86) — differentiable, torch CPU ops

def modulated_conv2d_ref(x, weight, styles, noise=None, up=1, down=1, padding=0, resample_filter=None,
                         demodulate=True, flip_weight=True, fused_modconv=True):
    N = x.shape[0]
    O, I, kh, kw = weight.shape
    w = None
    dcoefs = None
    if demodulate or fused_modconv:
        w = weight.unsqueeze(0) * styles.reshape(N, 1, I, 1, 1)             # :57
    if demodulate:
        dcoefs = (w.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt()              # :59
    if demodulate and fused_modconv:
        w = w * dcoefs.reshape(N, O, 1, 1, 1)                               # :61
    if not fused_modconv:                                                   # :64-74
        x = x * styles.reshape(N, I, 1, 1)
        x = conv2d_resample_ref(x, weight, f=resample_filter, up=up, down=down, padding=padding, flip_weight=flip_weight)
        if demodulate and noise is not None:
            x = torch.addcmul(noise, x, dcoefs.reshape(N, O, 1, 1))
        elif demodulate:
            x = x * dcoefs.reshape(N, O, 1, 1)
        elif noise is not None:
            x = x + noise
        return x
    x = x.reshape(1, N * I, *x.shape[2:])                                   # :79-86
    w = w.reshape(N * O, I, kh, kw)
    x = conv2d_resample_ref(x, w, f=resample_filter, up=up, down=down, padding=padding, groups=N, flip_weight=flip_weight)
    x = x.reshape(N, O, *x.shape[2:])
    if noise is not None:
        x = x + noise
    return x
