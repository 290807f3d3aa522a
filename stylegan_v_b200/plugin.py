"""Tensor-level shims with the signatures of the reference's two pybind plugins.

  plugin.upfirdn2d(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain) -> Tensor
      stands in for `_plugin.upfirdn2d` (src/torch_utils/ops/upfirdn2d.cpp:16,98-101)
  plugin.bias_act(x, b, xref, yref, dy, grad, dim, act, alpha, gain, clamp) -> Tensor
      stands in for `_plugin.bias_act` (src/torch_utils/ops/bias_act.cpp:32,94-97); an empty tensor means "absent"

PyTorch is used only for what the reference's C++ side used ATen for: allocating the output on the
input's device / memory format, the device guard and the current stream.  Compute is libsgv_b200.
"""
import ctypes
import torch

from . import _lib

_DTYPES = {torch.float32: _lib.SGV_F32, torch.float16: _lib.SGV_F16, torch.float64: _lib.SGV_F64}
INT_MAX = 2 ** 31 - 1


def _require(cond, msg):
    if not cond:
        raise RuntimeError(msg)


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _suggest_format(x):
    # x.suggest_memory_format() of ATen: channels_last only when strides say so unambiguously
    if x.ndim == 4 and x.shape[1] > 1 and x.stride(1) == 1 and x.is_contiguous(memory_format=torch.channels_last) \
            and not x.is_contiguous():
        return torch.channels_last
    return torch.contiguous_format


def upfirdn2d(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain, *, epilogue=None):
    """One FIR resampling pass on the GPU.  `epilogue` (optional, not part of the reference plugin) is a dict
    with keys scale [N,C] / noise [N|1,1,oh,ow] / bias [C] / act ('linear'|'lrelu') / alpha / gain / clamp applied to the result."""
    _require(x.is_cuda, 'x must reside on CUDA device')
    _require(f.device == x.device, 'f must reside on the same device as x')
    _require(f.dtype == torch.float32, 'f must be float32')
    _require(x.numel() <= INT_MAX and f.numel() <= INT_MAX, 'x is too large')
    _require(x.ndim == 4, 'x must be rank 4')
    _require(f.ndim == 2, 'f must be rank 2')
    _require(f.shape[0] >= 1 and f.shape[1] >= 1, 'f must be at least 1x1')
    _require(upx >= 1 and upy >= 1, 'upsampling factor must be at least 1')
    _require(downx >= 1 and downy >= 1, 'downsampling factor must be at least 1')
    _require(x.dtype in _DTYPES, f'unsupported dtype {x.dtype}')
    L = _lib.lib()
    N, C, H, W = x.shape
    fh, fw = f.shape
    out_w = (W * upx + padx0 + padx1 - fw + downx) // downx if (W * upx + padx0 + padx1 - fw + downx) >= 0 else 0
    out_h = (H * upy + pady0 + pady1 - fh + downy) // downy if (H * upy + pady0 + pady1 - fh + downy) >= 0 else 0
    _require(out_w >= 1 and out_h >= 1, 'output must be at least 1x1')
    y = torch.empty([N, C, out_h, out_w], dtype=x.dtype, device=x.device, memory_format=_suggest_format(x))
    _require(y.numel() <= INT_MAX, 'output is too large')
    if y.numel() == 0:            # empty batch / no channels: nothing to launch (the reference's zero-sized grid raises instead)
        return y

    p = _lib.UpfirdnParams()
    p.x, p.f, p.y = x.data_ptr(), f.data_ptr(), y.data_ptr()
    p.dtype = _DTYPES[x.dtype]
    p.up_x, p.up_y, p.down_x, p.down_y = upx, upy, downx, downy
    p.pad_x0, p.pad_x1, p.pad_y0, p.pad_y1 = padx0, padx1, pady0, pady1
    p.flip, p.gain = int(bool(flip)), float(gain)
    p.in_w, p.in_h, p.in_c, p.in_n = W, H, C, N
    p.in_stride_x, p.in_stride_y, p.in_stride_c, p.in_stride_n = x.stride(3), x.stride(2), x.stride(1), x.stride(0)
    p.f_w, p.f_h, p.f_stride_x, p.f_stride_y = fw, fh, f.stride(1), f.stride(0)
    p.out_w, p.out_h = out_w, out_h
    p.out_stride_x, p.out_stride_y, p.out_stride_c, p.out_stride_n = y.stride(3), y.stride(2), y.stride(1), y.stride(0)
    keep = []
    if epilogue is not None:
        _require(x.dtype == torch.float32, 'fused epilogue is float32 only')
        sc, bi = epilogue.get('scale'), epilogue.get('bias')
        if sc is not None:
            sc = sc.to(torch.float32).contiguous(); keep.append(sc)
            _require(sc.shape == (N, C), 'epilogue scale must be [N, C]')
            p.epi_scale = sc.data_ptr()
        if bi is not None:
            bi = bi.to(torch.float32).contiguous(); keep.append(bi)
            _require(bi.shape == (C,), 'epilogue bias must be [C]')
            p.epi_bias = bi.data_ptr()
        p.epi_act = {'linear': 1, 'lrelu': 3}[epilogue.get('act', 'linear')]
        p.epi_alpha = float(epilogue.get('alpha', 0.2))
        p.epi_gain = float(epilogue.get('gain', 1.0))
        cl = epilogue.get('clamp')
        p.epi_clamp = float(cl) if cl is not None else -1.0
        p.epi_round_tf32 = int(bool(epilogue.get('round_tf32', False)))
        nz = epilogue.get('noise')
        if nz is not None:      # [N or 1, 1, out_h, out_w] (or [out_h, out_w]) plane added after the scale, before the bias
            nz = nz.to(torch.float32)
            nz = nz.reshape(1, out_h, out_w) if nz.ndim == 2 else nz.reshape(nz.shape[0], out_h, out_w)
            _require(nz.shape[0] in (1, N) and nz.device == x.device, 'epilogue noise must be [N or 1, 1, out_h, out_w] on the same device')
            keep.append(nz)
            p.epi_noise = nz.data_ptr()
            p.epi_noise_stride_n = nz.stride(0) if (nz.shape[0] == N and N > 1) else 0
            p.epi_noise_stride_y, p.epi_noise_stride_x = nz.stride(1), nz.stride(2)
    with torch.cuda.device(x.device):
        _lib.check(L.sgv_upfirdn2d(ctypes.byref(p), _stream_ptr(x.device)), 'sgv_upfirdn2d')
    return y


def _same_layout(a, b):
    if a.ndim != b.ndim:
        return False
    return all(a.shape[i] == b.shape[i] and (a.shape[i] < 2 or a.stride(i) == b.stride(i)) for i in range(a.ndim))


def _dense(x):
    if x.numel() <= 1:
        return True
    dims = sorted((d for d in range(x.ndim) if x.shape[d] > 1), key=lambda d: x.stride(d))
    expect = 1
    for d in dims:
        if x.stride(d) != expect:
            return False
        expect *= x.shape[d]
    return True


def bias_act(x, b, xref, yref, dy, grad, dim, act, alpha, gain, clamp, *, db_accum=None):
    """One fused bias/activation pass on the GPU.  `db_accum` (optional, not in the reference plugin):
    float32 [size(dim)] buffer into which the per-channel sum of the result is accumulated atomically."""
    _require(x.is_cuda, 'x must reside on CUDA device')
    for name, t in (('b', b), ('xref', xref), ('yref', yref), ('dy', dy)):
        _require(t.numel() == 0 or (t.dtype == x.dtype and t.device == x.device), f'{name} must have the same dtype and device as x')
    for name, t in (('xref', xref), ('yref', yref), ('dy', dy)):
        _require(t.numel() == 0 or t.shape == x.shape, f'{name} must have the same shape as x')
        _require(t.numel() == 0 or _same_layout(t, x), f'{name} must have the same layout as x')
    _require(x.numel() <= INT_MAX, 'x is too large')
    _require(b.ndim == 1, 'b must have rank 1')
    _require(b.numel() == 0 or (0 <= dim < x.ndim), 'dim is out of bounds')
    _require(b.numel() == 0 or b.numel() == x.shape[dim], 'b has wrong number of elements')
    _require(grad >= 0, 'grad must be non-negative')
    _require(_dense(x), 'x must be non-overlapping and dense')
    _require(b.is_contiguous(), 'b must be contiguous')
    _require(x.dtype in _DTYPES, f'unsupported dtype {x.dtype}')
    L = _lib.lib()
    y = torch.empty_like(x)
    _require(_same_layout(y, x), 'y must have the same layout as x')

    p = _lib.BiasActParams()
    opt = lambda t: t.data_ptr() if t.numel() else None
    p.x, p.b, p.xref, p.yref, p.dy, p.y = x.data_ptr(), opt(b), opt(xref), opt(yref), opt(dy), y.data_ptr()
    p.dtype = _DTYPES[x.dtype]
    p.grad, p.act = int(grad), int(act)
    p.alpha, p.gain, p.clamp = float(alpha), float(gain), float(clamp)
    p.size_x = x.numel()
    has_index = b.numel() > 0 or db_accum is not None
    p.size_b = (x.shape[dim] if has_index else 0)
    p.step_b = (x.stride(dim) if has_index and x.numel() else 1)
    if db_accum is not None:
        _require(db_accum.dtype == torch.float32 and db_accum.is_contiguous() and db_accum.numel() == x.shape[dim],
                 'db_accum must be a contiguous float32 [size(dim)] buffer')
        p.db_accum = db_accum.data_ptr()
    if x.numel() == 0:
        return y
    with torch.cuda.device(x.device):
        _lib.check(L.sgv_bias_act(ctypes.byref(p), _stream_ptr(x.device)), 'sgv_bias_act')
    return y
