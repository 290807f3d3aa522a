"""Tensor-level interface to the tcgen05 implicit-GEMM convolution of libsgv_b200 (include/sgv_b200_conv.h).

Activations are NHWC in memory = torch tensors of logical shape [N, C, H, W] with channels_last strides.
`igemm_conv` is the raw kernel call; the autograd-aware layers that use it live in stylegan_v_b200/synthesis.py.
"""
import ctypes
import torch

from . import _lib
from . import precision as _precision


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _req(cond, msg):
    if not cond:
        raise RuntimeError(msg)


def prep_weights(w, taps, rows_dim=0, cols_dim=1, scale=1.0, x3=None):
    """[rows, cols, kh, kw]-like weight (any strides) -> TF32 slabs for the kernel: [ntaps, rows, cols], or in tf32x3 mode
    [2, ntaps, rows, cols] (hi parts, then the TF32 residuals).  `scale` is multiplied in before rounding (equalised-lr gain).

    taps: list of (ky, kx) kernel positions; rows_dim / cols_dim say which weight dims play GEMM-N (output
    channels of the contraction) and GEMM-K (its input channels).  x3=None follows stylegan_v_b200.precision."""
    _req(w.is_cuda and w.dtype == torch.float32 and w.ndim == 4, 'weight must be a CUDA float32 4-D tensor')
    L = _lib.lib()
    x3 = _precision.is_x3() if x3 is None else bool(x3)
    rows, cols = w.shape[rows_dim], w.shape[cols_dim]
    nt = len(taps)
    wp = torch.empty([2, nt, rows, cols] if x3 else [nt, rows, cols], dtype=torch.float32, device=w.device)
    ky = (ctypes.c_int32 * nt)(*[int(t[0]) for t in taps])
    kx = (ctypes.c_int32 * nt)(*[int(t[1]) for t in taps])
    with torch.cuda.device(w.device):
        _lib.check(L.sgv_conv_prep_weights_ex(w.data_ptr(), w.stride(rows_dim), w.stride(cols_dim), w.stride(2), w.stride(3),
                                              rows, cols, nt, ky, kx, float(scale), wp.data_ptr(), wp[1].data_ptr() if x3 else None,
                                              _stream(w.device)), 'sgv_conv_prep_weights_ex')
    return wp


def prep_weights_pair(w, taps_fwd, taps_dgrad, x3=None):
    """Dense [O, I, kh, kw] weight -> (wp_fwd [len(taps_fwd), O, I], wp_dgrad [len(taps_dgrad), I, O]), TF32-rounded, ONE launch
    (sgv_conv_prep_weights_pair[_x3]); in tf32x3 mode each gets a leading [2] = (hi, residual).
    Equivalent to prep_weights(w, taps_fwd) and prep_weights(w, taps_dgrad, rows_dim=1, cols_dim=0)."""
    _req(w.is_cuda and w.dtype == torch.float32 and w.ndim == 4 and w.is_contiguous(), 'weight must be a dense CUDA float32 [O, I, kh, kw] tensor')
    x3 = _precision.is_x3() if x3 is None else bool(x3)
    O, I, kh, kw = w.shape
    na, nb = len(taps_fwd), len(taps_dgrad)
    wa = torch.empty(([2] if x3 else []) + [na, O, I], dtype=torch.float32, device=w.device)
    wb = torch.empty(([2] if x3 else []) + [nb, I, O], dtype=torch.float32, device=w.device)
    arr = lambda taps, j: (ctypes.c_int32 * max(len(taps), 1))(*[int(t[j]) for t in taps])
    L = _lib.lib()
    with torch.cuda.device(w.device):
        _lib.check(L.sgv_conv_prep_weights_pair_x3(w.data_ptr(), O, I, kh, kw, na, arr(taps_fwd, 0), arr(taps_fwd, 1), wa.data_ptr(),
                                                   wa[1].data_ptr() if x3 else None, nb, arr(taps_dgrad, 0), arr(taps_dgrad, 1), wb.data_ptr(),
                                                   wb[1].data_ptr() if x3 else None, _stream(w.device)), 'sgv_conv_prep_weights_pair_x3')
    return wa, wb


def slab_taps(wp, start, stop):
    """Taps [start, stop) of a slab tensor from prep_weights / prep_weights_pair (tap axis = -3), in either precision layout."""
    return wp[..., start:stop, :, :]


def igemm_conv(x, wp, tap_offsets, out=None, out_hw=None, out_view=None, in_stride=1,
               a_scale=None, o_scale=None, bias=None, act='linear', alpha=0.2, gain=1.0, clamp=None, accumulate=False,
               red_x=None, red_out=None, a_ready=False, noise=None, query=False):
    """y[n,oy,ox,o] = epi(sum_{t,i} x[n, oy*in_stride+dy_t, ox*in_stride+dx_t, i] * a_scale[n,i] * wp[t,o,i]).

    x: [N, Cin, H, W] channels_last fp32.  wp: [ntaps, Cout, Cin] from prep_weights — or [2, ntaps, Cout, Cin] (hi, residual),
    which selects the fp32-grade tf32x3 arithmetic.  tap_offsets: [(dy, dx)].
    noise: [N or 1, 1, out_h, out_w] (or [out_h, out_w]) plane added after o_scale, before the bias (modulated_conv2d's noise-add).
    Output: a new channels_last [N, Cout, out_h, out_w] tensor, or — for polyphase writes — `out_view`, a strided
    view [N, Cout, out_h, out_w] (channel stride 1) of a larger channels_last tensor.
    query=True: nothing is launched; returns the kernel variant the call would run (sgv_conv2d_tf32_variant) as a dict."""
    _req(x.is_cuda and x.dtype == torch.float32 and x.ndim == 4, 'x must be a CUDA float32 [N,C,H,W] tensor')
    N, Cin, H, W = x.shape
    dense = x.stride(1) == 1 and x.stride(3) == Cin and x.stride(2) == W * Cin and x.stride(0) == H * W * Cin
    _req(dense or (x.stride(1) == 1 and all(st % 4 == 0 for st in (x.stride(0), x.stride(2), x.stride(3)))),
         'x must be channels_last (NHWC), dense or a pixel-strided view with unit channel stride')
    x3 = wp.ndim == 4
    _req(wp.ndim in (3, 4) and (not x3 or wp.shape[0] == 2), 'wp must be [ntaps, Cout, Cin] or [2, ntaps, Cout, Cin]')
    nt, Cout, Cin2 = wp.shape[-3:]
    wp_hi, wp_lo = (wp[0], wp[1]) if x3 else (wp, None)
    _req(Cin2 == Cin and nt == len(tap_offsets) and wp_hi.is_contiguous() and (wp_lo is None or wp_lo.is_contiguous()), 'wp does not match x / taps')
    if out_view is not None:
        y = out_view
        oh, ow = y.shape[2], y.shape[3]
        _req(y.shape[0] == N and y.shape[1] == Cout and y.stride(1) == 1, 'out_view must be [N,Cout,oh,ow] with unit channel stride')
        y_ptr, y_strides = y.data_ptr(), (y.stride(0), y.stride(2), y.stride(3))
    else:
        oh, ow = out_hw if out_hw is not None else (H, W)
        y_strides = (oh * ow * Cout, ow * Cout, Cout)                 # NHWC, exact strides also for 1-pixel / 1-channel extents
        if query:                                                     # a variant query touches no output: any non-NULL, 16-byte aligned address will do
            y, y_ptr = None, x.data_ptr()
        else:
            y = torch.empty_strided([N, Cout, oh, ow], [y_strides[0], 1, y_strides[1], y_strides[2]], dtype=torch.float32, device=x.device)
            y_ptr = y.data_ptr()
    L = _lib.lib()
    p = _lib.ConvParams()
    p.x, p.wp, p.y = x.data_ptr(), wp_hi.data_ptr(), y_ptr
    if x3:
        p.wp_lo = wp_lo.data_ptr()
    p.n, p.h, p.w, p.cin, p.cout = N, H, W, Cin, Cout
    p.out_h, p.out_w = oh, ow
    p.out_stride_n, p.out_stride_y, p.out_stride_x = y_strides
    p.in_stride, p.ntaps = in_stride, nt
    for i, (dy, dx) in enumerate(tap_offsets):
        p.tap_dy[i], p.tap_dx[i] = int(dy), int(dx)
    keep = []
    for name, t, shape in (('a_scale', a_scale, (N, Cin)), ('o_scale', o_scale, (N, Cout)), ('bias', bias, (Cout,))):
        if t is not None:
            t = t.to(torch.float32).contiguous()
            _req(tuple(t.shape) == shape and t.device == x.device, f'{name} must be a float32 {shape} tensor on the same device')
            keep.append(t)
            setattr(p, name, t.data_ptr())
    p.act = {'linear': 1, 'lrelu': 3}[act]
    p.alpha, p.gain = float(alpha), float(gain)
    p.clamp = float(clamp) if clamp is not None else -1.0
    if not dense:
        p.in_stride_n, p.in_stride_y, p.in_stride_x = x.stride(0), x.stride(2), x.stride(3)
    p.accumulate = int(bool(accumulate))
    p.a_ready = int(bool(a_ready and a_scale is None and not x3))      # x is already TF32-exact: no staging pass (persistent kernel)
    if noise is not None:
        nz = noise.to(torch.float32)
        nz = nz.reshape(1, oh, ow) if nz.ndim == 2 else nz.reshape(nz.shape[0], oh, ow)
        _req(nz.shape[0] in (1, N) and nz.device == x.device, 'noise must be [N or 1, 1, out_h, out_w] on the same device')
        keep.append(nz)
        p.noise = nz.data_ptr()
        p.noise_stride_n, p.noise_stride_y, p.noise_stride_x = (nz.stride(0) if nz.shape[0] == N and N > 1 else 0), nz.stride(1), nz.stride(2)
    if red_out is not None:
        _req(red_x is not None and tuple(red_x.shape) == (N, Cout, oh, ow) and (red_x.stride(0), red_x.stride(2), red_x.stride(3)) == tuple(y_strides)
             and red_x.stride(1) == 1 and red_x.dtype == torch.float32, 'red_x must have the shape and strides of the output')
        _req(red_out.dtype == torch.float32 and red_out.is_contiguous() and tuple(red_out.shape) == (N, Cout), 'red_out must be a contiguous float32 [N, Cout] buffer')
        p.red_x, p.red_out = red_x.data_ptr(), red_out.data_ptr()
    if query:
        v = _lib.ConvVariant()
        with torch.cuda.device(x.device):
            _lib.check(L.sgv_conv2d_tf32_variant(ctypes.byref(p), ctypes.byref(v)), 'sgv_conv2d_tf32_variant')
        return {k: int(getattr(v, k)) for k, _ in v._fields_}
    with torch.cuda.device(x.device):
        _lib.check(L.sgv_conv2d_tf32(ctypes.byref(p), _stream(x.device)), 'sgv_conv2d_tf32')
    return y


TAPS_3x3 = [(ky, kx) for ky in range(3) for kx in range(3)]


def conv3x3_taps():
    """Correlation with padding 1: tap (ky,kx) reads input offset (ky-1, kx-1)."""
    return TAPS_3x3, [(ky - 1, kx - 1) for ky, kx in TAPS_3x3]


def igemm_wgrad(g, x, taps_g, taps_x, out_hw, g_stride=1, x_stride=1, g_scale=None, x_scale=None, out=None, slots=None, g_ready=False, x_ready=False,
                x3=None, query=False):
    """dw[t,o,i] = sum_{n,p} g[n, p*gs + dg_t, o] * g_scale[n,o] * x[n, p*xs + dx_t, i] * x_scale[n,i]  ->  [ntaps, Cout, Cin].

    g: [N, Cout, gh, gw], x: [N, Cin, xh, xw], both channels_last fp32; taps_*: per-tap (dy, dx) pixel offsets.
    g_ready / x_ready: that operand already holds TF32-representable, fully scaled values (its scale must be None): no staging pass.
    out / slots: accumulate tap t into out[slots[t]] of a caller-provided (zeroed) [S, Cout, Cin] buffer instead of a fresh result.
    x3: fp32-grade tf32x3 arithmetic (three passes over hi / residual operand parts; None follows stylegan_v_b200.precision).
    query=True: nothing is launched; returns the kernel variant (sgv_conv2d_wgrad_tf32_variant) as a dict."""
    x3 = _precision.is_x3() if x3 is None else bool(x3)
    for name, t in (('g', g), ('x', x)):
        _req(t.is_cuda and t.dtype == torch.float32 and t.ndim == 4, f'{name} must be a CUDA float32 [N,C,H,W] tensor')
    _req(_is_nhwc(g), 'g must be dense channels_last (NHWC)')
    x_dense = _is_nhwc(x)
    _req(x_dense or (x.stride(1) == 1 and all(st % 4 == 0 for st in (x.stride(0), x.stride(2), x.stride(3)))),
         'x must be channels_last (NHWC), dense or a pixel-strided view with unit channel stride')
    N, Cout, gh, gw = g.shape
    N2, Cin, xh, xw = x.shape
    _req(N == N2 and len(taps_g) == len(taps_x), 'g / x / taps mismatch')
    nt = len(taps_g)
    if out is None:
        dw = torch.zeros([nt, Cout, Cin] if not query else [4], dtype=torch.float32, device=x.device)
    else:
        dw = out
        _req(slots is not None and len(slots) == nt and dw.is_contiguous() and dw.dtype == torch.float32 and tuple(dw.shape[1:]) == (Cout, Cin)
             and all(0 <= int(sl) < dw.shape[0] for sl in slots), 'out must be a contiguous float32 [S, Cout, Cin] buffer and slots index its first dim')
    L = _lib.lib()
    p = _lib.WgradParams()
    p.g, p.x, p.dw = g.data_ptr(), x.data_ptr(), dw.data_ptr()
    p.n, p.gh, p.gw, p.xh, p.xw, p.cin, p.cout = N, gh, gw, xh, xw, Cin, Cout
    p.out_h, p.out_w = out_hw
    p.g_stride, p.x_stride, p.ntaps = g_stride, x_stride, nt
    for i in range(nt):
        p.g_dy[i], p.g_dx[i] = int(taps_g[i][0]), int(taps_g[i][1])
        p.x_dy[i], p.x_dx[i] = int(taps_x[i][0]), int(taps_x[i][1])
    keep = []
    for name, t, shape in (('g_scale', g_scale, (N, Cout)), ('x_scale', x_scale, (N, Cin))):
        if t is not None:
            t = t.to(torch.float32).contiguous()
            _req(tuple(t.shape) == shape, f'{name} must be {shape}')
            keep.append(t)
            setattr(p, name, t.data_ptr())
    if not x_dense:
        p.x_stride_n, p.x_stride_y, p.x_stride_x = x.stride(0), x.stride(2), x.stride(3)
    p.g_ready, p.x_ready = int(bool(g_ready and g_scale is None and not x3)), int(bool(x_ready and x_scale is None and not x3))
    p.precision = int(x3)
    if out is not None:
        p.use_dw_slot = 1
        for i, sl in enumerate(slots):
            p.dw_slot[i] = int(sl)
    if query:
        v = _lib.WgradVariant()
        with torch.cuda.device(x.device):
            _lib.check(L.sgv_conv2d_wgrad_tf32_variant(ctypes.byref(p), ctypes.byref(v)), 'sgv_conv2d_wgrad_tf32_variant')
        return {k: int(getattr(v, k)) for k, _ in v._fields_}
    with torch.cuda.device(x.device):
        _lib.check(L.sgv_conv2d_wgrad_tf32(ctypes.byref(p), _stream(x.device)), 'sgv_conv2d_wgrad_tf32')
    return dw


# ---- one-pass NHWC companions of the fused layer (csrc/layer_elementwise.cu) ----

def _is_nhwc(t):
    N, C, H, W = t.shape
    return t.stride(1) == 1 and t.stride(3) == C and t.stride(2) == W * C and t.stride(0) == H * W * C


def _ptr(t):
    return t.data_ptr() if t is not None else None


def act_bwd(dy, y, bias, act, gain, want_db, want_dd, alpha=0.2, dyimg=None, wmod=None, oscale=None):
    """dz, db[C], dd[N,C] (see sgv_modconv_act_bwd).  dy, y: [N,C,H,W] channels_last.

    With dyimg [N,3,H,W] (contiguous) and wmod [N,3,C] the ToRGB branch reading the same activation is folded in
    (sgv_modconv_act_bwd_rgb): dy may then be None, and a fourth result dwmod [N,3,C] is returned.
    oscale [N,C]: dz is returned as tf32_rn(dz * oscale) (db / dd are reduced from the unscaled dz) — sgv_modconv_act_bwd_ex."""
    _req(y.is_cuda and y.dtype == torch.float32 and _is_nhwc(y), 'y must be an NHWC float32 CUDA tensor')
    _req(dy is None or (dy.dtype == torch.float32 and _is_nhwc(dy) and dy.shape == y.shape), 'dy must match y (NHWC float32)')
    N, C, H, W = y.shape
    rgb = dyimg is not None
    _req(rgb or dy is not None, 'dy is required without a ToRGB branch')
    dz = torch.empty_like(y)
    db = torch.zeros([C], dtype=torch.float32, device=y.device) if want_db else None
    dd = torch.zeros([N, C], dtype=torch.float32, device=y.device) if want_dd else None
    b = bias.to(torch.float32).contiguous() if bias is not None else None
    L = _lib.lib()
    tail = (N, H * W, C, {'linear': 1, 'lrelu': 3}[act], float(alpha), float(gain), _stream(y.device))
    dwmod = None
    if rgb:
        dyimg = dyimg.to(torch.float32).contiguous()
        wmod = wmod.to(torch.float32).contiguous()
        _req(tuple(dyimg.shape) == (N, 3, H, W) and tuple(wmod.shape) == (N, 3, C), 'dyimg must be [N,3,H,W] and wmod [N,3,C]')
        dwmod = torch.zeros([N, 3, C], dtype=torch.float32, device=y.device)
    if oscale is not None:
        oscale = oscale.to(torch.float32).contiguous()
        _req(tuple(oscale.shape) == (N, C), 'oscale must be [N, C]')
    with torch.cuda.device(y.device):
        _lib.check(L.sgv_modconv_act_bwd_ex(_ptr(dy), y.data_ptr(), _ptr(b), dz.data_ptr(), _ptr(db), _ptr(dd), _ptr(dyimg), _ptr(wmod), _ptr(dwmod),
                                            _ptr(oscale), *tail), 'sgv_modconv_act_bwd_ex')
    return (dz, db, dd, dwmod) if rgb else (dz, db, dd)


def scale_reduce(dxs, x, s, want_dx=True, want_ds=True):
    """dx = dxs * s[n,c] (written IN PLACE into dxs), ds[N,C] = sum_hw dxs * x."""
    _req(dxs.is_cuda and dxs.dtype == torch.float32 and _is_nhwc(dxs) and _is_nhwc(x) and dxs.shape == x.shape, 'dxs / x must be matching NHWC float32 CUDA tensors')
    N, C, H, W = dxs.shape
    s = s.to(torch.float32).contiguous()
    ds = torch.zeros([N, C], dtype=torch.float32, device=x.device) if want_ds else None
    L = _lib.lib()
    with torch.cuda.device(x.device):
        _lib.check(L.sgv_modconv_scale_reduce(dxs.data_ptr(), x.data_ptr(), s.data_ptr(), dxs.data_ptr() if want_dx else None, _ptr(ds),
                                              N, H * W, C, _stream(x.device)), 'sgv_modconv_scale_reduce')
    return (dxs if want_dx else None), ds


def torgb_fwd(x, wmod, bias):
    """x [N,C,H,W] NHWC, wmod [N,3,C], bias [3] -> y [N,3,H,W] contiguous (NCHW)."""
    _req(x.is_cuda and x.dtype == torch.float32 and _is_nhwc(x), 'x must be an NHWC float32 CUDA tensor')
    N, C, H, W = x.shape
    wmod = wmod.to(torch.float32).contiguous()
    _req(tuple(wmod.shape) == (N, 3, C), 'wmod must be [N, 3, C]')
    y = torch.empty([N, 3, H, W], dtype=torch.float32, device=x.device)
    b = bias.to(torch.float32).contiguous() if bias is not None else None
    L = _lib.lib()
    with torch.cuda.device(x.device):
        _lib.check(L.sgv_torgb_fwd(x.data_ptr(), wmod.data_ptr(), _ptr(b), y.data_ptr(), N, H * W, C, _stream(x.device)), 'sgv_torgb_fwd')
    return y


def torgb_bwd(dy, x, wmod):
    """dy [N,3,H,W] contiguous, x NHWC, wmod [N,3,C] -> dx (NHWC), dwmod [N,3,C]."""
    N, C, H, W = x.shape
    dy = dy.contiguous()
    wmod = wmod.to(torch.float32).contiguous()
    dx = torch.empty_like(x)
    dwmod = torch.zeros([N, 3, C], dtype=torch.float32, device=x.device)
    L = _lib.lib()
    with torch.cuda.device(x.device):
        _lib.check(L.sgv_torgb_bwd(dy.data_ptr(), x.data_ptr(), wmod.data_ptr(), dx.data_ptr(), dwmod.data_ptr(), N, H * W, C, _stream(x.device)), 'sgv_torgb_bwd')
    return dx, dwmod
