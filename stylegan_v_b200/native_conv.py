"""Lowers torch-style conv2d / conv_transpose2d calls (what `conv2d_gradfix` receives from the reference's
`conv2d_resample`, networks and layers) onto the tcgen05 implicit-GEMM kernels of libsgv_b200.

Covered (everything the synthesis and discriminator blocks issue, conv2d_resample.py:106-147):
  conv2d            k in {1,3}, stride 1, any symmetric padding        -> one launch (halo-patch kernel)
  conv2d            k in {1,3}, stride 2 (discriminator down path)     -> one launch, TMA element strides
  conv_transpose2d  k = 3, stride 2, padding 0 (generator up path)     -> 4 polyphase launches into the (2h+1) lattice
  conv_transpose2d  k in {1,3}, stride 1 (data gradient of conv2d)     -> one launch with transposed weights
and the matching weight gradients (stylegan_v_b200/csrc/wgrad_tf32.cu).  Requirements: CUDA fp32, groups = 1,
dilation 1, GEMM-K channels % 32 == 0, GEMM-N channels 32 or a multiple of 64.
Anything else returns None and the caller uses the library (cuDNN) call exactly like the reference.

Tensors are converted to channels_last (NHWC) on entry if necessary and results are returned channels_last — the
reference's ops accept either memory format (bias_act.py:148, upfirdn2d.cpp:35), so activations then stay NHWC.
Arithmetic: TF32 products (operands rounded to nearest), fp32 accumulation.
"""
import torch

from . import conv as _conv

enabled = True


def _ok_channels(k_ch, n_ch):
    """forward / data-gradient kernel: GEMM-K channels % 32, GEMM-N channels % 64 (or exactly 32).  16 output channels are NOT routed here: the
    kernel's epilogue works in 32-column TMEM slices, so a 16-wide tile stored nothing (wrong 32 -> 16 channel down layers of a tiny
    discriminator in the round-1 GPU run, scripts/debug_d_layers.py); no configuration of the reference has such a layer (narrowest: 64
    channels at 256^2, 32 at 1024^2) — those calls take the library path, and the C ABI rejects them."""
    return k_ch % 32 == 0 and (n_ch % 64 == 0 or n_ch == 32)


def _common_ok(x, w, dilation, groups):
    return (enabled and x.is_cuda and x.dtype == torch.float32 and w.dtype == torch.float32 and groups == 1
            and tuple(dilation) == (1, 1) and w.shape[2] == w.shape[3] and w.shape[2] in (1, 3) and x.ndim == 4)


def _nhwc(t):
    return t.contiguous(memory_format=torch.channels_last)


def _taps(k):
    return [(ky, kx) for ky in range(k) for kx in range(k)]


def conv_forward(x, w, b, transpose, stride, padding, output_padding, dilation, groups, weight_scale=1.0, x3=None):
    """Returns the convolution result (channels_last) or None when the call is outside the native envelope.
    weight_scale is folded into the weight-preparation pass (equalised-lr gain without a scaled weight copy); x3 selects the arithmetic
    mode (None = stylegan_v_b200.precision at call time; autograd nodes pass the mode of their forward pass)."""
    if not _common_ok(x, w, dilation, groups):
        return None
    k = w.shape[2]
    N, _, H, W = x.shape
    taps = _taps(k)
    if not transpose:
        O, I = w.shape[0], w.shape[1]
        if x.shape[1] != I or not _ok_channels(I, O) or stride[0] != stride[1] or padding[0] != padding[1]:
            return None
        s, p = stride[0], padding[0]
        if s not in (1, 2):
            return None
        oh, ow = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
        if oh < 1 or ow < 1:
            return None
        wp = _conv.prep_weights(w, taps, scale=weight_scale, x3=x3)
        return _conv.igemm_conv(_nhwc(x), wp, [(ky - p, kx - p) for ky, kx in taps], out_hw=(oh, ow), in_stride=s, bias=b)
    # transposed: weight [Cin, Cout, k, k]
    I, O = w.shape[0], w.shape[1]
    if x.shape[1] != I or not _ok_channels(I, O) or stride[0] != stride[1] or padding[0] != padding[1]:
        return None
    s, p = stride[0], padding[0]
    if s == 1:
        if tuple(output_padding) != (0, 0):
            return None
        oh, ow = H + k - 1 - 2 * p, W + k - 1 - 2 * p
        if oh < 1 or ow < 1:
            return None
        wp = _conv.prep_weights(w, taps, rows_dim=1, cols_dim=0, scale=weight_scale, x3=x3)
        return _conv.igemm_conv(_nhwc(x), wp, [(p - ky, p - kx) for ky, kx in taps], out_hw=(oh, ow), bias=b)
    if s == 2 and k == 3 and p == 0:
        op = tuple(output_padding)
        oh, ow = 2 * H + 1 + op[0], 2 * W + 1 + op[1]
        # allocated channels_last directly (an NCHW allocation + .contiguous(channels_last) is a full-size copy of an uninitialised tensor:
        # 0.8 GB per call for the 256^2 discriminator layers — found in the round-2 G+D timeline)
        u = torch.empty([N, O, oh, ow], dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
        if op != (0, 0):
            u.zero_()           # the polyphase launches below do not cover the output-padding row / column
        xn = _nhwc(x)
        for a in (0, 1):
            for c in (0, 1):
                ph_taps = [(ky, kx) for ky in range(a, 3, 2) for kx in range(c, 3, 2)]
                offs = [(-((ky - a) // 2), -((kx - c) // 2)) for ky, kx in ph_taps]
                view = u[:, :, a:2 * H + 1:2, c:2 * W + 1:2]
                _conv.igemm_conv(xn, _conv.prep_weights(w, ph_taps, rows_dim=1, cols_dim=0, scale=weight_scale, x3=x3), offs, out_view=view, bias=b)
        return u
    return None


def stride2_phases(k, p):
    """Pixel-parity decomposition of a stride-2 correlation with a k x k kernel and padding p: x[2 o + t - p] = x_phase(a)[o + (t - p - a) / 2]
    with a = (t - p) mod 2.  Returns [(a, c, offsets, slots)]: for the view x[:, :, a::2, c::2] the taps (ky, kx) that fall on it, as stride-1
    pixel offsets into that view and as their slots ky * k + kx in the [k * k, O, I] result."""
    out = []
    for a in (0, 1):
        for c in (0, 1):
            ph = [(ky, kx) for ky in range(k) for kx in range(k) if (ky - p) % 2 == a and (kx - p) % 2 == c]
            if ph:
                out.append((a, c, [((ky - p - a) // 2, (kx - p - c) // 2) for ky, kx in ph], [ky * k + kx for ky, kx in ph]))
    return out


def conv_weight_grad(gy, x, w_shape, transpose, stride, padding, output_padding, dilation, groups, x3=None):
    """Weight gradient (shape w_shape) of the op above, or None when outside the native envelope."""
    k = w_shape[2]
    if not (enabled and x.is_cuda and x.dtype == torch.float32 and gy.dtype == torch.float32 and groups == 1 and tuple(dilation) == (1, 1)
            and w_shape[2] == w_shape[3] and k in (1, 3) and stride[0] == stride[1] and padding[0] == padding[1]):
        return None
    s, p = stride[0], padding[0]
    taps = _taps(k)
    N, _, H, W = x.shape
    if not transpose:
        O, I = w_shape[0], w_shape[1]
        if I % 32 or O % 32 or s not in (1, 2):
            return None
        oh, ow = gy.shape[2], gy.shape[3]
        if s == 2 and k == 3 and ow >= 8 and oh >= 4:
            # stride 2 (the discriminator's down layers): x[2 oy + ky - p] = x_phase(a)[oy + (ky - p - a) / 2] with a = (ky - p) mod 2, so each of
            # the four pixel-parity views of x is a STRIDE-1 weight gradient over its 4 / 2 / 2 / 1 taps — and runs on the grouped-tap kernel
            # (one patch per dy row) instead of nine per-tap passes of the strided kernel (0.87 ms -> see profiles/timeline_gd_step_r2*.txt for
            # 64 -> 128 channels at 128^2 x 48 frames).  The four calls accumulate into their tap slots of one buffer.
            g, xn = _nhwc(gy), _nhwc(x)
            dwt = torch.zeros([9, O, I], dtype=torch.float32, device=x.device)
            for a, c, offs, slots in stride2_phases(k, p):
                _conv.igemm_wgrad(g, xn[:, :, a::2, c::2], [(0, 0)] * len(offs), offs, (oh, ow), out=dwt, slots=slots, x3=x3)
            return dwt.reshape(3, 3, O, I).permute(2, 3, 0, 1)
        dw = _conv.igemm_wgrad(_nhwc(gy), _nhwc(x), [(0, 0)] * len(taps), [(ky - p, kx - p) for ky, kx in taps], (oh, ow), x_stride=s, x3=x3)
        return dw.reshape(k, k, O, I).permute(2, 3, 0, 1)
    I, O = w_shape[0], w_shape[1]
    if I % 32 or O % 32:
        return None
    if s == 1:
        # y[Y] = sum_ky x[Y + p - ky] w[ky]  =>  dw[i, o, ky] = sum_Y gy[Y, o] x[Y + p - ky, i]; lattice = gy's extent
        oh, ow = gy.shape[2], gy.shape[3]
        dw = _conv.igemm_wgrad(_nhwc(gy), _nhwc(x), [(0, 0)] * len(taps), [(p - ky, p - kx) for ky, kx in taps], (oh, ow), x3=x3)
        return dw.reshape(k, k, O, I).permute(3, 2, 0, 1)
    if s == 2 and k == 3 and p == 0:
        g = _nhwc(gy)
        xn = _nhwc(x)
        if H >= 8 and W >= 8:
            dwt = torch.zeros([9, I, O], dtype=torch.float32, device=x.device)      # the four polyphase calls accumulate into their tap slots
            for a in (0, 1):
                for c in (0, 1):
                    ph_taps = [(ky, kx) for ky in range(a, 3, 2) for kx in range(c, 3, 2)]
                    offs = [(ky // 2, kx // 2) for ky, kx in ph_taps]
                    _conv.igemm_wgrad(xn, g[:, :, a:2 * H + 1:2, c:2 * W + 1:2], [(0, 0)] * len(ph_taps), offs, (H, W),
                                      out=dwt, slots=[ky * 3 + kx for ky, kx in ph_taps], x3=x3)
            return dwt.reshape(3, 3, I, O).permute(2, 3, 0, 1)
        dw = _conv.igemm_wgrad(g, xn, taps, [(0, 0)] * 9, (H, W), g_stride=2, x3=x3)
        return dw.reshape(3, 3, O, I).permute(3, 2, 0, 1)
    return None
