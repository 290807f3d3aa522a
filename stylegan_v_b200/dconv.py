"""Fused convolution + bias + activation node for the (non-modulated) convolution layers of the discriminator — the
`Conv2dLayer` arithmetic of the reference (src/training/layers.py:184-197: conv2d_resample then bias_act) with the bias /
leaky-ReLU / gain epilogue inside the implicit-GEMM launch and a one-pass activation-gradient kernel in the backward:

    forward    ONE tcgen05 launch  y = act(conv(x, w, stride s, padding p) + b) * gain         (csrc/conv_tf32*.cu epilogue)
    backward   ONE pass  dz = act'(y) * dy * gain, db = sum dz                                  (sgv_modconv_act_bwd: y -> pre-activation)
               data gradient   = the transposed contraction (polyphase launches for s = 2)      (native_conv.conv_forward, transpose)
               weight gradient = split-K weight-gradient kernel                                 (native_conv.conv_weight_grad)

instead of conv + bias_act (2 kernels, one extra read + write of the activation) forward and bias_act-grad + sum + conv
gradients backward.  The low-pass FIR of the down-sampling layers stays the drop-in `upfirdn2d` op in front of this node
(conv2d_resample.py:100-110 geometry), so a down layer is FIR launch + ONE conv launch.

First-order only (`once_differentiable`): the R1 phase, which differentiates the discriminator twice, uses the unfused drop-in ops.
CUDA / fp32 / channel counts inside the native envelope only — `supported()` says whether a call qualifies; there is no fallback
inside the node.
"""
import torch

from . import _lib
from . import conv as _conv
from . import native_conv as _native
from . import precision as _precision


def supported(x, w, stride, padding):
    """True when forward, data gradient and weight gradient of this call all run on the tcgen05 kernels."""
    if not (_native.enabled and x.is_cuda and x.dtype == torch.float32 and w.dtype == torch.float32 and x.ndim == 4):
        return False
    o, i, kh, kw = w.shape
    if kh != kw or kh not in (1, 3) or stride not in (1, 2) or x.shape[1] != i:
        return False
    if (x.shape[2] + 2 * padding - kh) // stride + 1 < 1 or (x.shape[3] + 2 * padding - kw) // stride + 1 < 1:
        return False
    return _native._ok_channels(i, o) and _native._ok_channels(o, i) and i % 32 == 0 and o % 32 == 0


class _FusedConvAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, stride, padding, act, gain, weight_gain):
        x = x.contiguous(memory_format=torch.channels_last)
        k = w.shape[2]
        taps = [(ky, kx) for ky in range(k) for kx in range(k)]
        oh, ow = (x.shape[2] + 2 * padding - k) // stride + 1, (x.shape[3] + 2 * padding - k) // stride + 1
        offs = [(ky - padding, kx - padding) for ky, kx in taps]
        x3 = _precision.is_x3()
        # the equalised-lr weight gain (layers.py:186) rides on the weight-preparation pass: no scaled weight copy, no multiply in the backward
        y = _conv.igemm_conv(x, _conv.prep_weights(w, taps, scale=weight_gain, x3=x3), offs, out_hw=(oh, ow), in_stride=stride, bias=b, act=act, gain=gain)
        # a linear layer's gradient does not need y (the block adds the two branches IN PLACE into the skip branch's output, networks.py:481)
        ctx.save_for_backward(x, w, y if act != 'linear' else x.new_empty(0), b if b is not None else x.new_empty(0))
        ctx.cfg = (stride, padding, act, gain, b is not None, weight_gain)
        ctx.x3 = x3
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        x, w, y, b = ctx.saved_tensors
        stride, padding, act, gain, has_b, weight_gain = ctx.cfg
        dy = dy.contiguous(memory_format=torch.channels_last)
        want_db = has_b and ctx.needs_input_grad[2]
        if act == 'linear':
            dz = dy * gain if gain != 1 else dy
            db = dz.sum(dim=[0, 2, 3]) if want_db else None
        else:
            dz, db, _ = _conv.act_bwd(dy, y, b if has_b else None, act, gain, want_db, False)
        gx = gw = None
        k = w.shape[2]
        taps = [(ky, kx) for ky in range(k) for kx in range(k)]
        if ctx.needs_input_grad[0]:
            s2, p2 = (stride, stride), (padding, padding)
            op = tuple(x.shape[i + 2] - (dz.shape[i + 2] - 1) * stride - (1 - 2 * padding) - (k - 1) for i in range(2))   # conv2d_gradfix.py:95-104
            gx = _native.conv_forward(dz, w, None, True, s2, p2, op, (1, 1), 1, weight_scale=weight_gain, x3=ctx.x3)
            assert gx is not None and gx.shape == x.shape
        if ctx.needs_input_grad[1]:
            gw = _native.conv_weight_grad(dz, x, tuple(w.shape), False, (stride, stride), (padding, padding), (0, 0), (1, 1), 1, x3=ctx.x3)
            assert gw is not None
            if weight_gain != 1:
                gw = gw * weight_gain
        return gx, gw, db, None, None, None, None, None


def fused_conv_act(x, w, b=None, stride=1, padding=0, act='linear', gain=1.0, weight_gain=1.0):
    """act(conv2d(x, w * weight_gain, stride, padding) + b) * gain as one autograd node; act in {'linear', 'lrelu'}.  Caller checks supported()."""
    assert act in ('linear', 'lrelu')
    return _FusedConvAct.apply(x, w, b, int(stride), int(padding), act, float(gain), float(weight_gain))


class _FromRgb(torch.autograd.Function):
    """y (NHWC) = act(conv1x1(img (NCHW, <= 4 channels), w * weight_gain) + b) * gain: one streaming pass each way (csrc/disc_ops.cu)."""

    @staticmethod
    def forward(ctx, img, w, b, act, gain, weight_gain):
        img = img.contiguous()
        N, J, H, W = img.shape
        O = w.shape[0]
        y = torch.empty_strided([N, O, H, W], [H * W * O, 1, W * O, O], dtype=torch.float32, device=img.device)
        w2 = w.reshape(O, J).contiguous()
        with torch.cuda.device(img.device):
            _lib.check(_lib.lib().sgv_fromrgb_fwd(img.data_ptr(), w2.data_ptr(), b.data_ptr() if b is not None else None, y.data_ptr(), N, H * W, O, J,
                                                  float(weight_gain), {'linear': 1, 'lrelu': 3}[act], 0.2, float(gain), _conv._stream(img.device)), 'sgv_fromrgb_fwd')
        ctx.save_for_backward(img, w2, y if act != 'linear' else img.new_empty(0))
        ctx.cfg = (act, gain, weight_gain, b is not None, tuple(w.shape))
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        img, w2, y = ctx.saved_tensors
        act, gain, weight_gain, has_b, w_shape = ctx.cfg
        N, J, H, W = img.shape
        O = w2.shape[0]
        dy = dy.contiguous(memory_format=torch.channels_last)
        want_db = has_b and ctx.needs_input_grad[2]
        if act == 'linear':
            dz = dy * gain if gain != 1 else dy
            db = dz.sum(dim=[0, 2, 3]) if want_db else None
        else:
            dz, db, _ = _conv.act_bwd(dy, y, None, act, gain, want_db, False)
        dimg = torch.empty_like(img) if ctx.needs_input_grad[0] else None
        dw = torch.zeros([O, J], dtype=torch.float32, device=img.device)
        with torch.cuda.device(img.device):
            _lib.check(_lib.lib().sgv_fromrgb_bwd(dz.data_ptr(), img.data_ptr(), w2.data_ptr(), dimg.data_ptr() if dimg is not None else None, dw.data_ptr(),
                                                  N, H * W, O, J, float(weight_gain), _conv._stream(img.device)), 'sgv_fromrgb_bwd')
        return dimg, (dw.reshape(w_shape) if ctx.needs_input_grad[1] else None), db, None, None, None


def fromrgb_supported(img, w):
    o = w.shape[0]
    return (_native.enabled and img.is_cuda and img.dtype == torch.float32 and w.dtype == torch.float32 and img.ndim == 4 and img.shape[1] <= 4
            and tuple(w.shape[2:]) == (1, 1) and o % 4 == 0 and o // 4 <= 32 and (o // 4) & (o // 4 - 1) == 0)


def fromrgb(img, w, b=None, act='lrelu', gain=1.0, weight_gain=1.0):
    """The discriminator's first 1x1 layer on the frames (networks.py:467-470): NCHW image in, NHWC activation out.  Caller checks fromrgb_supported()."""
    return _FromRgb.apply(img, w, b, act, float(gain), float(weight_gain))


class _MinibatchStd(torch.autograd.Function):
    """MinibatchStdLayer + concat (+ zero channels up to `cpad`) -> NHWC [N, cpad, H, W]; csrc/disc_ops.cu.  First order."""

    @staticmethod
    def forward(ctx, x, group, num_channels, cpad):
        N, Cc, H, W = x.shape
        y = torch.empty_strided([N, cpad, H, W], [H * W * cpad, 1, W * cpad, cpad], dtype=torch.float32, device=x.device)
        sd = torch.empty([N // group, num_channels], dtype=torch.float32, device=x.device)
        assert x.stride(2) == W * x.stride(3), 'pixels of a sample must be addressable with one stride'
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().sgv_mbstd_fwd(x.data_ptr(), x.stride(0), x.stride(1), x.stride(3), y.data_ptr(), sd.data_ptr(), N, Cc, H * W, cpad,
                                                group, num_channels, _conv._stream(x.device)), 'sgv_mbstd_fwd')
        ctx.save_for_backward(x)
        ctx.cfg = (group, num_channels, cpad)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        x, = ctx.saved_tensors
        group, num_channels, cpad = ctx.cfg
        N, Cc, H, W = x.shape
        dy = dy.contiguous(memory_format=torch.channels_last)
        dx = torch.empty_strided([N, Cc, H, W], [H * W * Cc, 1, W * Cc, Cc], dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().sgv_mbstd_bwd(dy.data_ptr(), x.data_ptr(), x.stride(0), x.stride(1), x.stride(3), dx.data_ptr(), N, Cc, H * W, cpad,
                                                group, num_channels, _conv._stream(x.device)), 'sgv_mbstd_bwd')
        return dx, None, None, None


def minibatch_std_concat(x, group_size, num_channels=1, pad_to=64):
    """[N, C, H, W] -> NHWC [N, roundup(C + num_channels, pad_to), H, W]: x, the per-group standard-deviation statistic as extra channel(s),
    zero channels after that (the consumer pads its weight with zero input channels).  CUDA float32; N % G == 0 like the reference's reshape."""
    N, Cc = x.shape[0], x.shape[1]
    G = min(group_size, N) if group_size is not None else N
    cpad = (Cc + num_channels + pad_to - 1) // pad_to * pad_to
    return _MinibatchStd.apply(x, int(G), int(num_channels), int(cpad))
