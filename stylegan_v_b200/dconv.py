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

from . import conv as _conv
from . import native_conv as _native


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
    def forward(ctx, x, w, b, stride, padding, act, gain):
        x = x.contiguous(memory_format=torch.channels_last)
        k = w.shape[2]
        taps = [(ky, kx) for ky in range(k) for kx in range(k)]
        oh, ow = (x.shape[2] + 2 * padding - k) // stride + 1, (x.shape[3] + 2 * padding - k) // stride + 1
        y = _conv.igemm_conv(x, _conv.prep_weights(w, taps), [(ky - padding, kx - padding) for ky, kx in taps], out_hw=(oh, ow),
                             in_stride=stride, bias=b, act=act, gain=gain)
        # a linear layer's gradient does not need y (the block adds the two branches IN PLACE into the skip branch's output, networks.py:481)
        ctx.save_for_backward(x, w, y if act != 'linear' else x.new_empty(0), b if b is not None else x.new_empty(0))
        ctx.cfg = (stride, padding, act, gain, b is not None)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        x, w, y, b = ctx.saved_tensors
        stride, padding, act, gain, has_b = ctx.cfg
        dy = dy.contiguous(memory_format=torch.channels_last)
        want_db = has_b and ctx.needs_input_grad[2]
        if act == 'linear':
            dz = dy * gain if gain != 1 else dy
            db = dz.sum(dim=[0, 2, 3]) if want_db else None
        else:
            dz, db, _ = _conv.act_bwd(dy, y, b if has_b else None, act, gain, want_db, False)
        gx = gw = None
        s2, p2 = (stride, stride), (padding, padding)
        if ctx.needs_input_grad[0]:
            k = w.shape[2]
            op = tuple(x.shape[i + 2] - (dz.shape[i + 2] - 1) * stride - (1 - 2 * padding) - (k - 1) for i in range(2))   # conv2d_gradfix.py:95-104
            gx = _native.conv_forward(dz, w, None, True, s2, p2, op, (1, 1), 1)
            assert gx is not None and gx.shape == x.shape
        if ctx.needs_input_grad[1]:
            gw = _native.conv_weight_grad(dz, x, tuple(w.shape), False, s2, p2, (0, 0), (1, 1), 1)
            assert gw is not None
        return gx, gw, db, None, None, None, None


def fused_conv_act(x, w, b=None, stride=1, padding=0, act='linear', gain=1.0):
    """act(conv2d(x, w, stride, padding) + b) * gain as one autograd node; act in {'linear', 'lrelu'}.  Caller checks supported()."""
    assert act in ('linear', 'lrelu')
    return _FusedConvAct.apply(x, w, b, int(stride), int(padding), act, float(gain))
