"""Drop-in for `src.torch_utils.ops.conv2d_gradfix` (reference: src/torch_utils/ops/conv2d_gradfix.py).

Public surface kept: conv2d(), conv_transpose2d() with torch.nn.functional signatures, module globals
`enabled` / `weight_gradients_disabled`, and the `no_weight_gradients()` context manager
(conv2d_gradfix.py:22-43).

The reference's custom op only activates on torch 1.7-1.10 (conv2d_gradfix.py:53-56) because it reaches
into `aten::cudnn_convolution_backward_weight`; on newer torch it silently degrades to F.conv2d and
`no_weight_gradients()` stops having any effect.  Here the op is rebuilt on the stable
`aten::convolution_backward` entry point so it works on current torch:

  forward        dense contraction: the tcgen05 implicit-GEMM kernels of libsgv_b200 for the shapes they cover
                 (stylegan_v_b200/native_conv.py: every conv the synthesis / discriminator blocks issue with channel
                 counts % 32), the cuDNN library call otherwise — like the reference;
  grad input     the transposed op of the same class (arbitrary-order differentiable, conv2d_gradfix.py:125-128);
  grad weight    separate autograd node that is skipped entirely when `weight_gradients_disabled`
                 (conv2d_gradfix.py:130-132) and is itself differentiable (conv2d_gradfix.py:151-165).
"""
import contextlib
import torch

from .. import native_conv as _native

enabled = False                     # the reference's training loop sets this to True (training_loop.py:143)
weight_gradients_disabled = False


@contextlib.contextmanager
def no_weight_gradients():
    global weight_gradients_disabled
    saved = weight_gradients_disabled
    weight_gradients_disabled = True
    try:
        yield
    finally:
        weight_gradients_disabled = saved


def _pair(v):
    v = tuple(v) if isinstance(v, (tuple, list)) else (v, v)
    assert len(v) == 2 and all(isinstance(i, int) for i in v)
    return v


def _use_custom(x):
    assert isinstance(x, torch.Tensor)
    return enabled and x.device.type == 'cuda'


def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    if _use_custom(input):
        return _Conv.apply(input, weight, bias, False, _pair(stride), _pair(padding), (0, 0), _pair(dilation), groups)
    return torch.nn.functional.conv2d(input=input, weight=weight, bias=bias, stride=stride, padding=padding, dilation=dilation, groups=groups)


def conv_transpose2d(input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    if _use_custom(input):
        return _Conv.apply(input, weight, bias, True, _pair(stride), _pair(padding), _pair(output_padding), _pair(dilation), groups)
    return torch.nn.functional.conv_transpose2d(input=input, weight=weight, bias=bias, stride=stride, padding=padding,
                                                output_padding=output_padding, groups=groups, dilation=dilation)


def _forward_op(x, w, b, transpose, stride, padding, output_padding, dilation, groups):
    y = _native.conv_forward(x, w, b, transpose, stride, padding, output_padding, dilation, groups)   # tcgen05 kernels when the shape allows
    if y is not None:
        # The kernels compute in NHWC.  Like F.conv2d, the op hands back the memory format of its INPUT: an NCHW-contiguous caller (the unmodified
        # reference networks: their .view() of activations assumes it, networks.py:659-662) gets an NCHW-contiguous result, a channels_last caller
        # (the native modules) keeps channels_last and pays no conversion.
        if x.is_contiguous() and not (x.shape[1] == 1 or (x.shape[2] == 1 and x.shape[3] == 1)):
            y = y.contiguous()
        return y
    F = torch.nn.functional
    if transpose:
        return F.conv_transpose2d(x, w, b, stride=stride, padding=padding, output_padding=output_padding, groups=groups, dilation=dilation)
    return F.conv2d(x, w, b, stride=stride, padding=padding, dilation=dilation, groups=groups)


class _Conv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, transpose, stride, padding, output_padding, dilation, groups):
        ctx.cfg = (transpose, stride, padding, output_padding, dilation, groups)
        ctx.save_for_backward(x, w)
        return _forward_op(x, w, b, transpose, stride, padding, output_padding, dilation, groups)

    @staticmethod
    def backward(ctx, gy):
        transpose, stride, padding, output_padding, dilation, groups = ctx.cfg
        x, w = ctx.saved_tensors
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            if transpose:
                op = (0, 0)
            else:   # output_padding that makes the transposed op reproduce x's extent (conv2d_gradfix.py:95-104)
                op = tuple(x.shape[i + 2] - (gy.shape[i + 2] - 1) * stride[i] - (1 - 2 * padding[i]) - dilation[i] * (w.shape[i + 2] - 1)
                           for i in range(2))
            gx = _Conv.apply(gy, w, None, not transpose, stride, padding, op, dilation, groups)
            assert gx.shape == x.shape
        if ctx.needs_input_grad[1] and not weight_gradients_disabled:
            gw = _ConvGradWeight.apply(gy, x, tuple(w.shape), transpose, stride, padding, output_padding, dilation, groups)
        if ctx.needs_input_grad[2]:
            gb = gy.sum([0, 2, 3])
        return gx, gw, gb, None, None, None, None, None, None


class _ConvGradWeight(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gy, x, w_shape, transpose, stride, padding, output_padding, dilation, groups):
        ctx.cfg = (w_shape, transpose, stride, padding, output_padding, dilation, groups)
        ctx.save_for_backward(gy, x)
        gw = _native.conv_weight_grad(gy, x, w_shape, transpose, stride, padding, output_padding, dilation, groups)
        if gw is not None:
            return gw
        w_like = torch.empty(w_shape, dtype=x.dtype, device=x.device)
        _, gw, _ = torch.ops.aten.convolution_backward(gy, x, w_like, None, list(stride), list(padding), list(dilation),
                                                      transpose, list(output_padding), groups, [False, True, False])
        assert gw.shape == tuple(w_shape)
        return gw

    @staticmethod
    def backward(ctx, ggw):
        w_shape, transpose, stride, padding, output_padding, dilation, groups = ctx.cfg
        gy, x = ctx.saved_tensors
        g_gy = g_x = None
        if ctx.needs_input_grad[0]:
            g_gy = _Conv.apply(x, ggw, None, transpose, stride, padding, output_padding, dilation, groups)
            assert g_gy.shape == gy.shape
        if ctx.needs_input_grad[1]:
            if transpose:
                op = (0, 0)
            else:
                op = tuple(x.shape[i + 2] - (gy.shape[i + 2] - 1) * stride[i] - (1 - 2 * padding[i]) - dilation[i] * (w_shape[i + 2] - 1)
                           for i in range(2))
            g_x = _Conv.apply(gy, ggw, None, not transpose, stride, padding, op, dilation, groups)
            assert g_x.shape == x.shape
        return g_gy, g_x, None, None, None, None, None, None, None
