"""FullyConnectedLayer (src/training/layers.py:108-138) on the tcgen05 contraction kernels: the mapping networks of G and D and the
discriminator epilogue's dense layers.

The reference runs these as torch.addmm / matmul (cuBLAS) + bias_act.  Here a layer is ONE launch of the implicit-GEMM kernel of
libsgv_b200 (include/sgv_b200_conv.h) on x [M, K] viewed as the NHWC image [M, 1, 1, K] with one tap: the equalised-lr weight gain is
folded into the weight preparation pass, bias / leaky ReLU / gain run in the epilogue, and both gradients run on the same kernels.
The generic line form (x [B, C, 1, L], k taps along L) is kept because it is what the kernel computes; k = 1, L = 1 is the dense layer.

Always the fp32-grade `tf32x3` arithmetic, whatever stylegan_v_b200.precision says (the reference computes these layers in fp32 and
their FLOPs are negligible).  Measured accuracy (tests/test_dense_gpu.py): 7e-6 of fp64 at K = 512, 5e-5 at K = 8192 — the tensor
core's fp32 accumulation is not round-to-nearest, so the error grows with K.  That is why the TIME ENCODER stays on true-fp32 library
GEMMs / conv1d (stylegan_v_b200/time_encoder.py: its outputs are multiplied by phase scales up to 64 before sin / cos), and the stacked
style affines of the synthesis network stay on cuBLAS for speed (stylegan_v_b200/synthesis.py::_all_styles).
CUDA / float32 only; anything else (CPU tensors, fp16) takes the PyTorch formulation in the calling module, like the reference's ops do.
"""
import torch

from . import conv as _conv


def _as_image(x2d):
    """[M, K] contiguous -> the same memory as an NHWC [M, K, 1, 1] tensor (channel stride 1, every pixel stride = K)."""
    M, K = x2d.shape
    return x2d.as_strided([M, K, 1, 1], [K, 1, K, K])


def _nhwc_exact(t):
    """[N, C, H, W] in any strides -> the same values with strides exactly (H*W*C, 1, W*C, C) (no copy when the memory already is NHWC;
    size-1 dims make torch's own channels_last test ambiguous, the kernels' tensor maps are not)."""
    N, Cc, H, W = t.shape
    m = t.permute(0, 2, 3, 1).contiguous()
    return m.as_strided([N, Cc, H, W], [H * W * Cc, 1, W * Cc, Cc])


def supported(x, weight, act='linear'):
    """Channel counts the contraction kernels accept as GEMM-K / GEMM-N for forward AND both gradients (and, for lrelu, the one-pass
    activation-gradient kernel's layout rule)."""
    o, i = weight.shape[0], weight.shape[1]
    ok = lambda k, n: k % 32 == 0 and (n % 64 == 0 or n == 32)
    act_ok = act == 'linear' or (o % 4 == 0 and 256 % (o // 4) == 0)
    return x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and ok(i, o) and ok(o, i) and act_ok


class _DenseConv(torch.autograd.Function):
    """y = act(conv_k(x, w * weight_gain) + b) on NHWC lines ([B, C, 1, L]); k = 1, L = 1 is the fully-connected layer.  First order."""

    @staticmethod
    def forward(ctx, x4, weight, bias, weight_gain, act, gain):
        O, I, k = weight.shape
        B, _, _, L = x4.shape
        Lout = L - k + 1
        taps = [(0, j) for j in range(k)]
        w4 = weight.unsqueeze(2)                                                     # [O, I, 1, k]
        wp = _conv.prep_weights(w4, taps, scale=weight_gain, x3=True)
        y = _conv.igemm_conv(x4, wp, taps, out_hw=(1, Lout), bias=bias, act=act, gain=gain)
        ctx.save_for_backward(x4, weight, y if act != 'linear' else x4.new_empty(0))
        ctx.cfg = (weight_gain, act, k, gain)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        x4, weight, y = ctx.saved_tensors
        weight_gain, act, k, gain = ctx.cfg
        O, I, _ = weight.shape
        B, _, _, L = x4.shape
        Lout = L - k + 1
        dy = _nhwc_exact(dy)
        taps = [(0, j) for j in range(k)]
        db = None
        if act == 'linear':
            dz = dy if gain == 1 else dy * gain
            if ctx.needs_input_grad[2]:
                db = dz.sum(dim=[0, 2, 3])
        else:
            dz, db, _ = _conv.act_bwd(dy, y, None, act, gain, ctx.needs_input_grad[2], False)    # slope from the saved output (bias already inside y)
        dx = dw = None
        w4 = weight.unsqueeze(2)
        if ctx.needs_input_grad[0]:
            wpt = _conv.prep_weights(w4, taps, rows_dim=1, cols_dim=0, scale=weight_gain, x3=True)
            dx = _conv.igemm_conv(dz, wpt, [(0, -j) for j in range(k)], out_hw=(1, L))            # full correlation: out-of-range gradient reads as zero
        if ctx.needs_input_grad[1]:
            dwt = _conv.igemm_wgrad(dz, x4, [(0, 0)] * k, taps, (1, Lout), x3=True)               # [k, O, I]
            dw = dwt.permute(1, 2, 0) * weight_gain
        return dx, dw, db, None, None, None


def linear(x, weight, bias=None, weight_gain=1.0, bias_gain=1.0, act='linear', gain=1.0):
    """act(x @ (weight * weight_gain).T + bias * bias_gain) * gain for x [M, K], weight [O, K]; act in {'linear', 'lrelu'} (bias_act's
    default sqrt(2) for lrelu is the caller's `gain`).  Caller checks supported()."""
    assert act in ('linear', 'lrelu') and x.ndim == 2
    b = None
    if bias is not None:
        b = bias if bias_gain == 1 else bias * bias_gain
    y = _DenseConv.apply(_as_image(x.contiguous()), weight.unsqueeze(2), b, float(weight_gain), act, float(gain))
    return y.reshape(x.shape[0], weight.shape[0])
