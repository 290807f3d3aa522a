"""modulated_conv2d on the drop-in ops — the arbitrarily-differentiable formulation of the layer the fused kernels implement.

Same semantics and argument meaning as the reference's `modulated_conv2d` (src/training/networks.py:30-86): train-mode
path (scale activations by the styles, shared-weight convolution, scale by the demodulation coefficients) and the
eval-mode per-sample-weight grouped convolution.  It exists next to the fused layer (stylegan_v_b200/modconv.py) for the two
things that one cannot do: gradients of gradients (path-length regularisation differentiates d img / d ws a second time,
loss.py:108-119) and CPU tensors.  Every primitive is a drop-in op, i.e. on CUDA the FIR / bias_act / contraction kernels of
libsgv_b200.

Differences from the reference implementation, none of them numerical in fp32:
  * the demodulation coefficients come from [O, I] and [N, I] sized tensors (sum_k W^2 contracted with styles^2) instead of a
    materialised [N, O, I, k, k] product (302 MB per 512-channel layer at N = 32) whenever the per-sample weights themselves
    are not needed;
  * fp16 inputs get the reference's infinity-norm pre-normalisation (networks.py:50-52) unchanged.
"""
import numpy as np
import torch

from . import conv2d_resample, fma


def modulated_conv2d(x, weight, styles, noise=None, up=1, down=1, padding=0, resample_filter=None, demodulate=True,
                     flip_weight=True, fused_modconv=True):
    n = x.shape[0]
    o, i, kh, kw = weight.shape
    assert x.shape[1] == i and tuple(styles.shape) == (n, i)

    if x.dtype == torch.float16 and demodulate:
        weight = weight * (1 / np.sqrt(i * kh * kw) / weight.norm(float('inf'), dim=[1, 2, 3], keepdim=True))
        styles = styles / styles.norm(float('inf'), dim=1, keepdim=True)

    per_sample = None
    dcoefs = None
    if fused_modconv:
        per_sample = weight.unsqueeze(0) * styles.reshape(n, 1, i, 1, 1)                    # [N, O, I, kh, kw]
        if demodulate:
            dcoefs = (per_sample.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt()
            per_sample = per_sample * dcoefs.reshape(n, o, 1, 1, 1)
    elif demodulate:
        dcoefs = (styles.square() @ weight.square().sum(dim=[2, 3]).t() + 1e-8).rsqrt()     # [N, O]

    if not fused_modconv:
        y = x * styles.to(x.dtype).reshape(n, i, 1, 1)
        y = conv2d_resample.conv2d_resample(x=y, w=weight.to(x.dtype), f=resample_filter, up=up, down=down, padding=padding,
                                            flip_weight=flip_weight)
        if dcoefs is not None:
            d = dcoefs.to(x.dtype).reshape(n, o, 1, 1)
            y = fma.fma(y, d, noise.to(x.dtype)) if noise is not None else y * d
        elif noise is not None:
            y = y.add_(noise.to(x.dtype))
        return y

    # one grouped convolution with per-sample weights (eval mode, networks.py:76-86)
    y = conv2d_resample.conv2d_resample(x=x.reshape(1, n * i, *x.shape[2:]), w=per_sample.reshape(n * o, i, kh, kw).to(x.dtype),
                                        f=resample_filter, up=up, down=down, padding=padding, groups=n, flip_weight=flip_weight)
    y = y.reshape(n, o, *y.shape[2:])
    if noise is not None:
        y = y.add_(noise)
    return y
