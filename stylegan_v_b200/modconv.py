"""Fused modulated-convolution layer op for the native synthesis path (NHWC, fp32 storage, TF32 tensor-core math).

One call covers what the reference runs as 4-6 kernels per layer in training mode
(src/training/networks.py:64-74 + 141-143):

    x * styles  ->  conv2d / conv_transpose2d(stride 2) [-> upfirdn2d 4x4]  ->  * dcoefs  ->  + bias -> lrelu -> * gain -> clamp

  up = 1 : ONE tcgen05 implicit-GEMM launch (styles folded into the A-operand staging, dcoefs/bias/act in the epilogue)
  up = 2 : four polyphase implicit-GEMM launches write the (2h+1)x(2w+1) transposed-conv result, then ONE FIR
           launch whose epilogue applies dcoefs/bias/act (conv2d_resample.py:125-139 geometry: pad 1, gain 4).

Backward (first order; second order is served by the unfused drop-in ops in stylegan_v_b200/ops):
  activation gradient + bias gradient   bias_act CUDA kernel (grad=1) with the per-channel reduction fused
  data gradient                         the same implicit-GEMM kernel with transposed weights; dcoefs are folded
                                        into its A-operand staging (stride-2 variant for up = 2 via TMA element strides)
  weight gradient                       implicit-GEMM weight-gradient kernel (stylegan_v_b200/csrc/wgrad_tf32.cu) when
                                        available for the shape, else the cuDNN library call the reference uses
                                        (conv2d_gradfix.py:140-148)
  d styles, d dcoefs                    small reductions; dcoefs themselves are computed by differentiable torch ops
                                        outside this Function (see demod_coefs), so their dependence on styles / weight
                                        is handled by autograd on [N,C]-sized tensors.
"""
import os

import numpy as np
import torch

from . import _lib
from . import conv as _conv
from . import plugin as _plugin

_TAPS3 = _conv.TAPS_3x3
_OFFS3_FWD = [(ky - 1, kx - 1) for ky, kx in _TAPS3]          # correlation, padding 1
_OFFS3_DGRAD = [(1 - ky, 1 - kx) for ky, kx in _TAPS3]        # its adjoint
_FIR_1331 = None


class _Demod(torch.autograd.Function):
    """dcoefs from (weight, styles) in one launch, both gradients in one launch each (csrc/demod.cu).  First order."""

    @staticmethod
    def forward(ctx, weight, styles):
        O, I = weight.shape[0], weight.shape[1]
        N = styles.shape[0]
        if styles.stride(1) != 1 or styles.stride(0) % 4 != 0 or styles.data_ptr() % 16 != 0:
            styles = styles.contiguous()
        w = weight.contiguous()
        dc = torch.empty([N, O], dtype=torch.float32, device=weight.device)
        with torch.cuda.device(weight.device):
            _lib.check(_lib.lib().sgv_demod_fwd(w.data_ptr(), styles.data_ptr(), styles.stride(0), dc.data_ptr(), N, O, I, 9, 1e-8,
                                                _conv._stream(weight.device)), 'sgv_demod_fwd')
        ctx.save_for_backward(w, styles, dc)
        return dc

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, ddc):
        w, styles, dc = ctx.saved_tensors
        O, I = w.shape[0], w.shape[1]
        N = styles.shape[0]
        ddc = ddc.contiguous()
        dw = ds = None
        L, st = _lib.lib(), _conv._stream(w.device)
        with torch.cuda.device(w.device):
            if ctx.needs_input_grad[0]:
                dw = torch.empty_like(w)
                _lib.check(L.sgv_demod_bwd_weight(w.data_ptr(), styles.data_ptr(), styles.stride(0), dc.data_ptr(), ddc.data_ptr(), dw.data_ptr(),
                                                  N, O, I, 9, st), 'sgv_demod_bwd_weight')
            if ctx.needs_input_grad[1]:
                ds = torch.zeros([N, I], dtype=torch.float32, device=w.device)
                _lib.check(L.sgv_demod_bwd_styles(w.data_ptr(), styles.data_ptr(), styles.stride(0), dc.data_ptr(), ddc.data_ptr(), ds.data_ptr(), I,
                                                  N, O, I, 9, st), 'sgv_demod_bwd_styles')
        return dw, ds


def demod_coefs(weight, styles):
    """dcoefs[n,o] = rsqrt(sum_{i,k} (W[o,i,k] * s[n,i])^2 + 1e-8)   (networks.py:57-59) without materialising [N,O,I,k,k].
    CUDA fp32 3x3 weights: one launch (and one per gradient) of csrc/demod.cu; otherwise differentiable torch ops on [O, I]-sized tensors."""
    if (weight.is_cuda and weight.dtype == torch.float32 and styles.dtype == torch.float32 and tuple(weight.shape[2:]) == (3, 3)
            and weight.shape[1] % 4 == 0):
        return _Demod.apply(weight, styles)
    wsq = weight.square().sum(dim=[2, 3])                  # [O, I]
    return (styles.square() @ wsq.t() + 1e-8).rsqrt()      # [N, O]


def _phase_taps(a, b):
    """Taps of the stride-2 transposed 3x3 conv that land on output phase (a, b), with their input offsets."""
    taps = [(ky, kx) for ky in range(a, 3, 2) for kx in range(b, 3, 2)]
    offs = [(-((ky - a) // 2), -((kx - b) // 2)) for ky, kx in taps]
    return taps, offs


def _nhwc(t):
    return t.contiguous(memory_format=torch.channels_last)


def _fir(device):
    global _FIR_1331
    if _FIR_1331 is None or _FIR_1331.device != device:
        k = torch.tensor([1.0, 3.0, 3.0, 1.0], dtype=torch.float32)
        k = torch.outer(k, k)
        _FIR_1331 = (k / k.sum()).to(device)
    return _FIR_1331


def _wgrad_native(gout, x, styles, dscale, w_shape, up, g_ready=False, x3=False):
    """Weight gradient on the tcgen05 split-K kernel (csrc/wgrad_tf32.cu) with styles / dcoefs folded into operand staging."""
    O, I, kh, kw = w_shape
    N, _, H, W = x.shape
    if up == 1:
        taps_x = _OFFS3_FWD if kh == 3 else [(0, 0)]
        dw = _conv.igemm_wgrad(gout, x, [(0, 0)] * len(taps_x), taps_x, (H, W), g_scale=dscale, x_scale=styles, g_ready=g_ready, x3=x3)
    elif H >= 8 and W >= 8:
        # transposed (stride-2) conv: per polyphase sub-lattice (a, b) of the output gradient, a stride-1 correlation between the
        # unshifted input (M side) and the pixel-strided VIEW gout[:, :, a::2, b::2] shifted by (ky//2, kx//2) (N side): the grouped-tap
        # kernel applies; it returns dW^T [tap][I][O]
        dwt = torch.zeros([kh * kw, I, O], dtype=torch.float32, device=x.device)       # all four calls accumulate into their tap slots
        for a in (0, 1):
            for b in (0, 1):
                taps = [(ky, kx) for ky in range(a, 3, 2) for kx in range(b, 3, 2)]
                offs = [(ky // 2, kx // 2) for ky, kx in taps]
                _conv.igemm_wgrad(x, gout[:, :, a::2, b::2], [(0, 0)] * len(taps), offs, (H, W), g_scale=styles, x_scale=dscale,
                                  out=dwt, slots=[ky * kw + kx for ky, kx in taps], x_ready=g_ready, x3=x3)
        return dwt.reshape(kh, kw, I, O).permute(3, 2, 0, 1)
    else:
        dw = _conv.igemm_wgrad(gout, x, _TAPS3, [(0, 0)] * 9, (H, W), g_stride=2, g_scale=dscale, x_scale=styles, x3=x3)
    return dw.reshape(kh, kw, O, I).permute(2, 3, 0, 1)


USE_NATIVE_WGRAD = True      # False -> cuDNN library call (kept for A/B measurements)
PRESCALE_GRADIENT = os.environ.get('SGV_PRESCALE', '1') != '0'     # fold dcoefs + TF32 rounding into the activation-gradient pass (see backward)


def prepare_weights(weight, up, flip_weight):
    """Tap-major TF32 weight slabs of one layer for the forward and the data-gradient launches (non-differentiable; the weight
    gradient is produced directly by the wgrad kernel).  Depends on the weight only, so a caller may build it ahead of the layer
    (SynthesisNetwork does, on its parameter stream) and hand it to fused_modulated_conv(prep=...).  In the tf32x3 precision mode
    (stylegan_v_b200.precision) the slabs carry their TF32 residuals too ([2, ntaps, ., .]); the layer then runs fp32-grade."""
    O, I, kh, kw = weight.shape
    # a flipped kernel is the same weight read with mirrored tap indices: no flip kernel
    flipped = (not flip_weight) if up == 1 else flip_weight      # conv2d_resample.py:35-36 (up = 1) / :138 (transposed conv: flag inverted)
    mirror = (lambda t: (kh - 1 - t[0], kw - 1 - t[1])) if flipped else (lambda t: t)
    base = _TAPS3 if kh == 3 else [(0, 0)]
    if up == 1:
        groups = [base]
    else:
        groups = [_phase_taps(a, b)[0] for a in (0, 1) for b in (0, 1)]
    order = [t for g in groups for t in g]
    pair_ok = weight.is_contiguous() and O % 32 == 0 and I % 32 == 0
    if pair_ok:
        wf, wd = _conv.prep_weights_pair(weight, [mirror(t) for t in order], [mirror(t) for t in base])
    else:
        wsrc = weight.flip([2, 3]) if flipped else weight
        wf = _conv.prep_weights(wsrc, order)
        wd = _conv.prep_weights(wsrc, base, rows_dim=1, cols_dim=0)
    fwd, start = [], 0
    for g in groups:
        fwd.append(_conv.slab_taps(wf, start, start + len(g)))
        start += len(g)
    return dict(fwd=fwd, dgrad=wd)


class WeightGradBox:
    """Mailbox between a fused layer node and the WeightGradNode of its weight (see WeightGradNode)."""
    __slots__ = ('job', 'stream')

    def __init__(self, stream):
        self.job, self.stream = None, stream


class WeightGradNode(torch.autograd.Function):
    """Identity on a layer weight, applied on the stream that should run the layer's weight-gradient contraction.

    Autograd runs a node's backward on the stream its forward ran on and orders it after the producers of its incoming gradients.
    Routing `weight` through this node on the network's parameter stream therefore moves the (tensor-core bound) weight-gradient
    kernels off the activation-gradient stream: the fused layer's backward posts the contraction as a job in the box and hands a
    shape-only placeholder down; this node executes the job and returns the real gradient.  Results are identical; only the
    stream the kernels are issued on changes."""

    @staticmethod
    def forward(ctx, weight, box):
        ctx.box = box
        return weight.view_as(weight)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad):
        job, ctx.box.job = ctx.box.job, None
        return (job() if job is not None else grad), None


_PLACEHOLDERS = {}


def _placeholder(like):
    """A zero-stride tensor of `like`'s shape (no memory traffic): stands in for a gradient that a later node fills in."""
    key = (like.device, like.dtype)
    if key not in _PLACEHOLDERS:
        _PLACEHOLDERS[key] = torch.zeros(1, dtype=like.dtype, device=like.device)
    return _PLACEHOLDERS[key].expand(like.shape)


class _FusedModConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, styles, dcoefs, bias, up, act, gain, flip_weight, prep, wmod=None, rgb_bias=None, wbox=None, noise=None):
        assert x.is_cuda and x.dtype == torch.float32, 'the fused layer op is CUDA / float32 only (no CPU path)'
        x = _nhwc(x)
        O, I, kh, kw = weight.shape
        N, _, H, W = x.shape
        assert kh == kw and kh in (1, 3)
        if prep is None:
            prep = prepare_weights(weight, up, flip_weight)
        ctx.x3 = prep['dgrad'].ndim == 4          # the precision mode the slabs were prepared in; backward follows it
        if noise is not None:
            noise = noise.detach().to(torch.float32).contiguous()
        if up == 1:
            offs = _OFFS3_FWD if kh == 3 else [(0, 0)]
            y = _conv.igemm_conv(x, prep['fwd'][0], offs, a_scale=styles, o_scale=dcoefs, bias=bias, act=act, gain=gain, noise=noise)
        else:
            assert up == 2 and kh == 3
            u = torch.empty([N, O, 2 * H + 1, 2 * W + 1], dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
            for j, (a, b) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
                _conv.igemm_conv(x, prep['fwd'][j], _phase_taps(a, b)[1], a_scale=styles, out_view=u[:, :, a::2, b::2])
            y = _plugin.upfirdn2d(u, _fir(x.device), 1, 1, 1, 1, 1, 1, 1, 1, False, 4.0,
                                  epilogue=dict(scale=dcoefs, bias=bias, act=act, alpha=0.2, gain=gain, clamp=None, noise=noise))
        ctx.save_for_backward(x, weight, styles, dcoefs if dcoefs is not None else x.new_empty(0), bias if bias is not None else x.new_empty(0), y,
                              noise if noise is not None else x.new_empty(0))
        ctx.has_noise = noise is not None
        ctx.cfg = (up, act, gain, flip_weight, dcoefs is not None, bias is not None)
        ctx.wp_dgrad = prep['dgrad']
        ctx.wbox = wbox
        if wmod is None:
            ctx.wmod = None
            return y
        # ToRGB branch on the same activation (networks.py:262-265): second output, its backward is folded into this node's
        ctx.wmod = wmod.detach()
        ctx.set_materialize_grads(False)
        return y, _conv.torgb_fwd(y, wmod, rgb_bias)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy, drgb=None):
        x, weight, styles, dcoefs, bias, y, noise = ctx.saved_tensors
        up, act, gain, flip_weight, has_d, has_b = ctx.cfg
        O, I, kh, kw = weight.shape
        N, _, H, W = x.shape
        dy = _nhwc(dy) if dy is not None else None
        # ---- activation (+ bias) gradient dz, bias gradient and the dcoefs reduction in ONE pass over (dy, y); with a ToRGB
        #      branch also its data gradient (added to dy on the fly) and d(wmod) ----
        want_db = has_b and ctx.needs_input_grad[4]
        want_dd = has_d and ctx.needs_input_grad[3]
        dwmod = drgb_bias = None
        # PRESCALE: dz leaves the kernel as tf32_rn(dz * dcoefs) (the FIR adjoint of the up layers re-rounds its result), i.e. exactly the
        # operand both contractions below need — they then skip scaling, and the weight-gradient kernel skips the staging pass of
        # that operand altogether (`ready`).  db / dd are reduced from the unscaled gradient inside the same kernel.
        native_w = USE_NATIVE_WGRAD and I % 32 == 0 and O % 32 == 0
        # (not in the tf32x3 mode, whose contractions split the UNROUNDED gradient themselves, and not with a noise input, whose own
        #  gradient needs the unscaled dz)
        x3 = ctx.x3
        prescale = PRESCALE_GRADIENT and has_d and native_w and not x3 and not ctx.has_noise
        osc = dcoefs if prescale else None
        if ctx.wmod is not None and drgb is not None:
            dz, db, dd, dwmod = _conv.act_bwd(dy, y, bias if has_b else None, act, gain, want_db, want_dd, dyimg=drgb, wmod=ctx.wmod, oscale=osc)
            drgb_bias = drgb.sum(dim=[0, 2, 3])
        else:
            if dy is None:
                return (None,) * 14
            dz, db, dd = _conv.act_bwd(dy, y, bias if has_b else None, act, gain, want_db, want_dd, oscale=osc)
        dnoise = None
        if ctx.has_noise:
            # y = act(conv * dcoefs + noise + bias): the pre-bias value the kernel recovers from y contains the noise, so its part of the
            # dcoefs reduction is taken out again, and d(noise)[n, hw] = sum_c dz (fma.py:36-58).  Two extra passes over dz, only in this
            # (non-default: configs/model/stylegan-v.yaml use_noise = false) mode.
            nz = noise.reshape(-1, 1, y.shape[2], y.shape[3])
            if want_dd:
                dd = dd - (dz * nz).sum(dim=[2, 3])
            if ctx.needs_input_grad[13]:
                dnoise = dz.sum(dim=1, keepdim=True)
                if nz.shape[0] == 1 and N > 1:
                    dnoise = dnoise.sum(dim=0, keepdim=True)
        ddcoefs = dd / dcoefs if want_dd else None
        dscale = dcoefs if (has_d and not prescale) else None
        wp = ctx.wp_dgrad
        # ---- data gradient: ONE launch gives dx = dxs * styles (epilogue scale) and dstyles = sum_hw dxs * x (fused reduction) ----
        want_ds = ctx.needs_input_grad[2]
        ds = torch.zeros([N, I], dtype=torch.float32, device=x.device) if want_ds else None
        red = dict(red_x=x, red_out=ds) if want_ds else {}
        if up == 1:
            gout = dz
            offs = _OFFS3_DGRAD if kh == 3 else [(0, 0)]
            dx = _conv.igemm_conv(dz, wp, offs, a_scale=dscale, o_scale=styles, a_ready=prescale, **red)
        else:
            # adjoint of the FIR pass (upfirdn2d.py:246-261): padding (fw - p - 1) = 2, flipped filter, same gain
            gout = _plugin.upfirdn2d(dz, _fir(x.device), 1, 1, 1, 1, 2, 2, 2, 2, True, 4.0,
                                     epilogue=dict(act='linear', round_tf32=True) if prescale else None)
            # data gradient of the stride-2 transposed conv = stride-2 correlation: ONE launch with TMA element strides.
            # (Measured alternative: 4 accumulate-launches over polyphase views of `gout` on the halo-patch kernel — 1.2 ms/step
            #  slower at config 2; kept available through igemm_conv(accumulate=True).)
            dx = _conv.igemm_conv(gout, wp, _TAPS3, out_hw=(H, W), in_stride=2, a_scale=dscale, o_scale=styles, a_ready=prescale, **red)
        # ---- weight gradient ----
        def weight_grad():
            if native_w:
                dw = _wgrad_native(gout, x, styles, dscale, (O, I, kh, kw), up, g_ready=prescale, x3=x3)
            else:
                xs = x * styles.reshape(N, I, 1, 1)
                g = gout * dscale.reshape(N, O, 1, 1) if has_d else gout
                if up == 1:
                    dw = _wgrad_library(g, xs, (O, I, kh, kw), False, 1)
                else:
                    dw = _wgrad_library(g, xs, (I, O, kh, kw), True, 2).transpose(0, 1)
            if (up == 1 and not flip_weight) or (up == 2 and flip_weight):
                dw = dw.flip([2, 3])
            return dw
        dw = None
        if ctx.needs_input_grad[1]:
            box = ctx.wbox
            if box is None:
                dw = weight_grad()
            else:
                # deferred: the WeightGradNode that fed `weight` into this op runs the contraction when the engine reaches it, on the
                # stream IT was created on (the parameter stream) — i.e. concurrently with the next layers' HBM-bound gradient kernels
                # on this stream.  The node receives a shape-only placeholder and substitutes the real gradient.
                for t in (gout, x, styles, dscale):
                    if t is not None and box.stream is not None:
                        t.record_stream(box.stream)
                box.job = weight_grad
                dw = _placeholder(weight)
        return dx, dw, ds, ddcoefs, db, None, None, None, None, None, dwmod, drgb_bias, None, dnoise


def fused_modulated_conv(x, weight, styles, bias=None, up=1, demodulate=True, act='lrelu', gain=None, flip_weight=True, dcoefs=None, prep=None,
                         torgb_wmod=None, torgb_bias=None, wbox=None, noise=None):
    """y = clamp-free bias_act(modulated_conv2d(x, weight, styles, noise, up, demodulate), bias, act, gain) on NHWC fp32 tensors.

    Equivalent (up to TF32 rounding of the contraction operands; fp32-grade in the tf32x3 precision mode) to the reference's
    training-mode sequence modulated_conv2d(..., fused_modconv=False) + bias_act (networks.py:30-86,141-143).
    noise: [N or 1, 1, H_out, W_out] plane(s) already multiplied by the noise strength (networks.py:130-134), added inside the
    contraction's epilogue (up = 1) / the FIR pass's epilogue (up = 2) between the demodulation and the bias."""
    if gain is None:
        gain = float(np.sqrt(2)) if act == 'lrelu' else 1.0
    if dcoefs is None and demodulate:
        dcoefs = demod_coefs(weight, styles)
    if torgb_wmod is not None:
        # -> (y, rgb): rgb[n,j,hw] = sum_c y[n,hw,c] * torgb_wmod[n,j,c] + torgb_bias[j]  (ToRGBLayer arithmetic, one autograd node)
        return _FusedModConv.apply(x, weight, styles, dcoefs, bias, up, act, float(gain), flip_weight, prep, torgb_wmod, torgb_bias, wbox, noise)
    return _FusedModConv.apply(x, weight, styles, dcoefs, bias, up, act, float(gain), flip_weight, prep, None, None, wbox, noise)
