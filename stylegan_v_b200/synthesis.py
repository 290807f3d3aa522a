"""Native (B200) synthesis network of StyleGAN-V: same parameters and results as the reference's
SynthesisLayer / ToRGBLayer / SynthesisBlock / SynthesisNetwork (src/training/networks.py:90-366, 'skip'
architecture, concat_const time conditioning, temporal input — the stylegan-v.yaml model), executed on the fused
sm_100a kernels of libsgv_b200 with NHWC activations.

State-dict keys equal the reference's (`b{res}.conv{0,1}.{weight,bias,affine.weight,affine.bias}`,
`b{res}.torgb.*`, `b4.input.input.const`, `motion_encoder.*`, `*.resample_filter`), so a reference checkpoint's
`G.synthesis.state_dict()` loads with `load_state_dict`.

What differs from the reference is how a layer is executed (see stylegan_v_b200/modconv.py): per layer one
implicit-GEMM launch (plus one FIR launch for up=2) instead of x*styles, cuDNN conv, upfirdn2d, *dcoefs, bias_act;
all style affines of a forward pass are evaluated as ONE stacked GEMM up front.
"""
import contextlib
import os

import numpy as np
import torch

from . import _lib
from . import conv as _conv
from . import dense as _dense
from .modconv import fused_modulated_conv, demod_coefs, prepare_weights, WeightGradBox, WeightGradNode
from .ops import upfirdn2d as _upfirdn2d
from .ops import bias_act as _bias_act
from .ops.modulated_conv import modulated_conv2d as _modulated_conv2d
from .time_encoder import EqualizedLinear, MotionMappingNetwork


def _setup_filter(taps):
    return _upfirdn2d.setup_filter(taps)


class SynthesisLayer(torch.nn.Module):
    def __init__(self, in_channels, out_channels, w_dim, resolution, kernel_size=3, up=1, resample_filter=(1, 3, 3, 1), use_noise=False):
        super().__init__()
        self.resolution, self.up, self.use_noise = resolution, up, use_noise
        self.register_buffer('resample_filter', _setup_filter(list(resample_filter)))
        self.affine = EqualizedLinear(w_dim, in_channels, bias_init=1)
        self.weight = torch.nn.Parameter(torch.randn(out_channels, in_channels, kernel_size, kernel_size))
        if use_noise:                                   # networks.py:119-121 (same registration order => same state dict / RNG draws)
            self.register_buffer('noise_const', torch.randn([resolution, resolution]))
            self.noise_strength = torch.nn.Parameter(torch.zeros([]))
        self.bias = torch.nn.Parameter(torch.zeros(out_channels))

    def make_noise(self, batch, noise_mode, device):
        """The noise input of modulated_conv2d (networks.py:130-134): [N,1,res,res] fresh Gaussian draws ('random') or the stored
        plane ('const'), times the learned strength; None without use_noise / for 'none'."""
        assert noise_mode in ('random', 'const', 'none')
        if not self.use_noise or noise_mode == 'none':
            return None
        if noise_mode == 'random':
            return torch.randn([batch, 1, self.resolution, self.resolution], device=device) * self.noise_strength
        return (self.noise_const * self.noise_strength).reshape(1, 1, self.resolution, self.resolution)

    def plan(self, styles, async_wgrad_stream=None):
        """Everything of this layer that depends on parameters and styles only (not on activations): demodulation coefficients
        and the TF32 weight slabs.  SynthesisNetwork evaluates the plans of all layers on its parameter stream."""
        plan = dict(styles=styles, dcoefs=demod_coefs(self.weight, styles), prep=prepare_weights(self.weight, self.up, self.up == 1),
                    weight=self.weight, wbox=None)
        if async_wgrad_stream is not None and self.weight.requires_grad and torch.is_grad_enabled():
            # the weight-gradient contraction of this layer will be issued on the parameter stream (see modconv.WeightGradNode)
            plan['wbox'] = WeightGradBox(async_wgrad_stream)
            plan['weight'] = WeightGradNode.apply(self.weight, plan['wbox'])
        return plan

    def forward(self, x, w=None, styles=None, gain=1.0, plan=None, torgb_wmod=None, torgb_bias=None, noise_mode='random', noise=None):
        """With torgb_wmod [N,3,C] / torgb_bias [3] the block's ToRGB layer is evaluated in the same autograd node: returns (x, rgb).
        noise: explicit [N|1,1,res,res] noise input (already scaled) overriding noise_mode (tests feed the oracle's draws)."""
        if plan is None:
            plan = dict(styles=styles if styles is not None else self.affine(w), dcoefs=None, prep=None, weight=self.weight, wbox=None)
        if noise is None:
            noise = self.make_noise(x.shape[0], noise_mode, x.device)
        return fused_modulated_conv(x, plan['weight'], plan['styles'], self.bias, up=self.up, demodulate=True, act='lrelu',
                                    gain=float(np.sqrt(2)) * gain, flip_weight=(self.up == 1), dcoefs=plan['dcoefs'], prep=plan['prep'],
                                    torgb_wmod=torgb_wmod, torgb_bias=torgb_bias, wbox=plan['wbox'], noise=noise)


def _layer_unfused(layer, x, w, fused_modconv, gain=1.0, conv_clamp=None, noise_mode='random'):
    """SynthesisLayer.forward of the reference (networks.py:124-144) on the drop-in ops."""
    styles = layer.affine(w)
    noise = layer.make_noise(x.shape[0], noise_mode, x.device)
    x = _modulated_conv2d(x=x, weight=layer.weight, styles=styles, noise=noise, up=layer.up, padding=1, resample_filter=layer.resample_filter,
                          flip_weight=(layer.up == 1), fused_modconv=fused_modconv)
    clamp = conv_clamp * gain if conv_clamp is not None else None
    return _bias_act.bias_act(x, layer.bias.to(x.dtype), act='lrelu', gain=float(np.sqrt(2)) * gain, clamp=clamp)


def _torgb_unfused(layer, x, w, fused_modconv, conv_clamp=None):
    """ToRGBLayer.forward of the reference (networks.py:159-163) on the drop-in ops."""
    styles = layer.affine(w) * layer.weight_gain
    x = _modulated_conv2d(x=x, weight=layer.weight, styles=styles, demodulate=False, fused_modconv=fused_modconv)
    return _bias_act.bias_act(x, layer.bias.to(x.dtype), clamp=conv_clamp)


class _ToRGB(torch.autograd.Function):
    """y[n,j,hw] = sum_c x[n,hw,c] * wmod[n,j,c] + b[j] on the NHWC activation (csrc/layer_elementwise.cu)."""

    @staticmethod
    def forward(ctx, x, wmod, bias):
        x = x.contiguous(memory_format=torch.channels_last)
        ctx.save_for_backward(x, wmod)
        return _conv.torgb_fwd(x, wmod, bias)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        x, wmod = ctx.saved_tensors
        dx, dwmod = _conv.torgb_bwd(dy, x, wmod)
        return dx, dwmod, dy.sum(dim=[0, 2, 3])


class _ToRgbWmod(torch.autograd.Function):
    """wmod[n, j, c] = weight[j, c] * styles[n, c] * gain (networks.py:159-160) and both gradients: csrc/layer_elementwise.cu.  First order."""

    @staticmethod
    def forward(ctx, weight, styles, gain):
        J, C = weight.shape[0], weight.shape[1]
        N = styles.shape[0]
        w2 = weight.reshape(J, C).contiguous()
        wmod = torch.empty([N, J, C], dtype=torch.float32, device=styles.device)
        with torch.cuda.device(styles.device):
            _lib.check(_lib.lib().sgv_torgb_wmod_fwd(w2.data_ptr(), styles.data_ptr(), styles.stride(0), wmod.data_ptr(), N, C, J, gain,
                                                     _conv._stream(styles.device)), 'sgv_torgb_wmod_fwd')
        ctx.save_for_backward(w2, styles)
        ctx.cfg = (gain, tuple(weight.shape))
        return wmod

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dwmod):
        w2, styles = ctx.saved_tensors
        gain, w_shape = ctx.cfg
        J, C = w2.shape
        N = styles.shape[0]
        dwmod = dwmod.contiguous()
        ds = torch.empty([N, C], dtype=torch.float32, device=styles.device) if ctx.needs_input_grad[1] else None
        dw = torch.empty([J, C], dtype=torch.float32, device=styles.device) if ctx.needs_input_grad[0] else None
        with torch.cuda.device(styles.device):
            _lib.check(_lib.lib().sgv_torgb_wmod_bwd(dwmod.data_ptr(), w2.data_ptr(), styles.data_ptr(), styles.stride(0), ds.data_ptr() if ds is not None else None,
                                                     dw.data_ptr() if dw is not None else None, N, C, J, gain, _conv._stream(styles.device)), 'sgv_torgb_wmod_bwd')
        return (dw.reshape(w_shape) if dw is not None else None), ds, None


class ToRGBLayer(torch.nn.Module):
    """1x1 modulated conv to RGB without demodulation + bias (networks.py:148-163).  With 3 output channels this is
    memory-bound; it runs as one batched [HW, C] x [C, 3] product per sample straight from the NHWC activations."""

    def __init__(self, in_channels, out_channels, w_dim):
        super().__init__()
        self.affine = EqualizedLinear(w_dim, in_channels, bias_init=1)
        self.weight = torch.nn.Parameter(torch.randn(out_channels, in_channels, 1, 1))
        self.bias = torch.nn.Parameter(torch.zeros(out_channels))
        self.weight_gain = 1 / np.sqrt(in_channels)

    def plan(self, styles):
        C = self.weight.shape[1]
        if styles.is_cuda and styles.dtype == torch.float32 and self.weight.dtype == torch.float32 and self.weight.shape[0] <= 4 and styles.stride(1) == 1:
            return dict(wmod=_ToRgbWmod.apply(self.weight, styles, float(self.weight_gain)))          # [N, 3, C]: one launch, one in the backward
        return dict(wmod=self.weight.reshape(1, -1, C) * (styles * self.weight_gain).unsqueeze(1))   # [N, 3, C]  (tiny, differentiable torch ops)

    def forward(self, x, w=None, styles=None, plan=None):
        if plan is None:
            plan = self.plan(styles if styles is not None else self.affine(w))
        wmod = plan['wmod']
        N, C, H, W = x.shape
        if wmod.shape[1] == 3 and x.is_cuda:
            return _ToRGB.apply(x, wmod, self.bias)                                        # one pass over the NHWC activation
        xf = x.permute(0, 2, 3, 1).reshape(N, H * W, C)
        y = torch.baddbmm(self.bias.reshape(1, 1, -1), xf, wmod.transpose(1, 2))
        return y.reshape(N, H, W, -1).permute(0, 3, 1, 2).contiguous()                     # NCHW fp32 like the reference (networks.py:261)


class _TemporalInput(torch.nn.Module):
    def __init__(self, channel_dim, motion_v_dim):
        super().__init__()
        self.motion_v_dim = motion_v_dim
        self.const = torch.nn.Parameter(torch.randn(1, channel_dim, 4, 4))

    def get_dim(self):
        return self.motion_v_dim + self.const.shape[1]

    def forward(self, motion_v):
        n = motion_v.shape[0]
        x = torch.cat([self.const.expand(n, -1, -1, -1), motion_v[:, :, None, None].expand(-1, -1, 4, 4)], dim=1)
        return x.contiguous(memory_format=torch.channels_last)


class _GenInput(torch.nn.Module):
    def __init__(self, channel_dim, motion_v_dim):
        super().__init__()
        self.input = _TemporalInput(channel_dim, motion_v_dim)
        self.total_dim = self.input.get_dim()

    def forward(self, motion_v):
        return self.input(motion_v)


class SynthesisBlock(torch.nn.Module):
    def __init__(self, in_channels, out_channels, w_dim, motion_v_dim, resolution, img_channels, resample_filter=(1, 3, 3, 1),
                 use_fp16=False, conv_clamp=None, use_noise=False):
        super().__init__()
        self.in_channels, self.resolution = in_channels, resolution
        self.use_fp16, self.conv_clamp = use_fp16, conv_clamp            # mixed-precision mode of the reference (train.py:173-174); unfused path only
        self.register_buffer('resample_filter', _setup_filter(list(resample_filter)))
        self.num_conv = 0
        if in_channels == 0:
            self.input = _GenInput(out_channels, motion_v_dim)
            conv1_in = self.input.total_dim
        else:
            self.conv0 = SynthesisLayer(in_channels, out_channels, w_dim, resolution, up=2, resample_filter=resample_filter, use_noise=use_noise)
            self.num_conv += 1
            conv1_in = out_channels
        self.conv1 = SynthesisLayer(conv1_in, out_channels, w_dim, resolution, resample_filter=resample_filter, use_noise=use_noise)
        self.num_conv += 1
        self.torgb = ToRGBLayer(out_channels, img_channels, w_dim)
        self.num_torgb = 1

    def layers(self):
        return ([] if self.in_channels == 0 else [self.conv0]) + [self.conv1, self.torgb]

    def forward(self, x, img, plans, motion_v=None, noise_mode='random'):
        """plans: list of per-layer plans in layer order (conv0?, conv1, torgb); each is a callable returning the plan dict
        (it makes the compute stream wait for the parameter stream first)."""
        it = iter(plans)
        if self.in_channels == 0:
            x = self.input(motion_v)
        else:
            x = self.conv0(x, plan=next(it)(), noise_mode=noise_mode)
        conv1_plan, rgb_plan = next(it)(), next(it)()
        if rgb_plan['wmod'].shape[1] == 3:
            x, y = self.conv1(x, plan=conv1_plan, torgb_wmod=rgb_plan['wmod'], torgb_bias=self.torgb.bias, noise_mode=noise_mode)    # conv1 + ToRGB: one node
        else:
            x = self.conv1(x, plan=conv1_plan, noise_mode=noise_mode)
            y = self.torgb(x, plan=rgb_plan)
        if img is not None:
            img = _upfirdn2d.upsample2d(img, self.resample_filter)
        img = img.add_(y) if img is not None else y
        return x, img

    def forward_unfused(self, x, img, ws, motion_v=None, fused_modconv=None, force_fp32=False, noise_mode='random'):
        """SynthesisBlock.forward of the reference (networks.py:224-266, 'skip' architecture) layer by layer on the drop-in ops:
        differentiable to any order, runnable on CPU tensors, and the home of the reference's mixed-precision mode (fp16 activations +
        conv_clamp in the `use_fp16` blocks).  ws [N, num_conv + num_torgb, w_dim]."""
        w_iter = iter(ws.unbind(dim=1))
        dtype = torch.float16 if self.use_fp16 and not force_fp32 else torch.float32
        if fused_modconv is None:               # networks.py:232: per-sample-weight conv in eval mode, unless fp16 with a batch
            fused_modconv = (not self.training) and (dtype == torch.float32 or (x is not None and int(x.shape[0]) == 1))
        if self.in_channels == 0:
            x = self.input(motion_v)
            if not x.is_cuda:
                x = x.contiguous()              # CUDA keeps the NHWC layout of the kernels: the drop-in ops hand back their input's memory format
        else:
            x = _layer_unfused(self.conv0, x.to(dtype), next(w_iter), fused_modconv, conv_clamp=self.conv_clamp, noise_mode=noise_mode)
        x = _layer_unfused(self.conv1, x, next(w_iter), fused_modconv, conv_clamp=self.conv_clamp, noise_mode=noise_mode)
        if img is not None:
            img = _upfirdn2d.upsample2d(img, self.resample_filter)
        y = _torgb_unfused(self.torgb, x, next(w_iter), fused_modconv, conv_clamp=self.conv_clamp)
        y = y.to(dtype=torch.float32, memory_format=torch.contiguous_format)
        img = img.add_(y) if img is not None else y
        return x, img


class SynthesisNetwork(torch.nn.Module):
    def __init__(self, w_dim=512, img_resolution=256, img_channels=3, channel_base=16384, channel_max=512,
                 motion_z_dim=512, motion_v_dim=512, motion_kernel_size=11, motion_z_distance=16, time_enc_dim=256,
                 min_period_len=16, max_period_len=1024, max_num_frames=1024, resample_filter=(1, 3, 3, 1), num_fp16_res=0, conv_clamp=None,
                 use_noise=False):
        assert img_resolution >= 4 and img_resolution & (img_resolution - 1) == 0
        super().__init__()
        # the reference's mixed-precision mode (fp16 in the num_fp16_res highest resolutions + activation clamp, train.py:173-174,
        # networks.py:292,305).  The fused NHWC layers are fp32-storage / TF32-math only, so a network built with it runs the unfused ops.
        self.mixed_precision = num_fp16_res > 0 or conv_clamp is not None
        fp16_resolution = max(2 ** (int(np.log2(img_resolution)) + 1 - num_fp16_res), 8)
        self.w_dim, self.img_resolution, self.img_channels = w_dim, img_resolution, img_channels
        self.block_resolutions = [2 ** i for i in range(2, int(np.log2(img_resolution)) + 1)]
        self.motion_encoder = MotionMappingNetwork(motion_z_dim, motion_v_dim, motion_kernel_size, motion_z_distance,
                                                   time_enc_dim, min_period_len, max_period_len, max_num_frames)
        self.motion_v_dim = self.motion_encoder.get_dim()
        ch = {res: min(channel_base // res, channel_max) for res in self.block_resolutions}
        self.num_ws = 0
        for res in self.block_resolutions:
            block = SynthesisBlock(ch[res // 2] if res > 4 else 0, ch[res], w_dim, self.motion_v_dim, res, img_channels, resample_filter,
                                   use_fp16=(res >= fp16_resolution), conv_clamp=conv_clamp, use_noise=use_noise)
            self.num_ws += block.num_conv
            if res == img_resolution:
                self.num_ws += block.num_torgb
            setattr(self, f'b{res}', block)

    @classmethod
    def from_config(cls, cfg):
        """cfg: anything with the attributes of oracle-side SynthesisConfig / the reference's yaml values."""
        return cls(w_dim=cfg.w_dim, img_resolution=cfg.img_resolution, img_channels=cfg.img_channels, channel_base=cfg.channel_base,
                   channel_max=cfg.channel_max, motion_z_dim=cfg.motion_z_dim, motion_v_dim=cfg.motion_v_dim,
                   motion_kernel_size=cfg.motion_kernel_size, motion_z_distance=cfg.motion_z_distance, time_enc_dim=cfg.time_enc_dim,
                   min_period_len=cfg.min_period_len, max_period_len=cfg.max_period_len, max_num_frames=cfg.max_num_frames,
                   resample_filter=tuple(cfg.resample_filter), use_noise=bool(getattr(cfg, 'use_noise', False)))

    def affine_layout(self):
        """Which w row every style affine reads, and the column layout of the stacked affine product: returns (groups, order, layers, col) with
        groups {w index: [layers]}, order = the w indices ascending, layers = all layers in that order, col[g] .. col[g + 1] = the columns of the
        stacked output that belong to order[g].  Layer l of block b reads ws[:, w_idx(b) + l], and a block's ToRGB shares its row with the next
        block's first conv (networks.py:350-357)."""
        groups, w_idx = {}, 0
        for res in self.block_resolutions:
            block = getattr(self, f'b{res}')
            for j, layer in enumerate(block.layers()):
                groups.setdefault(w_idx + j, []).append(layer)
            w_idx += block.num_conv
        order = sorted(groups)
        layers = [l for wi in order for l in groups[wi]]
        col = [0]
        for wi in order:
            col.append(col[-1] + sum(l.affine.weight.shape[0] for l in groups[wi]))
        return groups, order, layers, col

    def _all_styles(self, ws):
        """Evaluates every style affine of the forward pass (networks.py:124-126,159-160 for all layers) up front.
        Layer l of block b reads ws[:, w_idx(b) + l] (networks.py:350-357).  CUDA fp32: ONE launch of the exact-fp32 dense kernel for all
        layers — the affine weights are stacked along the output dimension in w-index order and column group g reads ws[:, g, :]
        (stylegan_v_b200/dense.py::stacked_affine; the tcgen05 kernels are the wrong tool here: M = 32 rows fill a quarter of one 128-row
        tile and the K = 512 loop is latency-bound — 1.3 ms per step measured against 0.6 ms for cuBLAS, profiles/launches_r2b_summary.txt).
        Otherwise one library GEMM per distinct w index."""
        groups, order, layers, col = self.affine_layout()
        out = {}
        first = layers[0].affine
        if ws.is_cuda and ws.dtype == torch.float32 and first.weight.dtype == torch.float32 and ws.shape[2] % 4 == 0 \
                and all(l.affine.weight.shape[0] % 8 == 0 for l in layers):
            key = (str(ws.device), ws.shape[1], ws.shape[2])
            if getattr(self, '_affine_groups_key', None) != key:
                self._affine_groups = _dense.make_groups(col, order, ws.shape[2], ws.device)
                self._affine_groups_key = key
            wcat = torch.cat([l.affine.weight for l in layers], dim=0)
            bcat = torch.cat([l.affine.bias for l in layers], dim=0)
            s = _dense.stacked_affine(ws, wcat, bcat, self._affine_groups, first.weight_gain)
            for l, piece in zip(layers, s.split([l.affine.weight.shape[0] for l in layers], dim=1)):
                out[id(l)] = piece
            return out
        for wi, layers in groups.items():
            wcat = torch.cat([l.affine.weight for l in layers], dim=0) * layers[0].affine.weight_gain
            bcat = torch.cat([l.affine.bias for l in layers], dim=0)
            s = torch.addmm(bcat.unsqueeze(0), ws[:, wi], wcat.t())
            for l, piece in zip(layers, s.split([l.affine.weight.shape[0] for l in layers], dim=1)):
                out[id(l)] = piece
        return out

    def forward(self, ws, t, c=None, motion_z=None, motion_v=None, t_max=None, unfused=None, fused_modconv=None, noise_mode='random'):
        """ws [B, num_ws, w_dim], t [B, F] -> img [B*F, 3, R, R] (fp32, NCHW) — networks.py:324-366 semantics.

        unfused=None (default): CUDA inputs run the fused NHWC layers (first-order differentiable); CPU inputs, or unfused=True,
        run the layer-by-layer formulation on the drop-in ops (any-order differentiable: path-length regularisation).
        fused_modconv only applies to the unfused formulation: None = the reference's rule (grouped per-sample-weight conv in eval
        mode, shared-weight conv in training mode, networks.py:232).  Networks built with num_fp16_res / conv_clamp always run unfused."""
        assert t.ndim == 2 and len(ws) == len(t)
        assert ws.shape[1] == self.num_ws and ws.shape[2] == self.w_dim
        if motion_v is None:
            motion_v = self.motion_encoder(t, motion_z=motion_z, t_max=t_max)['motion_v']
        ws = ws.to(torch.float32).repeat_interleave(t.shape[1], dim=0)
        if unfused is None:
            unfused = (not ws.is_cuda) or self.mixed_precision
        assert unfused or not self.mixed_precision, 'fp16 / conv_clamp blocks run on the unfused ops only'
        if unfused:
            x = img = None
            w_idx = 0
            for res in self.block_resolutions:
                block = getattr(self, f'b{res}')
                x, img = block.forward_unfused(x, img, ws.narrow(1, w_idx, block.num_conv + block.num_torgb), motion_v=motion_v,
                                               fused_modconv=fused_modconv, noise_mode=noise_mode)
                w_idx += block.num_conv
            return img
        plans = self._plan_layers(ws)
        x = img = None
        for res in self.block_resolutions:
            block = getattr(self, f'b{res}')
            x, img = block(x, img, [plans[id(l)] for l in block.layers()], motion_v=motion_v, noise_mode=noise_mode)
        return img

    # evaluate the per-layer plans on a second CUDA stream (SGV_PARAM_STREAM=0 or False: same stream, for A/B measurements)
    param_stream = os.environ.get('SGV_PARAM_STREAM', '1') != '0'
    # issue the weight-gradient contractions on the parameter stream too (SGV_ASYNC_WGRAD=0: on the compute stream, inside the layer node)
    async_wgrad = os.environ.get('SGV_ASYNC_WGRAD', '1') != '0'

    def _plan_layers(self, ws):
        """Style affines, demodulation coefficients, ToRGB modulated weights and TF32 weight slabs of EVERY layer depend on
        (parameters, ws) only.  They are ~25 tiny launches per layer; evaluated here on a second stream they overlap the
        activation-sized kernels of the compute stream instead of sitting between them, and autograd replays their backward on
        that same stream.  Returns {id(layer): thunk}; the thunk makes the current stream wait for that layer's plan."""
        main = torch.cuda.current_stream(ws.device)
        use_side = self.param_stream and ws.is_cuda
        if use_side:
            if getattr(self, '_pstream', None) is None or self._pstream.device != ws.device:
                self._pstream = torch.cuda.Stream(ws.device)
            side = self._pstream
            side.wait_stream(main)
        out = {}
        with torch.cuda.stream(side) if use_side else contextlib.nullcontext():
            styles = self._all_styles(ws)
            for res in self.block_resolutions:
                for layer in getattr(self, f'b{res}').layers():
                    if isinstance(layer, SynthesisLayer):
                        plan = layer.plan(styles[id(layer)], async_wgrad_stream=side if (use_side and self.async_wgrad) else None)
                    else:
                        plan = layer.plan(styles[id(layer)])
                    ev = None
                    if use_side:
                        ev = torch.cuda.Event()
                        ev.record(side)
                    out[id(layer)] = _PlanThunk(plan, ev, main if use_side else None)
        return out


def _plan_tensors(obj):
    if isinstance(obj, torch.Tensor):
        yield obj
    elif isinstance(obj, dict):
        for v in obj.values():
            yield from _plan_tensors(v)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            yield from _plan_tensors(v)


class _PlanThunk:
    def __init__(self, plan, event, consumer_stream):
        self.plan, self.event, self.consumer = plan, event, consumer_stream

    def __call__(self):
        if self.event is not None:
            self.consumer.wait_event(self.event)
            for t in _plan_tensors(self.plan):
                t.record_stream(self.consumer)      # allocated on the parameter stream, consumed on the compute stream
        return self.plan
