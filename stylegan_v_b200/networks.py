"""Generator / Discriminator of StyleGAN-V around the native synthesis network — the callers either side of the hot path
(SURVEY.md §8 rows a14-a16 and §8f rank 4), state-dict compatible with the reference so its checkpoints load:

    Generator        mapping (MappingNetwork) + synthesis (stylegan_v_b200.synthesis.SynthesisNetwork)   networks.py:370-404
    Discriminator    b{256..8} DiscriminatorBlock ('resnet'), frame concat at `concat_res`, time-difference
                     conditioning (TemporalDifferenceEncoder -> MappingNetwork -> projection), b4 epilogue       networks.py:407-673

Every convolution layer is `Conv2dLayer` = conv2d_resample + bias_act on the drop-in ops (layers.py:141-197): on CUDA the FIR
and bias/activation kernels of libsgv_b200 and, through conv2d_gradfix, the tcgen05 implicit-GEMM kernels for every shape with
channel counts % 32 (the 3-channel fromrgb and the 513-channel epilogue conv use the library call like the reference).  These
ops are differentiable to any order, which the R1 penalty needs (loss.py:151-160).  fp16 blocks (`num_fp16_res`) are not built:
the fp32 contract of BASELINE.json applies.
"""
import numpy as np
import torch

from . import dconv
from . import dense as _dense
from .ops import bias_act, conv2d_resample, upfirdn2d
from .synthesis import SynthesisNetwork


def normalize_2nd_moment(x, dim=1, eps=1e-8):
    return x * (x.square().mean(dim=dim, keepdim=True) + eps).rsqrt()


class FullyConnectedLayer(torch.nn.Module):
    """Equalised-lr dense layer with optional activation (layers.py:108-138)."""

    def __init__(self, in_features, out_features, bias=True, activation='linear', lr_multiplier=1.0, bias_init=0.0):
        super().__init__()
        self.activation = activation
        self.weight = torch.nn.Parameter(torch.randn(out_features, in_features) / lr_multiplier)
        self.bias = torch.nn.Parameter(torch.full([out_features], float(bias_init))) if bias else None
        self.weight_gain = lr_multiplier / np.sqrt(in_features)
        self.bias_gain = lr_multiplier

    def forward(self, x, fused=False):
        """fused=True (first-order callers): ONE launch of the exact-fp32 dense kernel with the weight gain, bias, leaky ReLU and gain in its
        epilogue (stylegan_v_b200/dense.py -> csrc/dense_f32.cu); otherwise the reference's addmm / matmul + bias_act formulation,
        differentiable to any order (the R1 penalty differentiates the discriminator's dense layers twice)."""
        if fused and x.ndim == 2 and self.activation in ('linear', 'lrelu') and _dense.supported(x, self.weight, self.activation):
            return _dense.linear(x, self.weight, self.bias, self.weight_gain, self.bias_gain, act=self.activation,
                                 gain=bias_act.activation_funcs[self.activation].def_gain)
        w = self.weight.to(x.dtype) * self.weight_gain
        b = self.bias
        if b is not None:
            b = b.to(x.dtype)
            if self.bias_gain != 1:
                b = b * self.bias_gain
        if self.activation == 'linear' and b is not None:
            return torch.addmm(b.unsqueeze(0), x, w.t())
        return bias_act.bias_act(x.matmul(w.t()), b, act=self.activation)


class MappingNetwork(torch.nn.Module):
    """z (and / or an embedded condition c) -> w, broadcast to num_ws rows, with the running w_avg (layers.py:22-104)."""

    def __init__(self, z_dim, c_dim, w_dim, num_ws, num_layers=8, embed_features=None, layer_features=None, activation='lrelu',
                 lr_multiplier=0.01, w_avg_beta=0.995):
        super().__init__()
        self.z_dim, self.c_dim, self.w_dim, self.num_ws, self.num_layers, self.w_avg_beta = z_dim, c_dim, w_dim, num_ws, num_layers, w_avg_beta
        embed_features = 0 if c_dim == 0 else (w_dim if embed_features is None else embed_features)
        layer_features = w_dim if layer_features is None else layer_features
        widths = [z_dim + embed_features] + [layer_features] * (num_layers - 1) + [w_dim]
        if c_dim > 0:
            self.embed = FullyConnectedLayer(c_dim, embed_features)
        for i in range(num_layers):
            setattr(self, f'fc{i}', FullyConnectedLayer(widths[i], widths[i + 1], activation=activation, lr_multiplier=lr_multiplier))
        if num_ws is not None and w_avg_beta is not None:
            self.register_buffer('w_avg', torch.zeros([w_dim]))

    def forward(self, z, c, truncation_psi=1, truncation_cutoff=None, skip_w_avg_update=False):
        parts = []
        if self.z_dim > 0:
            assert z.shape[1] == self.z_dim
            parts.append(normalize_2nd_moment(z.to(torch.float32)))
        if self.c_dim > 0:
            assert c.shape[1] == self.c_dim
            parts.append(normalize_2nd_moment(self.embed(c.to(torch.float32))))
        x = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
        for i in range(self.num_layers):
            # the mapping network is only ever differentiated once (path length and R1 differentiate the synthesis / discriminator bodies twice,
            # and reach the mapping through a plain first-order backward): its layers always take the kernel route where the shape allows
            x = getattr(self, f'fc{i}')(x, fused=True)
        if self.w_avg_beta is not None and self.training and not skip_w_avg_update:
            self.w_avg.copy_(x.detach().mean(dim=0).lerp(self.w_avg, self.w_avg_beta))
        if self.num_ws is not None:
            x = x.unsqueeze(1).repeat([1, self.num_ws, 1])
        if truncation_psi != 1:
            assert self.w_avg_beta is not None
            if self.num_ws is None or truncation_cutoff is None:
                x = self.w_avg.lerp(x, truncation_psi)
            else:
                x[:, :truncation_cutoff] = self.w_avg.lerp(x[:, :truncation_cutoff], truncation_psi)
        return x


class Conv2dLayer(torch.nn.Module):
    """Equalised-lr convolution with optional 2x resampling, bias, activation, gain, clamp (layers.py:141-197)."""

    def __init__(self, in_channels, out_channels, kernel_size, bias=True, activation='linear', up=1, down=1, resample_filter=(1, 3, 3, 1),
                 conv_clamp=None, trainable=True):
        super().__init__()
        self.activation, self.up, self.down, self.conv_clamp = activation, up, down, conv_clamp
        self.register_buffer('resample_filter', upfirdn2d.setup_filter(list(resample_filter)))
        self.padding = kernel_size // 2
        self.weight_gain = 1 / np.sqrt(in_channels * kernel_size ** 2)
        self.act_gain = bias_act.activation_funcs[activation].def_gain
        weight = torch.randn(out_channels, in_channels, kernel_size, kernel_size)
        b = torch.zeros(out_channels) if bias else None
        if trainable:
            self.weight = torch.nn.Parameter(weight)
            self.bias = torch.nn.Parameter(b) if b is not None else None
        else:                                                      # Freeze-D (networks.py:441-447)
            self.register_buffer('weight', weight)
            if b is not None:
                self.register_buffer('bias', b)
            else:
                self.bias = None

    def forward(self, x, gain=1, fused=False):
        """fused=True: layers inside the native envelope run as [FIR] + ONE implicit-GEMM launch with the equalised-lr weight gain folded into
        the weight pass and the bias / activation epilogue (stylegan_v_b200/dconv.py; first-order differentiable); a 1x1 layer on <= 4 input
        channels (fromrgb) runs as one streaming pass (dconv.fromrgb); everything else — and fused=False — is conv2d_resample + bias_act."""
        b = self.bias.to(x.dtype) if self.bias is not None else None
        if fused and self.up == 1 and self.down in (1, 2) and self.conv_clamp is None and self.activation in ('linear', 'lrelu'):
            k = self.weight.shape[2]
            if k == 1 and self.down == 1 and dconv.fromrgb_supported(x, self.weight):
                return dconv.fromrgb(x, self.weight, b, act=self.activation, gain=self.act_gain * gain, weight_gain=self.weight_gain)
            # down layers (conv2d_resample.py:100-110,119-122): k = 3 -> low-pass at full resolution, then the convolution strides;
            #                                                    k = 1 -> the FIR decimates, then a 1x1 convolution
            stride, pad = (2, 0) if (self.down == 2 and k == 3) else (1, self.padding if self.down == 1 else 0)
            if dconv.supported(x, self.weight, stride, pad):
                if self.down == 2:
                    fw = self.resample_filter.shape[-1]
                    p0, p1 = self.padding + (fw - self.down + 1) // 2, self.padding + (fw - self.down) // 2       # conv2d_resample.py:100-104
                    x = upfirdn2d.upfirdn2d(x, self.resample_filter, down=(1 if k == 3 else 2), padding=[p0, p1, p0, p1])
                return dconv.fused_conv_act(x, self.weight, b, stride=stride, padding=pad, act=self.activation, gain=self.act_gain * gain,
                                            weight_gain=self.weight_gain)
        w = self.weight * self.weight_gain
        x = conv2d_resample.conv2d_resample(x=x, w=w.to(x.dtype), f=self.resample_filter, up=self.up, down=self.down, padding=self.padding,
                                            flip_weight=(self.up == 1))
        clamp = self.conv_clamp * gain if self.conv_clamp is not None else None
        return bias_act.bias_act(x, b, act=self.activation, gain=self.act_gain * gain, clamp=clamp)


class DiscriminatorBlock(torch.nn.Module):
    """One resolution of the discriminator (networks.py:407-488): fromrgb (first block / 'skip'), conv0, conv1 (down 2) and, for
    'resnet', the 1x1 down-2 skip branch; both branches scaled by sqrt(1/2)."""

    def __init__(self, in_channels, tmp_channels, out_channels, resolution, img_channels, first_layer_idx, architecture='resnet',
                 activation='lrelu', resample_filter=(1, 3, 3, 1), conv_clamp=None, freeze_layers=0, use_fp16=False):
        assert architecture in ('orig', 'skip', 'resnet')
        super().__init__()
        self.in_channels, self.resolution, self.img_channels, self.architecture = in_channels, resolution, img_channels, architecture
        self.use_fp16 = use_fp16                                   # mixed-precision mode of the reference (networks.py:462); unfused ops only
        self.register_buffer('resample_filter', upfirdn2d.setup_filter(list(resample_filter)))
        self.num_layers = 0

        def trainable():
            flag = (first_layer_idx + self.num_layers) >= freeze_layers
            self.num_layers += 1
            return flag
        conv0_in = in_channels if in_channels > 0 else tmp_channels
        if in_channels == 0 or architecture == 'skip':
            self.fromrgb = Conv2dLayer(img_channels, tmp_channels, 1, activation=activation, trainable=trainable(), conv_clamp=conv_clamp)
        self.conv0 = Conv2dLayer(conv0_in, tmp_channels, 3, activation=activation, trainable=trainable(), conv_clamp=conv_clamp)
        self.conv1 = Conv2dLayer(tmp_channels, out_channels, 3, activation=activation, down=2, trainable=trainable(),
                                 resample_filter=resample_filter, conv_clamp=conv_clamp)
        if architecture == 'resnet':
            self.skip = Conv2dLayer(conv0_in, out_channels, 1, bias=False, down=2, trainable=trainable(), resample_filter=resample_filter)

    def forward(self, x, img, fused=False):
        dtype = torch.float16 if self.use_fp16 else torch.float32
        fused = fused and not self.use_fp16
        if x is not None:
            assert x.shape[1] == self.in_channels and x.shape[2] == x.shape[3] == self.resolution
            x = x.to(dtype)
        if self.in_channels == 0 or self.architecture == 'skip':
            assert img.shape[1] == self.img_channels and img.shape[2] == img.shape[3] == self.resolution
            img = img.to(dtype)
            if fused and img.is_cuda and not dconv.fromrgb_supported(img, self.fromrgb.weight):
                # library route of the 1x1 layer: a 3-channel NHWC view of the frames makes it write its output channels_last, the layout of the
                # fused conv0 that follows (the streaming fromrgb kernel does that by itself from the NCHW frames)
                img = img.contiguous(memory_format=torch.channels_last)
            y = self.fromrgb(img, fused=fused)
            x = x + y if x is not None else y
            img = upfirdn2d.downsample2d(img, self.resample_filter) if self.architecture == 'skip' else None
        if self.architecture == 'resnet':
            # (the add stays a separate in-place pass into the LINEAR skip output: conv1's own output must survive unchanged, its backward
            #  recovers the leaky-ReLU slope from it — an accumulate-epilogue into either branch's buffer was tried and dropped for that reason)
            y = self.skip(x, gain=np.sqrt(0.5), fused=fused)
            x = self.conv1(self.conv0(x, fused=fused), gain=np.sqrt(0.5), fused=fused)
            x = y.add_(x)
        else:
            x = self.conv1(self.conv0(x, fused=fused), fused=fused)
        return x, img


class MinibatchStdLayer(torch.nn.Module):
    """Appends the per-group standard deviation, averaged over channels and pixels, as extra feature maps (networks.py:492-516)."""

    def __init__(self, group_size, num_channels=1):
        super().__init__()
        self.group_size, self.num_channels = group_size, num_channels

    def forward(self, x):
        N, C, H, W = x.shape
        G = min(self.group_size, N) if self.group_size is not None else N
        F = self.num_channels
        y = x.reshape(G, -1, F, C // F, H, W)
        y = y - y.mean(dim=0)
        y = (y.square().mean(dim=0) + 1e-8).sqrt()
        y = y.mean(dim=[2, 3, 4]).reshape(-1, F, 1, 1).repeat(G, 1, H, W)
        return torch.cat([x, y], dim=1)


class DiscriminatorEpilogue(torch.nn.Module):
    """4x4 tail: minibatch-std, 3x3 conv, two dense layers, projection onto the mapped condition (networks.py:520-576)."""

    def __init__(self, in_channels, cmap_dim, resolution, img_channels, architecture='resnet', mbstd_group_size=4, mbstd_num_channels=1,
                 activation='lrelu', conv_clamp=None):
        assert architecture in ('orig', 'skip', 'resnet')
        super().__init__()
        self.in_channels, self.cmap_dim, self.resolution, self.img_channels, self.architecture = in_channels, cmap_dim, resolution, img_channels, architecture
        if architecture == 'skip':
            self.fromrgb = Conv2dLayer(img_channels, in_channels, 1, activation=activation)
        self.mbstd = MinibatchStdLayer(mbstd_group_size, mbstd_num_channels) if mbstd_num_channels > 0 else None
        self.conv = Conv2dLayer(in_channels + mbstd_num_channels, in_channels, 3, activation=activation, conv_clamp=conv_clamp)
        self.fc = FullyConnectedLayer(in_channels * resolution ** 2, in_channels, activation=activation)
        self.out = FullyConnectedLayer(in_channels, 1 if cmap_dim == 0 else cmap_dim)

    def forward(self, x, img, cmap, fused=False):
        assert x.shape[1] == self.in_channels and x.shape[2] == x.shape[3] == self.resolution
        w = self.conv.weight
        if (fused and x.is_cuda and x.dtype == torch.float32 and self.architecture != 'skip' and self.mbstd is not None and self.conv.conv_clamp is None
                and x.shape[0] % min(self.mbstd.group_size or x.shape[0], x.shape[0]) == 0 and w.shape[0] % 64 == 0 and x.shape[1] % 4 == 0):
            # minibatch-std + concat + zero padding of the 513 channels to 576 in ONE kernel (NHWC), then the 3x3 conv on the tcgen05 kernel with
            # zero input-channel padding of its weight (instead of the library call a 513-channel contraction needs)
            # (padding to a multiple of 64: the padded count is GEMM-N of the data-gradient contraction)
            xp = dconv.minibatch_std_concat(x, self.mbstd.group_size, self.mbstd.num_channels, pad_to=64)
            wpad = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, xp.shape[1] - w.shape[1]))
            assert dconv.supported(xp, wpad, 1, 1)
            x = dconv.fused_conv_act(xp, wpad, self.conv.bias, stride=1, padding=1, act=self.conv.activation, gain=self.conv.act_gain,
                                     weight_gain=self.conv.weight_gain)
        else:
            x = x.to(dtype=torch.float32, memory_format=torch.contiguous_format)
            if self.architecture == 'skip':
                x = x + self.fromrgb(img.to(dtype=torch.float32, memory_format=torch.contiguous_format))
            if self.mbstd is not None:
                x = self.mbstd(x)
            x = self.conv(x)
        # flatten in (C, H, W) order like the reference, whatever memory format the conv returned
        x = self.out(self.fc(x.contiguous().flatten(1), fused=fused), fused=fused)
        if self.cmap_dim > 0:
            assert cmap.shape[1] == self.cmap_dim
            x = (x * cmap).sum(dim=1, keepdim=True) * (1 / np.sqrt(self.cmap_dim))
        return x


def log_spaced_frequencies(max_num_frames, skip_small_t_freqs=0):
    """pi * 2^k / T for k < log2 T, T = max_num_frames rounded up to a power of two (layers.py:439-446)."""
    resolution = 2 ** int(np.ceil(np.log2(max_num_frames)))
    count = int(np.ceil(np.log2(resolution)))
    powers = (2 ** torch.arange(count))[:count - skip_small_t_freqs]
    return powers.unsqueeze(0).float() * np.pi / resolution


class FixedTimeEncoder(torch.nn.Module):
    """[sin, cos](coefs * t) with fixed log-spaced coefficients (layers.py:300-327)."""

    def __init__(self, max_num_frames, skip_small_t_freqs=0):
        super().__init__()
        assert max_num_frames >= 1
        self.register_buffer('fourier_coefs', log_spaced_frequencies(max_num_frames, skip_small_t_freqs))

    def get_dim(self):
        return self.fourier_coefs.shape[1] * 2

    def forward(self, t):
        assert t.ndim == 2
        raw = self.fourier_coefs * t.reshape(-1).float().unsqueeze(1)
        return torch.cat([raw.sin(), raw.cos()], dim=1)


class TemporalDifferenceEncoder(torch.nn.Module):
    """Condition of the discriminator: learned + Fourier embedding of the frame distances of a clip (layers.py:255-296)."""

    def __init__(self, num_frames_per_video, max_num_frames, sampling_type='random', skip_small_t_freqs=0):
        super().__init__()
        self.num_frames_per_video, self.sampling_type = num_frames_per_video, sampling_type
        if num_frames_per_video > 1:
            self.d = 256
            self.const_embed = torch.nn.Embedding(max_num_frames, self.d)
            self.time_encoder = FixedTimeEncoder(max_num_frames, skip_small_t_freqs=skip_small_t_freqs)

    def get_dim(self):
        if self.num_frames_per_video == 1:
            return 1
        per_diff = self.d + self.time_encoder.get_dim()
        return per_diff if self.sampling_type == 'uniform' else per_diff * (self.num_frames_per_video - 1)

    def forward(self, t):
        assert t.ndim == 2 and t.shape[1] == self.num_frames_per_video
        B = t.shape[0]
        if self.num_frames_per_video == 1:
            return torch.zeros(B, 1, device=t.device)
        diffs = (t[:, 1] - t[:, 0]) if self.sampling_type == 'uniform' else (t[:, 1:] - t[:, :-1]).reshape(-1)
        emb = torch.cat([self.const_embed(diffs.float().round().long()), self.time_encoder(diffs.unsqueeze(1))], dim=1)
        return emb.reshape(B, -1)


def _concat_frames(x, num_frames, keep_nhwc=False):
    """[B*F, C, h, w] -> [B, F*C, h, w] with channel index f * C + c (networks.py:659-662).  For an NHWC activation on the fused path the
    result is produced NHWC with ONE copy (pixel-major, the frames' channel vectors side by side) instead of NCHW-and-back."""
    BF, Cc, h, w = x.shape
    if keep_nhwc and x.is_cuda and x.stride(1) == 1 and Cc > 1:
        m = x.permute(0, 2, 3, 1).reshape(BF // num_frames, num_frames, h, w, Cc).permute(0, 2, 3, 1, 4).reshape(BF // num_frames, h, w, num_frames * Cc)
        return m.permute(0, 3, 1, 2)
    return x.contiguous().reshape(-1, num_frames * Cc, h, w)


class Discriminator(torch.nn.Module):
    """StyleGAN-V discriminator (networks.py:580-673).  Frames of a clip are processed independently down to `concat_res`, where the
    channel dimension of the F frames is concatenated; logits are conditioned on the frame time differences."""

    def __init__(self, c_dim=0, img_resolution=256, img_channels=3, architecture='resnet', channel_base=16384, channel_max=512,
                 conv_clamp=None, cmap_dim=None, num_frames_per_video=3, max_num_frames=1024, sampling_type='random', concat_res=16,
                 num_frames_div_factor=2, dummy_c=False, mbstd_group_size=4, mbstd_num_channels=1, mapping_layers=2, freeze_layers=0,
                 resample_filter=(1, 3, 3, 1), num_fp16_res=0):
        super().__init__()
        self.c_dim, self.img_resolution, self.img_channels = c_dim, img_resolution, img_channels
        self.num_frames_per_video, self.concat_res, self.dummy_c = num_frames_per_video, concat_res, dummy_c
        log2 = int(np.log2(img_resolution))
        self.block_resolutions = [2 ** i for i in range(log2, 2, -1)]
        ch = {res: min(channel_base // res, channel_max) for res in self.block_resolutions + [4]}
        fp16_resolution = max(2 ** (log2 + 1 - num_fp16_res), 8)                                       # networks.py:611
        if cmap_dim is None:
            cmap_dim = ch[4]
        self.time_encoder = TemporalDifferenceEncoder(num_frames_per_video, max_num_frames, sampling_type) if num_frames_per_video > 1 else None
        if c_dim == 0 and self.time_encoder is None:
            cmap_dim = 0
        total_c_dim = c_dim + (0 if self.time_encoder is None else self.time_encoder.get_dim())
        layer_idx = 0
        for res in self.block_resolutions:
            cin = ch[res] if res < img_resolution else 0
            cout = ch[res // 2]
            if res // 2 == concat_res:
                cout = cout // num_frames_div_factor
            if res == concat_res:
                cin = (cin // num_frames_div_factor) * num_frames_per_video
            block = DiscriminatorBlock(cin, ch[res], cout, res, img_channels, layer_idx, architecture=architecture, conv_clamp=conv_clamp,
                                       freeze_layers=freeze_layers, resample_filter=resample_filter, use_fp16=(res >= fp16_resolution))
            setattr(self, f'b{res}', block)
            layer_idx += block.num_layers
        if c_dim > 0 or self.time_encoder is not None:
            self.mapping = MappingNetwork(z_dim=0, c_dim=total_c_dim, w_dim=cmap_dim, num_ws=None, w_avg_beta=None, num_layers=mapping_layers)
        self.b4 = DiscriminatorEpilogue(ch[4], cmap_dim, 4, img_channels, architecture=architecture, mbstd_group_size=mbstd_group_size,
                                        mbstd_num_channels=mbstd_num_channels, conv_clamp=conv_clamp)

    @classmethod
    def from_reference_cfg(cls, cfg, img_resolution, img_channels=3, c_dim=0, channel_base=16384, channel_max=512, mbstd_group_size=4,
                           mapping_layers=2, **kwargs):
        """Builds the module from the reference's `cfg.model.discriminator` node (configs/model/stylegan-v.yaml:47-51 + `sampling`), given as a
        nested dict / attribute object, the way train.py:166-178 assembles D_kwargs."""
        get = (lambda o, k, d=None: o.get(k, d) if hasattr(o, 'get') else getattr(o, k, d))
        sampling = get(cfg, 'sampling')
        return cls(c_dim=c_dim, img_resolution=img_resolution, img_channels=img_channels, channel_base=channel_base, channel_max=channel_max,
                   num_frames_per_video=get(sampling, 'num_frames_per_video'), max_num_frames=get(sampling, 'max_num_frames'),
                   sampling_type=get(sampling, 'type', 'random'), concat_res=get(cfg, 'concat_res', 16),
                   num_frames_div_factor=get(cfg, 'num_frames_div_factor', 2), dummy_c=get(cfg, 'dummy_c', False),
                   mbstd_group_size=mbstd_group_size, mapping_layers=mapping_layers, **kwargs)

    def forward(self, img, c, t, fused=None):
        """fused=None: CUDA inputs use the fused conv + bias + activation nodes when no second-order gradient can be requested, i.e.
        when gradient mode is off or the caller says so explicitly; TrainingPhases passes fused=True for the main phases and False for R1."""
        assert t.ndim == 2 and len(img) == t.shape[0] * t.shape[1]
        fused = bool(fused) if fused is not None else (img.is_cuda and not torch.is_grad_enabled())
        if self.time_encoder is not None:
            c = torch.cat([c, self.time_encoder(t.reshape(-1, self.num_frames_per_video))], dim=1)
            if self.dummy_c:
                c = c * 0.0
        x = None
        for res in self.block_resolutions:
            if res == self.concat_res:
                x = _concat_frames(x, self.num_frames_per_video, keep_nhwc=fused)     # [B, F*C, h, w] in (frame, channel) order
            x, img = getattr(self, f'b{res}')(x, img, fused=fused)
        cmap = self.mapping(None, c) if c.shape[1] > 0 else None
        return {'image_logits': self.b4(x, img, cmap, fused=fused).squeeze(1)}


class Generator(torch.nn.Module):
    """z, c, t -> frames: mapping network + native synthesis network (networks.py:370-404)."""

    def __init__(self, z_dim=512, c_dim=0, w_dim=512, img_resolution=256, img_channels=3, mapping_layers=2, **synthesis_kwargs):
        super().__init__()
        self.z_dim, self.c_dim, self.w_dim, self.img_resolution, self.img_channels = z_dim, c_dim, w_dim, img_resolution, img_channels
        self.synthesis = SynthesisNetwork(w_dim=w_dim, img_resolution=img_resolution, img_channels=img_channels, **synthesis_kwargs)
        self.num_ws = self.synthesis.num_ws
        self.mapping = MappingNetwork(z_dim=z_dim, c_dim=c_dim, w_dim=w_dim, num_ws=self.num_ws, num_layers=mapping_layers)

    @classmethod
    def from_reference_cfg(cls, cfg, img_resolution, img_channels=3, channel_base=16384, channel_max=512, mapping_layers=2, **kwargs):
        """Builds the module from the reference's `cfg.model.generator` node (configs/model/stylegan-v.yaml:3-45 + `sampling`), nested dict or
        attribute object, the way train.py:163-172 assembles G_kwargs.  Options this implementation does not cover raise instead of being ignored."""
        get = (lambda o, k, d=None: o.get(k, d) if hasattr(o, 'get') else getattr(o, k, d))
        motion, tenc, sampling = get(cfg, 'motion'), get(cfg, 'time_enc'), get(cfg, 'sampling')
        unsupported = []
        if get(get(cfg, 'input', {}), 'type', 'temporal') != 'temporal':
            unsupported.append('input.type != temporal')
        if get(tenc, 'cond_type', 'concat_const') != 'concat_const':
            unsupported.append('time_enc.cond_type != concat_const')
        if get(motion, 'gen_strategy', 'conv') != 'conv' or not get(motion, 'fourier', True):
            unsupported.append('motion.gen_strategy != conv / fourier=false')
        if get(cfg, 'c_dim', 0) not in (0, None):
            unsupported.append('c_dim > 0')
        if unsupported:
            raise NotImplementedError('generator options outside the stylegan-v.yaml model: ' + ', '.join(unsupported))
        return cls(z_dim=get(cfg, 'z_dim', 512), c_dim=0, w_dim=get(cfg, 'w_dim', 512), img_resolution=img_resolution, img_channels=img_channels,
                   mapping_layers=mapping_layers, channel_base=channel_base, channel_max=channel_max,
                   motion_z_dim=get(motion, 'z_dim', 512), motion_v_dim=get(motion, 'v_dim', 512), motion_kernel_size=get(motion, 'kernel_size', 11),
                   motion_z_distance=get(motion, 'motion_z_distance', get(tenc, 'min_period_len', 16)), time_enc_dim=get(tenc, 'dim', 256),
                   min_period_len=get(tenc, 'min_period_len', 16), max_period_len=get(tenc, 'max_period_len', 1024),
                   max_num_frames=get(sampling, 'max_num_frames', 1024), use_noise=bool(get(cfg, 'use_noise', False)), **kwargs)

    def forward(self, z, c, t, truncation_psi=1, truncation_cutoff=None, **synthesis_kwargs):
        assert len(z) == len(c) == len(t) and t.ndim == 2
        ws = self.mapping(z, c, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff)
        return self.synthesis(ws, t=t, c=c, **synthesis_kwargs)
