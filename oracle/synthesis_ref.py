"""ORACLE (test infrastructure only — see oracle/__init__.py).

Functional CPU restatement (torch ops, differentiable through autograd) of the reference's generator
synthesis path and Fourier time-encoder.  Parameters live in a flat dict keyed by the reference's
state_dict names (so a reference `SynthesisNetwork.state_dict()` can be passed in unchanged).

  fully_connected          src/training/layers.py:125-138   (FullyConnectedLayer.forward)
  eqlr_conv1d              src/training/layers.py:356-373   (EqLRConv1d.forward)
  motion_encoder           src/training/motion.py:63-127 (trajectory, gather, lerp) + 132-156
  aligned_time_encoder     src/training/motion.py:185-214, frequencies 218-222, phase scales 176-178
  synthesis_layer          src/training/networks.py:124-144
  torgb_layer              src/training/networks.py:159-163
  synthesis_forward        src/training/networks.py:224-266 (block, 'skip' arch) + 324-366 (network, concat_const)
  temporal_input           src/training/layers.py:242-251
"""
from dataclasses import dataclass, field
import numpy as np
import torch
import torch.nn.functional as F

from . import ops_ref


@dataclass
class SynthesisConfig:
    """Shape-defining subset of configs/model/stylegan-v.yaml + train.py:158-174."""
    img_resolution: int = 256
    img_channels: int = 3
    w_dim: int = 512
    channel_base: int = 16384       # fmaps 0.5 * 32768 for res < 512 (train.py:158,167)
    channel_max: int = 512
    motion_z_dim: int = 512
    motion_v_dim: int = 512
    motion_kernel_size: int = 11
    motion_z_distance: int = 16     # = time_enc.min_period_len (stylegan-v.yaml:16)
    time_enc_dim: int = 256
    min_period_len: int = 16
    max_period_len: int = 1024
    max_num_frames: int = 1024
    resample_filter: tuple = (1, 3, 3, 1)
    use_noise: bool = False         # stylegan-v.yaml:6
    conv_clamp: float = None        # fp32 contract (num_fp16_res = 0)

    @property
    def block_resolutions(self):
        return [2 ** i for i in range(2, int(np.log2(self.img_resolution)) + 1)]

    def channels(self, res):
        return min(self.channel_base // res, self.channel_max)

    @property
    def motion_out_dim(self):       # AlignedTimeEncoder.get_dim (motion.py:182-183)
        return self.time_enc_dim * 2

    @property
    def num_ws(self):               # networks.py:300-321
        n = 0
        for res in self.block_resolutions:
            n += 1 if res == 4 else 2
        return n + 1

    def reference_generator_cfg(self):
        """EasyDict-like nested dict equal to what the reference networks read from cfg (for make_goldens)."""
        return dict(
            sampling=dict(max_num_frames=self.max_num_frames, num_frames_per_video=3, type='random',
                          total_dists=[1, 2, 4, 8, 16, 32], max_dist=32),
            use_noise=self.use_noise, input=dict(type='temporal'), w_dim=self.w_dim, z_dim=self.w_dim, c_dim=0,
            motion=dict(z_dim=self.motion_z_dim, v_dim=self.motion_v_dim, motion_z_distance=self.motion_z_distance,
                        gen_strategy='conv', kernel_size=self.motion_kernel_size, use_fractional_t=True, fourier=True),
            time_enc=dict(cond_type='concat_const', dim=self.time_enc_dim, min_period_len=self.min_period_len,
                          max_period_len=self.max_period_len, phase_dropout_std=1.0),
        )


# ----------------------------------------------------------------------------------------------
# small layers

def fully_connected(x, weight, bias=None, lr_multiplier=1.0, activation='linear'):
    w = weight * (lr_multiplier / np.sqrt(weight.shape[1]))
    b = bias
    if b is not None and lr_multiplier != 1:
        b = b * lr_multiplier
    if activation == 'linear' and b is not None:
        return torch.addmm(b.unsqueeze(0), x, w.t())
    x = x.matmul(w.t())
    return ops_ref.bias_act_ref_torch(x, b, act=activation)


def eqlr_conv1d(x, weight, bias, lr_multiplier=1.0, activation='lrelu'):
    w = weight * (lr_multiplier / np.sqrt(weight.shape[1] * weight.shape[2]))
    b = bias * lr_multiplier if lr_multiplier != 1 else bias
    y = F.conv1d(x, w, b)
    return F.leaky_relu(y, 0.2) if activation == 'lrelu' else y


def linspaced_frequencies(num_freqs, min_period_len, max_period_len):
    freqs = 2 * np.pi / (2 ** np.linspace(np.log2(min_period_len), np.log2(max_period_len), num_freqs))
    return torch.from_numpy(freqs[::-1].copy().astype(np.float32)).unsqueeze(0)


# ----------------------------------------------------------------------------------------------
# motion encoder (Fourier time-encoder)

def max_traj_len(cfg: SynthesisConfig, t_max: float) -> int:
    max_t = max(cfg.max_num_frames - 1, t_max)                                       # motion.py:64
    return int(np.ceil(max_t / cfg.motion_z_distance)) + 2 + (cfg.motion_kernel_size - 1) * 2   # :65,59,80


def motion_encoder(P, cfg: SynthesisConfig, t, motion_z, prefix='motion_encoder.'):
    """t [B,F] float, motion_z [B, >=max_traj_len, z_dim] -> motion_v [B*F, 2*time_enc_dim]."""
    B, Fr = t.shape
    L = max_traj_len(cfg, float(t.max()))
    traj_in = motion_z[:B, :L, :cfg.motion_z_dim]
    h = traj_in.permute(0, 2, 1)
    h = eqlr_conv1d(h, P[prefix + 'conv.0.weight'], P[prefix + 'conv.0.bias'], lr_multiplier=0.01)
    h = eqlr_conv1d(h, P[prefix + 'conv.1.weight'], P[prefix + 'conv.1.bias'], lr_multiplier=0.01)
    trajs = h.permute(0, 2, 1)                                                       # [B, L-20, v_dim]

    d = cfg.motion_z_distance
    left_idx = (t / d).floor().long()                                                # motion.py:105
    bidx = torch.arange(B).unsqueeze(1).repeat(1, Fr)
    u_left = trajs[bidx, left_idx]
    u_right = trajs[bidx, left_idx + 1]
    t_left = t - t % d
    t_right = t_left + d
    alpha = ((t % d) / d).unsqueeze(2).to(torch.float32)

    # AlignedTimeEncoder.forward (motion.py:185-214)
    te = prefix + 'time_encoder.'
    freqs = linspaced_frequencies(cfg.time_enc_dim, cfg.min_period_len, cfg.max_period_len).to(t.device)
    phase_scales = cfg.max_period_len / (2 * np.pi / freqs)
    uL = u_left.reshape(B * Fr, -1)
    uR = u_right.reshape(B * Fr, -1)
    periods = fully_connected(uL, P[te + 'periods_predictor.weight']).tanh() + 1
    phases = fully_connected(uL, P[te + 'phase_predictor.weight'])
    al_left = fully_connected(uL, P[te + 'aligners_predictor.weight'])
    al_right = fully_connected(uR, P[te + 'aligners_predictor.weight'])

    def emb(tt):
        raw = freqs * periods * tt.reshape(-1).float().unsqueeze(1) + phases * phase_scales
        return torch.cat([raw.sin(), raw.cos()], dim=1)
    a = alpha.reshape(-1, 1)
    remove = emb(t_left) * (1 - a) + emb(t_right) * a
    add = al_left * (1 - a) + al_right * a
    return emb(t) - remove + add


# ----------------------------------------------------------------------------------------------
# synthesis layers

def synthesis_layer(P, name, x, w, up, resample_filter, fused_modconv, gain=1.0, conv_clamp=None, noise_mode='none'):
    styles = fully_connected(w, P[name + '.affine.weight'], P[name + '.affine.bias'])
    weight = P[name + '.weight']
    noise = None                                                                     # networks.py:130-134 ('random' draws are the caller's business)
    if noise_mode == 'const' and (name + '.noise_const') in P:
        noise = P[name + '.noise_const'] * P[name + '.noise_strength']
    x = ops_ref.modulated_conv2d_ref(x, weight, styles, noise=noise, up=up, padding=weight.shape[2] // 2,
                                     resample_filter=resample_filter, flip_weight=(up == 1), fused_modconv=fused_modconv)
    act_gain = float(np.sqrt(2)) * gain
    act_clamp = conv_clamp * gain if conv_clamp is not None else None
    return ops_ref.bias_act_ref_torch(x, P[name + '.bias'], act='lrelu', gain=act_gain, clamp=act_clamp)


def torgb_layer(P, name, x, w, fused_modconv, conv_clamp=None):
    weight = P[name + '.weight']
    styles = fully_connected(w, P[name + '.affine.weight'], P[name + '.affine.bias']) * (1 / np.sqrt(weight.shape[1]))
    x = ops_ref.modulated_conv2d_ref(x, weight, styles, demodulate=False, fused_modconv=fused_modconv)
    return ops_ref.bias_act_ref_torch(x, P[name + '.bias'], clamp=conv_clamp)


def synthesis_forward(P, cfg: SynthesisConfig, ws, t, motion_z=None, motion_v=None, fused_modconv=False,
                      return_features=False, noise_mode='none'):
    """ws [B, num_ws, w_dim], t [B, F] -> img [B*F, 3, R, R] (networks.py:324-366, cond_type concat_const)."""
    B, Fr = t.shape
    assert ws.shape[1] == cfg.num_ws
    if motion_v is None:
        motion_v = motion_encoder(P, cfg, t, motion_z)
    ws = ws.repeat_interleave(Fr, dim=0).to(torch.float32)
    f = ops_ref.setup_filter(list(cfg.resample_filter))
    x = img = None
    w_idx = 0
    feats = {}
    for res in cfg.block_resolutions:
        b = f'b{res}'
        num_conv = 1 if res == 4 else 2
        cur = ws.narrow(1, w_idx, num_conv + 1)
        w_idx += num_conv
        wi = iter(cur.unbind(dim=1))
        if res == 4:
            const = P[b + '.input.input.const']                                         # layers.py:246-249
            x = torch.cat([const.repeat(B * Fr, 1, 1, 1),
                           motion_v.unsqueeze(2).unsqueeze(3).repeat(1, 1, 4, 4)], dim=1)
            x = synthesis_layer(P, b + '.conv1', x, next(wi), 1, f, fused_modconv, conv_clamp=cfg.conv_clamp, noise_mode=noise_mode)
        else:
            x = synthesis_layer(P, b + '.conv0', x, next(wi), 2, f, fused_modconv, conv_clamp=cfg.conv_clamp, noise_mode=noise_mode)
            x = synthesis_layer(P, b + '.conv1', x, next(wi), 1, f, fused_modconv, conv_clamp=cfg.conv_clamp, noise_mode=noise_mode)
        if img is not None:
            img = ops_ref.upfirdn2d_ref_torch(img, f, up=2, padding=[2, 1, 2, 1], gain=4)   # upsample2d (upfirdn2d.py:308-343)
        y = torgb_layer(P, b + '.torgb', x, next(wi), fused_modconv, conv_clamp=cfg.conv_clamp)
        img = img + y if img is not None else y
        feats[b] = x
    return (img, feats) if return_features else img


# ----------------------------------------------------------------------------------------------
# parameters

def init_params(cfg: SynthesisConfig, seed=0, dtype=torch.float32):
    """Reference initialisation (networks.py:116-122,152-157; layers.py:119-123,345-347,238)."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=dtype)
    P = {}
    me = 'motion_encoder.'
    k = cfg.motion_kernel_size
    P[me + 'conv.0.weight'] = rn(cfg.motion_z_dim, cfg.motion_z_dim, k) / 0.01
    P[me + 'conv.0.bias'] = torch.zeros(cfg.motion_z_dim, dtype=dtype)
    P[me + 'conv.1.weight'] = rn(cfg.motion_v_dim, cfg.motion_z_dim, k) / 0.01
    P[me + 'conv.1.bias'] = torch.zeros(cfg.motion_v_dim, dtype=dtype)
    P[me + 'time_encoder.periods_predictor.weight'] = rn(cfg.time_enc_dim, cfg.motion_v_dim)
    P[me + 'time_encoder.phase_predictor.weight'] = rn(cfg.time_enc_dim, cfg.motion_v_dim)
    P[me + 'time_encoder.aligners_predictor.weight'] = rn(cfg.time_enc_dim * 2, cfg.motion_v_dim)
    for res in cfg.block_resolutions:
        b = f'b{res}'
        out_c = cfg.channels(res)
        layers = []
        if res == 4:
            P[b + '.input.input.const'] = rn(1, out_c, 4, 4)
            layers.append((b + '.conv1', out_c + cfg.motion_out_dim, out_c, 3))
        else:
            layers.append((b + '.conv0', cfg.channels(res // 2), out_c, 3))
            layers.append((b + '.conv1', out_c, out_c, 3))
        layers.append((b + '.torgb', out_c, cfg.img_channels, 1))
        for name, ic, oc, ks in layers:
            P[name + '.affine.weight'] = rn(ic, cfg.w_dim)
            P[name + '.affine.bias'] = torch.ones(ic, dtype=dtype)
            P[name + '.weight'] = rn(oc, ic, ks, ks)
            P[name + '.bias'] = torch.zeros(oc, dtype=dtype)
    return P


def conv_flops_per_frame(cfg: SynthesisConfig) -> float:
    """Algorithmic conv FLOPs per frame: 2*Cin*Cout*k^2*H_out*W_out for stride-1,
    2*Cin*Cout*k^2*H_in*W_in for the stride-2 transposed conv."""
    total = 0.0
    for res in cfg.block_resolutions:
        oc = cfg.channels(res)
        if res == 4:
            total += 2.0 * (oc + cfg.motion_out_dim) * oc * 9 * res * res
        else:
            ic = cfg.channels(res // 2)
            total += 2.0 * ic * oc * 9 * (res // 2) ** 2
            total += 2.0 * oc * oc * 9 * res * res
        total += 2.0 * oc * cfg.img_channels * res * res
    return total
