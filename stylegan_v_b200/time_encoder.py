"""Continuous Fourier time-encoder of StyleGAN-V (motion codes), native-path module.

Parameter names and shapes follow the reference's MotionMappingNetwork / AlignedTimeEncoder
(src/training/motion.py:18-156,160-214) so reference checkpoints load: `conv.{0,1}.{weight,bias}`,
`time_encoder.{periods,phase,aligners}_predictor.weight`, buffers `time_encoder.{freqs,phase_scales}`.

Computation (same maths, reorganised):
  trajectory     z ~ N(0,1) [B, L, 512]  ->  two equalised-lr Conv1d(k=11, no padding) + lrelu   (motion.py:55-58,100)
  neighbours     left = floor(t / 16), right = left + 1, alpha = frac(t / 16)                    (motion.py:105-115)
  embedding      emb(tau) = [sin, cos](freqs * (tanh(P u_L) + 1) * tau + (Phi u_L) * phase_scales)
                 v = emb(t) - lerp(emb(t_L), emb(t_R), alpha) + lerp(A u_L, A u_R, alpha)         (motion.py:198-212)
The three predictors on u_L share one stacked GEMM ([P; Phi; A]).  Arguments of sin/cos reach ~800 rad, so accurate
(not fast-math) sin/cos are required for parity.  On CUDA the whole elementwise tail (remainder / neighbour positions,
tanh, three phase arguments, sin/cos, both lerps) is ONE kernel forward and ONE backward (csrc/time_encoder.cu,
`sgv_time_encoder_fwd/_bwd`) instead of ~25 + ~60 PyTorch launches; CPU tensors evaluate the same expression with
PyTorch ops like the reference does.
"""
import ctypes

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib
from . import dense as _dense
from . import precision as _precision


class _TimeEncoderTail(torch.autograd.Function):
    """out = emb(t) - lerp(emb(t_L), emb(t_R), a) + lerp(A u_L, A u_R, a) from the stacked predictor outputs (CUDA, fp32)."""

    @staticmethod
    def forward(ctx, heads_left, aligners_right, t, freqs, phase_scales, d):
        heads_left, aligners_right = heads_left.contiguous(), aligners_right.contiguous()
        t = t.to(torch.float32).contiguous()
        freqs, phase_scales = freqs.reshape(-1).contiguous(), phase_scales.reshape(-1).contiguous()
        m, nf = t.numel(), freqs.numel()
        assert heads_left.shape == (m, 4 * nf) and aligners_right.shape == (m, 2 * nf)
        assert heads_left.dtype == aligners_right.dtype == freqs.dtype == phase_scales.dtype == torch.float32
        out = torch.empty([m, 2 * nf], dtype=torch.float32, device=t.device)
        with torch.cuda.device(t.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)
            _lib.check(_lib.lib().sgv_time_encoder_fwd(heads_left.data_ptr(), aligners_right.data_ptr(), t.data_ptr(), freqs.data_ptr(),
                                                       phase_scales.data_ptr(), out.data_ptr(), m, nf, float(d), stream), 'sgv_time_encoder_fwd')
        ctx.save_for_backward(heads_left, t, freqs, phase_scales)
        ctx.d = float(d)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout):
        heads_left, t, freqs, phase_scales = ctx.saved_tensors
        m, nf = t.numel(), freqs.numel()
        dout = dout.to(torch.float32).contiguous()
        dhl = torch.empty_like(heads_left)
        dar = torch.empty([m, 2 * nf], dtype=torch.float32, device=t.device)
        with torch.cuda.device(t.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)
            _lib.check(_lib.lib().sgv_time_encoder_bwd(dout.data_ptr(), heads_left.data_ptr(), t.data_ptr(), freqs.data_ptr(), phase_scales.data_ptr(),
                                                       dhl.data_ptr(), dar.data_ptr(), m, nf, ctx.d, stream), 'sgv_time_encoder_bwd')
        return dhl, dar, None, None, None, None


def linspaced_frequencies(num_freqs, min_period_len, max_period_len):
    """2*pi / 2^linspace(log2 min, log2 max), highest frequency last (motion.py:218-222)."""
    periods = 2.0 ** np.linspace(np.log2(min_period_len), np.log2(max_period_len), num_freqs)
    return torch.from_numpy((2 * np.pi / periods)[::-1].copy().astype(np.float32)).unsqueeze(0)


class EqualizedLinear(torch.nn.Module):
    """Bias-free or biased dense layer with runtime weight scaling lr_mult / sqrt(fan_in) (layers.py:108-138)."""

    def __init__(self, in_features, out_features, bias=True, lr_multiplier=1.0, bias_init=0.0):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.randn(out_features, in_features) / lr_multiplier)
        self.bias = torch.nn.Parameter(torch.full([out_features], float(bias_init))) if bias else None
        self.weight_gain = lr_multiplier / np.sqrt(in_features)
        self.bias_gain = lr_multiplier

    def forward(self, x, fused=False):
        """fused=True (first-order callers: the motion encoder, whose nodes are never on a path that is differentiated twice): one launch of
        the exact-fp32 dense kernel with gains + bias in its epilogue; default: the reference's addmm / matmul formulation, differentiable to
        any order (the style affines sit between ws and the image, which path-length regularisation differentiates twice, loss.py:101-119)."""
        if fused and x.ndim == 2 and _dense.supported(x, self.weight):
            return _dense.linear(x, self.weight, self.bias, self.weight_gain, self.bias_gain)
        w = self.weight.to(x.dtype) * self.weight_gain
        if self.bias is None:
            return x.matmul(w.t())
        b = self.bias.to(x.dtype)
        if self.bias_gain != 1:
            b = b * self.bias_gain
        return torch.addmm(b.unsqueeze(0), x, w.t())


class _Conv1dFp32Fwd(torch.autograd.Function):
    """conv1d whose FORWARD is true fp32 (see EqualizedConv1d.forward) while the backward may use TF32 tensor cores:
    gradient noise of 1e-3 is harmless, and the fp32 cuDNN weight-gradient engine costs 1.1 ms per step here."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        ctx.tf32_backward = not _precision.is_x3()          # fp32-grade mode: the backward is true fp32 as well
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            return F.conv1d(x, w, b)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=ctx.tf32_backward):
            gx, gw, gb = torch.ops.aten.convolution_backward(gy, x, w, [w.shape[0]], [1], [0], [1], False, [0], 1,
                                                            [ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]])
        return gx, gw, gb


class EqualizedConv1d(torch.nn.Module):
    """Conv1d with equalised learning rate + leaky ReLU (layers.py:331-373)."""

    def __init__(self, in_features, out_features, kernel_size, lr_multiplier=1.0):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.randn(out_features, in_features, kernel_size) / lr_multiplier)
        self.bias = torch.nn.Parameter(torch.zeros(out_features))
        self.weight_gain = lr_multiplier / np.sqrt(in_features * kernel_size)
        self.bias_gain = lr_multiplier

    def forward(self, x):
        # true fp32: the embedding multiplies these features by phase scales up to 64 and takes sin/cos, so TF32 rounding
        # here (PyTorch's cuDNN default) would show up as ~1e-2 errors in motion_v; the reference trains with allow_tf32=False
        y = _Conv1dFp32Fwd.apply(x, self.weight * self.weight_gain, self.bias * self.bias_gain)
        return F.leaky_relu(y, 0.2)


class AlignedTimeEncoder(torch.nn.Module):
    def __init__(self, latent_dim, num_freqs, min_period_len, max_period_len):
        super().__init__()
        freqs = linspaced_frequencies(num_freqs, min_period_len, max_period_len)
        self.register_buffer('freqs', freqs)
        self.register_buffer('phase_scales', max_period_len / (2 * np.pi / freqs))
        self.periods_predictor = EqualizedLinear(latent_dim, num_freqs, bias=False)
        self.phase_predictor = EqualizedLinear(latent_dim, num_freqs, bias=False)
        self.aligners_predictor = EqualizedLinear(latent_dim, num_freqs * 2, bias=False)

    def get_dim(self):
        return self.freqs.shape[1] * 2

    def forward(self, t, u_left, u_right, alpha, t_left, t_right, motion_z_distance=None):
        """t, t_left, t_right [M]; u_left, u_right [M, latent]; alpha [M, 1] -> [M, 2*num_freqs].
        With `motion_z_distance` (d) given, CUDA inputs run the fused tail kernel, which derives t_left / t_right / alpha from
        (t, d) itself exactly as MotionMappingNetwork does (motion.py:111-115)."""
        nf = self.freqs.shape[1]
        # one stacked GEMM for the three heads on u_left, one for the aligners on u_right.  Exact fp32 on purpose (csrc/dense_f32.cu; the
        # reference's library GEMMs for CPU tensors): the phases are multiplied by phase_scales up to 64 before sin / cos, so these products and
        # the trajectory conv1d below need round-to-nearest fp32 accumulation.  Measured on the B200 (profiles/dense_precision_r2.txt): with
        # the tcgen05 tf32x3 kernels, whose accumulator truncates (error ~K, 4e-5 at K = 5632), motion_v is off by 2e-3 against the fp32 CPU
        # evaluation; in fp32 by 1e-4.
        heads = torch.cat([self.periods_predictor.weight, self.phase_predictor.weight, self.aligners_predictor.weight], dim=0)
        gain = self.periods_predictor.weight_gain
        if _dense.supported(u_left, heads):
            hl = _dense.linear(u_left, heads, None, gain)
        else:
            hl = u_left.matmul((heads * gain).t())
        if motion_z_distance is not None and t.is_cuda and hl.dtype == torch.float32:
            return _TimeEncoderTail.apply(hl, self.aligners_predictor(u_right, fused=True), t.reshape(-1), self.freqs, self.phase_scales, motion_z_distance)
        periods = hl[:, :nf].tanh() + 1
        phases = hl[:, nf:2 * nf]
        al_left = hl[:, 2 * nf:]
        al_right = self.aligners_predictor(u_right)
        base = self.freqs * periods
        shift = phases * self.phase_scales

        def emb(tau):
            raw = base * tau.reshape(-1, 1).float() + shift
            return torch.cat([raw.sin(), raw.cos()], dim=1)
        remove = emb(t_left) * (1 - alpha) + emb(t_right) * alpha
        add = al_left * (1 - alpha) + al_right * alpha
        return emb(t) - remove + add


def trajectory_slabs(left, B, Fr, L, k, C):
    """Which slabs of the [B, L, C] noise sequence the two valid conv1d layers (k taps each) must be evaluated on so that trajectory positions
    left and left + 1 of every frame exist (motion.py:100-115).  Returns (windows, left, base):
      windows = True   one slab of k + 1 layer-1 outputs (2 k input positions) per FRAME, starting at input position left[b, f]:
                       base[b * Fr + f] = (b * L + left) * C element offsets; chosen while Fr * (k + 1) <= L - k + 1, i.e. while the windows
                       are cheaper than the whole trajectory;
      windows = False  one slab per CLIP covering the whole sequence: base[b] = b * L * C.
    left is clamped into [0, L - 2 k] (the last position whose right neighbour exists; the reference would raise on an out-of-range index)."""
    dev = left.device
    if Fr * (k + 1) <= L - k + 1:
        left = left.clamp(0, L - 2 * k)
        return True, left, ((torch.arange(B, device=dev).unsqueeze(1) * L + left) * C).reshape(-1)
    return False, left, torch.arange(B, device=dev) * (L * C)


class MotionMappingNetwork(torch.nn.Module):
    def __init__(self, z_dim=512, v_dim=512, kernel_size=11, motion_z_distance=16, time_enc_dim=256,
                 min_period_len=16, max_period_len=1024, max_num_frames=1024):
        super().__init__()
        self.z_dim, self.v_dim = z_dim, v_dim
        self.motion_z_distance = motion_z_distance
        self.max_num_frames = max_num_frames
        self.num_additional_codes = (kernel_size - 1) * 2
        self.conv = torch.nn.Sequential(EqualizedConv1d(z_dim, z_dim, kernel_size, lr_multiplier=0.01),
                                        EqualizedConv1d(z_dim, v_dim, kernel_size, lr_multiplier=0.01))
        self.time_encoder = AlignedTimeEncoder(v_dim, time_enc_dim, min_period_len, max_period_len)

    def get_dim(self):
        return self.time_encoder.get_dim()

    def traj_len(self, t_max=None):
        """Trajectory length the reference derives from max(t) (motion.py:63-66,80).  Passing t_max=None assumes
        t <= max_num_frames - 1 (true for the training sampler) and avoids the reference's host sync on t.max()."""
        max_t = self.max_num_frames - 1 if t_max is None else max(self.max_num_frames - 1, t_max)
        return int(np.ceil(max_t / self.motion_z_distance)) + 2 + self.num_additional_codes

    def forward(self, t, motion_z=None, t_max=None):
        """t [B, F] (float frame positions) -> dict(motion_v [B*F, dim], motion_z)."""
        B, Fr = t.shape
        L = self.traj_len(t_max)
        if motion_z is None:
            motion_z = torch.randn(B, L, self.z_dim, device=t.device)
        d = self.motion_z_distance
        left = (t / d).floor().long()
        z = motion_z[:B, :L, :self.z_dim]
        c0, c1 = self.conv[0], self.conv[1]
        k = c0.weight.shape[2]
        if z.is_cuda and z.dtype == torch.float32 and not z.requires_grad and c0.weight.shape[1] % 4 == 0 and c1.weight.shape[1] % 4 == 0:
            # the two valid conv1d layers as exact-fp32 GEMMs over windows of the [B, L, C] sequence (stylegan_v_b200/dense.py), evaluated only
            # where the result is read: a frame needs trajectory positions left, left + 1 = 2 outputs of layer 2 = k + 1 outputs of layer 1.
            # With many frames per clip the windows overlap enough that the whole trajectory is cheaper; then every position is computed once.
            C = z.shape[2]
            z = z.contiguous()
            windows, left, base = trajectory_slabs(left, B, Fr, L, k, C)
            if windows:
                y1 = _dense.conv1d_slabs(z, base, k + 1, c0.weight, c0.bias, c0.weight_gain, c0.bias_gain, 'lrelu')          # [B*F, k+1, C]
                y2 = _dense.conv1d_slabs(y1, None, 2, c1.weight, c1.bias, c1.weight_gain, c1.bias_gain, 'lrelu')             # [B*F, 2, v]
                u_left, u_right = y2[:, 0], y2[:, 1]
            else:
                y1 = _dense.conv1d_slabs(z, base, L - k + 1, c0.weight, c0.bias, c0.weight_gain, c0.bias_gain, 'lrelu')      # [B, L-k+1, C]
                trajs = _dense.conv1d_slabs(y1, None, L - 2 * k + 2, c1.weight, c1.bias, c1.weight_gain, c1.bias_gain, 'lrelu')
                rows = torch.arange(B, device=t.device).unsqueeze(1).expand(B, Fr)
                u_left = trajs[rows, left].reshape(B * Fr, -1)
                u_right = trajs[rows, left + 1].reshape(B * Fr, -1)
        else:
            trajs = self.conv(z.permute(0, 2, 1)).permute(0, 2, 1)                   # [B, L - 20, v_dim]
            rows = torch.arange(B, device=t.device).unsqueeze(1).expand(B, Fr)
            u_left = trajs[rows, left].reshape(B * Fr, -1)
            u_right = trajs[rows, left + 1].reshape(B * Fr, -1)
        if t.is_cuda and (t.dtype == torch.float32 or not t.dtype.is_floating_point):
            # fused tail: remainder / neighbour positions / alpha are derived from (t, d) inside the kernel (exact for integer frame indices too)
            v = self.time_encoder(t.reshape(-1).to(torch.float32), u_left, u_right, None, None, None, motion_z_distance=d)
        else:
            t_left = t - t % d
            alpha = ((t % d) / d).reshape(-1, 1).to(torch.float32)
            v = self.time_encoder(t.reshape(-1), u_left, u_right, alpha, t_left.reshape(-1), (t_left + d).reshape(-1))
        return dict(motion_v=v, motion_z=motion_z)
