"""Adaptive discriminator augmentation pipe on the drop-in ops — the stage between G and D in every training phase when
`aug=ada` (the reference's default, configs/training/base.yaml; src/training/augment.py:117-436, used at loss.py:58-70).

Same constructor arguments, buffers (`p`, `Hz_geom`, `Hz_fbank`) and random-number stream as the reference class, so a run with the same
seed produces the same augmented frames (tests/test_augment_cpu.py compares against the unmodified reference on CPU).  What the heavy
steps run on with CUDA tensors:

    reflect pad -> 2x up-sampling with the 12-tap sym6 filter      upfirdn2d FIR kernels (separable: two 1-D passes, upfirdn2d.py:238-240)
    bilinear warp through the inverse transform                     grid_sample_gradfix (library sampler, gradient fix as the reference)
    2x down-sampling + crop (negative padding, flipped filter)      upfirdn2d FIR kernels
    per-sample 3x4 colour matrix                                    one batched GEMM on [B*F, 3, H*W]
    per-sample band amplification (imgfilter)                       two grouped 1-D convolutions (conv2d_gradfix -> library, groups = B*C)

Organisation differs from the reference's one long forward(): every augmentation is a (sampler, gate probability, identity value,
matrix builder) record; `_sample` draws value then gate — the order the reference consumes its generator in — and the records are folded
into the inverse geometric transform G_inv [B,3,3] and the colour transform C [B,4,4].  With `video_consistent_aug` the caller passes clips
as [B, F*3, H, W]; the colour matrix is then shared by the F frames of a clip (augment.py:352-356).

The one host synchronisation of the reference is kept: the reflect-padding margins are the batch maximum of the transformed corner
positions and size the padded tensor (augment.py:268-277).
"""
import numpy as np
import scipy.signal
import torch

from .ops import conv2d_gradfix, grid_sample_gradfix, upfirdn2d

# low-pass decomposition filters used by the pipe (Daubechies / symlet coefficients; augment.py:21-38 lists the full family)
_SYM2 = [-0.12940952255092145, 0.22414386804185735, 0.836516303737469, 0.48296291314469025]
_SYM6 = [0.015404109327027373, 0.0034907120842174702, -0.11799011114819057, -0.048311742585633, 0.4910559419267466, 0.787641141030194,
         0.3379294217276218, -0.07263752278646252, -0.021060292512300564, 0.04472490177066578, 0.0017677118642428036, -0.007800708325034148]


def _const(value, like=None, device=None, shape=None):
    t = torch.as_tensor(np.asarray(value), dtype=torch.float32, device=like.device if like is not None else device)
    if shape is not None:
        t = t.expand(shape) if t.ndim == 0 else t.reshape(shape)
    return t


def _mat(rows, device=None):
    """Homogeneous matrix from nested rows whose entries are python numbers or [B]-shaped tensors -> [rows, cols] or [B, rows, cols]."""
    flat = [e for r in rows for e in r]
    ref = next((e for e in flat if isinstance(e, torch.Tensor)), None)
    if ref is None:
        return _const(rows, device=device)
    cols = [e if isinstance(e, torch.Tensor) else _const(e, like=ref, shape=ref.shape) for e in flat]
    return torch.stack(cols, dim=-1).reshape(ref.shape + (len(rows), len(rows[0])))


def _translate2(tx, ty, **kw):
    return _mat([[1, 0, tx], [0, 1, ty], [0, 0, 1]], **kw)


def _scale2(sx, sy, **kw):
    return _mat([[sx, 0, 0], [0, sy, 0], [0, 0, 1]], **kw)


def _rotate2(theta, **kw):
    c, s = torch.cos(theta), torch.sin(theta)
    return _mat([[c, s * -1, 0], [s, c, 0], [0, 0, 1]], **kw)


def _translate3(tx, ty, tz, **kw):
    return _mat([[1, 0, 0, tx], [0, 1, 0, ty], [0, 0, 1, tz], [0, 0, 0, 1]], **kw)


def _scale3(sx, sy, sz, **kw):
    return _mat([[sx, 0, 0, 0], [0, sy, 0, 0], [0, 0, sz, 0], [0, 0, 0, 1]], **kw)


def _rotate3(axis, theta, **kw):
    """Rodrigues rotation about `axis` (augment.py:90-98)."""
    vx, vy, vz = axis[..., 0], axis[..., 1], axis[..., 2]
    s, c = torch.sin(theta), torch.cos(theta)
    cc = 1 - c
    return _mat([[vx * vx * cc + c, vx * vy * cc - vz * s, vx * vz * cc + vy * s, 0],
                 [vy * vx * cc + vz * s, vy * vy * cc + c, vy * vz * cc - vx * s, 0],
                 [vz * vx * cc - vy * s, vz * vy * cc + vx * s, vz * vz * cc + c, 0],
                 [0, 0, 0, 1]], **kw)


def _filter_bank():
    """Four band-pass filters built from the sym2 low-pass (augment.py:166-177): an octave filter bank with perfect reconstruction."""
    lo = np.asarray(_SYM2)
    hi = lo * ((-1) ** np.arange(lo.size))
    lo2 = np.convolve(lo, lo[::-1]) / 2
    hi2 = np.convolve(hi, hi[::-1]) / 2
    bank = np.eye(4, 1)
    for i in range(1, bank.shape[0]):
        bank = np.dstack([bank, np.zeros_like(bank)]).reshape(bank.shape[0], -1)[:, :-1]      # zero-stuff (up-sample by 2)
        bank = scipy.signal.convolve(bank, [lo2])
        mid = bank.shape[1]
        bank[i, (mid - hi2.size) // 2:(mid + hi2.size) // 2] += hi2
    return torch.as_tensor(bank, dtype=torch.float32)


class AugmentPipe(torch.nn.Module):
    def __init__(self, xflip=0, rotate90=0, xint=0, xint_max=0.125,
                 scale=0, rotate=0, aniso=0, xfrac=0, scale_std=0.2, rotate_max=1, aniso_std=0.2, xfrac_std=0.125,
                 brightness=0, contrast=0, lumaflip=0, hue=0, saturation=0, brightness_std=0.2, contrast_std=0.5, hue_max=1, saturation_std=1,
                 imgfilter=0, imgfilter_bands=(1, 1, 1, 1), imgfilter_std=1, noise=0, cutout=0, noise_std=0.1, cutout_size=0.5):
        super().__init__()
        self.register_buffer('p', torch.ones([]))            # overall probability multiplier, adapted by the training loop (training_loop.py:327-330)
        for k, v in dict(xflip=xflip, rotate90=rotate90, xint=xint, xint_max=xint_max, scale=scale, rotate=rotate, aniso=aniso, xfrac=xfrac,
                         scale_std=scale_std, rotate_max=rotate_max, aniso_std=aniso_std, xfrac_std=xfrac_std, brightness=brightness,
                         contrast=contrast, lumaflip=lumaflip, hue=hue, saturation=saturation, brightness_std=brightness_std,
                         contrast_std=contrast_std, hue_max=hue_max, saturation_std=saturation_std, imgfilter=imgfilter,
                         imgfilter_std=imgfilter_std, noise=noise, cutout=cutout, noise_std=noise_std, cutout_size=cutout_size).items():
            setattr(self, k, float(v))
        self.imgfilter_bands = list(imgfilter_bands)
        self.register_buffer('Hz_geom', upfirdn2d.setup_filter(_SYM6))
        self.register_buffer('Hz_fbank', _filter_bank())

    # -- random draws: value first, then the Bernoulli gate — the order the reference consumes the generator in ------------------------
    def _sample(self, draw, prob, identity, gate_shape, pct=None):
        value = draw()
        gate = torch.rand(gate_shape, device=value.device) < prob
        value = torch.where(gate, value, torch.full_like(value, identity) if not isinstance(identity, torch.Tensor) else identity)
        if pct is not None:
            value = pct(value)
        return value

    def forward(self, images, debug_percentile=None):
        assert isinstance(images, torch.Tensor) and images.ndim == 4
        B, C, H, W = images.shape
        dev = images.device
        dp = None if debug_percentile is None else torch.as_tensor(debug_percentile, dtype=torch.float32, device=dev)
        rand = lambda *s: torch.rand(list(s), device=dev)
        randn = lambda *s: torch.randn(list(s), device=dev)
        fill = lambda f: (None if dp is None else (lambda v: torch.full_like(v, f(dp))))       # debug mode: every sample gets the percentile's value
        erf = lambda std: (lambda q: torch.erfinv(q * 2 - 1) * std)
        p = self.p

        # ---- inverse geometric transform G_inv: output pixel -> input pixel ----------------------------------------------------------------
        eye3 = torch.eye(3, device=dev)
        G = eye3
        if self.xflip > 0:
            i = self._sample(lambda: torch.floor(rand(B) * 2), self.xflip * p, 0, [B], fill(lambda q: torch.floor(q * 2)))
            G = G @ _scale2(1 / (1 - 2 * i), 1)
        if self.rotate90 > 0:
            i = self._sample(lambda: torch.floor(rand(B) * 4), self.rotate90 * p, 0, [B], fill(lambda q: torch.floor(q * 4)))
            G = G @ _rotate2(np.pi / 2 * i)
        if self.xint > 0:
            t = self._sample(lambda: (rand(B, 2) * 2 - 1) * self.xint_max, self.xint * p, 0, [B, 1], fill(lambda q: (q * 2 - 1) * self.xint_max))
            G = G @ _translate2(-torch.round(t[:, 0] * W), -torch.round(t[:, 1] * H))
        if self.scale > 0:
            s = self._sample(lambda: torch.exp2(randn(B) * self.scale_std), self.scale * p, 1, [B], fill(lambda q: torch.exp2(erf(self.scale_std)(q))))
            G = G @ _scale2(1 / s, 1 / s)
        p_rot = 1 - torch.sqrt((1 - self.rotate * p).clamp(0, 1))                        # P(pre-rotation OR post-rotation) = rotate * p
        draw_theta = lambda: (rand(B) * 2 - 1) * np.pi * self.rotate_max
        if self.rotate > 0:
            th = self._sample(draw_theta, p_rot, 0, [B], fill(lambda q: (q * 2 - 1) * np.pi * self.rotate_max))
            G = G @ _rotate2(th)                                                          # before the anisotropic scaling
        if self.aniso > 0:
            s = self._sample(lambda: torch.exp2(randn(B) * self.aniso_std), self.aniso * p, 1, [B], fill(lambda q: torch.exp2(erf(self.aniso_std)(q))))
            G = G @ _scale2(1 / s, 1 / (1 / s))
        if self.rotate > 0:
            th = self._sample(draw_theta, p_rot, 0, [B], fill(lambda q: q * 0))
            G = G @ _rotate2(th)                                                          # after it
        if self.xfrac > 0:
            t = self._sample(lambda: randn(B, 2) * self.xfrac_std, self.xfrac * p, 0, [B, 1], fill(erf(self.xfrac_std)))
            G = G @ _translate2(-(t[:, 0] * W), -(t[:, 1] * H))

        if G is not eye3:
            images = self._warp(images, G)

        # ---- colour transform C: colour_in -> colour_out ---------------------------------------------------------------------------------
        eye4 = torch.eye(4, device=dev)
        Cm = eye4
        if self.brightness > 0:
            b = self._sample(lambda: randn(B) * self.brightness_std, self.brightness * p, 0, [B], fill(erf(self.brightness_std)))
            Cm = _translate3(b, b, b) @ Cm
        if self.contrast > 0:
            c = self._sample(lambda: torch.exp2(randn(B) * self.contrast_std), self.contrast * p, 1, [B], fill(lambda q: torch.exp2(erf(self.contrast_std)(q))))
            Cm = _scale3(c, c, c) @ Cm
        luma = _const(np.asarray([1, 1, 1, 0]) / np.sqrt(3), device=dev)
        if self.lumaflip > 0:
            i = self._sample(lambda: torch.floor(rand(B, 1, 1) * 2), self.lumaflip * p, 0, [B, 1, 1], fill(lambda q: torch.floor(q * 2)))
            Cm = (eye4 - 2 * luma.ger(luma) * i) @ Cm                                     # Householder reflection about the luma axis
        if self.hue > 0 and C > 1:
            th = self._sample(lambda: (rand(B) * 2 - 1) * np.pi * self.hue_max, self.hue * p, 0, [B], fill(lambda q: (q * 2 - 1) * np.pi * self.hue_max))
            Cm = _rotate3(luma, th) @ Cm
        if self.saturation > 0 and C > 1:
            s = self._sample(lambda: torch.exp2(randn(B, 1, 1) * self.saturation_std), self.saturation * p, 1, [B, 1, 1],
                             fill(lambda q: torch.exp2(erf(self.saturation_std)(q))))
            Cm = (luma.ger(luma) + (eye4 - luma.ger(luma)) * s) @ Cm
        if Cm is not eye4:
            images = self._recolour(images, Cm)

        if self.imgfilter > 0:
            images = self._band_filter(images, dp)

        # ---- corruptions --------------------------------------------------------------------------------------------------------------------
        if self.noise > 0:
            sigma = self._sample(lambda: randn(B, 1, 1, 1).abs() * self.noise_std, self.noise * p, 0, [B, 1, 1, 1],
                                 fill(lambda q: torch.erfinv(q) * self.noise_std))
            images = images + randn(B, C, H, W) * sigma
        if self.cutout > 0:
            size = torch.full([B, 2, 1, 1, 1], self.cutout_size, device=dev)
            size = torch.where(rand(B, 1, 1, 1, 1) < self.cutout * p, size, torch.zeros_like(size))
            center = rand(B, 2, 1, 1, 1)
            if dp is not None:
                size, center = torch.full_like(size, self.cutout_size), torch.full_like(center, dp)
            xs = (torch.arange(W, device=dev).reshape(1, 1, 1, -1) + 0.5) / W
            ys = (torch.arange(H, device=dev).reshape(1, 1, -1, 1) + 0.5) / H
            keep = torch.logical_or((xs - center[:, 0]).abs() >= size[:, 0] / 2, (ys - center[:, 1]).abs() >= size[:, 1] / 2)
            images = images * keep.to(torch.float32)
        return images

    # -- geometric execution: reflect pad, 2x up, warp, 2x down + crop (augment.py:263-296) --------------------------------------------------
    def _warp(self, images, G):
        B, C, H, W = images.shape
        dev = images.device
        cx, cy = (W - 1) / 2, (H - 1) / 2
        corners = _mat([[-cx, -cy, 1], [cx, -cy, 1], [cx, cy, 1], [-cx, cy, 1]], device=dev)            # image corners, centred coordinates
        moved = G @ corners.t()                                                                             # [B, xyz, corner]
        hz_pad = self.Hz_geom.shape[0] // 4
        reach = moved[:, :2, :].permute(1, 0, 2).flatten(1)                                                # [xy, B * corner]
        reach = torch.cat([-reach, reach]).max(dim=1).values                                               # furthest source coordinate per side
        reach = reach + _const([hz_pad * 2 - cx, hz_pad * 2 - cy] * 2, device=dev)
        reach = reach.max(_const([0, 0] * 2, device=dev)).min(_const([W - 1, H - 1] * 2, device=dev))
        mx0, my0, mx1, my1 = reach.ceil().to(torch.int32)                                                  # host sync: sizes the padded tensor
        images = torch.nn.functional.pad(input=images, pad=[mx0, mx1, my0, my1], mode='reflect')
        G = _translate2((mx0 - mx1) / 2, (my0 - my1) / 2) @ G
        images = upfirdn2d.upsample2d(x=images, f=self.Hz_geom, up=2)
        G = _scale2(2, 2, device=dev) @ G @ _scale2(1 / 2, 1 / 2, device=dev)
        G = _translate2(-0.5, -0.5, device=dev) @ G @ _translate2(0.5, 0.5, device=dev)
        shape = [B, C, (H + hz_pad * 2) * 2, (W + hz_pad * 2) * 2]
        G = _scale2(2 / images.shape[3], 2 / images.shape[2], device=dev) @ G @ _scale2(1 / (2 / shape[3]), 1 / (2 / shape[2]), device=dev)
        grid = torch.nn.functional.affine_grid(theta=G[:, :2, :], size=shape, align_corners=False)
        images = grid_sample_gradfix.grid_sample(images, grid)
        return upfirdn2d.downsample2d(x=images, f=self.Hz_geom, down=2, padding=-hz_pad * 2, flip_filter=True)

    # -- colour execution (augment.py:345-363) ------------------------------------------------------------------------------------------------
    def _recolour(self, images, Cm):
        B, C, H, W = images.shape
        x = images.reshape(B, C, H * W)
        if C > 3 and C % 3 == 0:                                     # a clip passed as [B, F*3, H, W]: the same colour matrix for its F frames
            x = x.reshape(B * (C // 3), 3, H * W)
            Cm = Cm.repeat_interleave(C // 3, dim=0)
        if C % 3 == 0:
            x = Cm[:, :3, :3] @ x + Cm[:, :3, 3:]
        elif C == 1:
            row = Cm[:, :3, :].mean(dim=1, keepdims=True)
            x = x * row[:, :, :3].sum(dim=2, keepdims=True) + row[:, :, 3:]
        else:
            raise ValueError('Image must be RGB (3 channels) or L (1 channel)')
        return x.reshape(B, C, H, W)

    # -- image-space filtering: random per-band amplification through the sym2 filter bank (augment.py:369-400) ----------------------------
    def _band_filter(self, images, dp):
        B, C, H, W = images.shape
        dev = images.device
        nb = self.Hz_fbank.shape[0]
        assert len(self.imgfilter_bands) == nb
        power = _const(np.array([10, 1, 1, 1]) / 13, device=dev)                                         # expected 1/f power spectrum
        gain = torch.ones([B, nb], device=dev)
        for i, strength in enumerate(self.imgfilter_bands):
            t_i = torch.exp2(torch.randn([B], device=dev) * self.imgfilter_std)
            t_i = torch.where(torch.rand([B], device=dev) < self.imgfilter * self.p * strength, t_i, torch.ones_like(t_i))
            if dp is not None:
                t_i = torch.full_like(t_i, torch.exp2(torch.erfinv(dp * 2 - 1) * self.imgfilter_std)) if strength > 0 else torch.ones_like(t_i)
            t = torch.ones([B, nb], device=dev)
            t[:, i] = t_i
            gain = gain * (t / (power * t.square()).sum(dim=-1, keepdims=True).sqrt())                  # power-normalised, accumulated over bands
        taps = (gain @ self.Hz_fbank).unsqueeze(1).repeat([1, C, 1]).reshape(B * C, 1, -1)               # one 1-D filter per (sample, channel)
        pad = self.Hz_fbank.shape[1] // 2
        x = torch.nn.functional.pad(input=images.reshape(1, B * C, H, W), pad=[pad, pad, pad, pad], mode='reflect')
        x = conv2d_gradfix.conv2d(input=x, weight=taps.unsqueeze(2), groups=B * C)
        x = conv2d_gradfix.conv2d(input=x, weight=taps.unsqueeze(3), groups=B * C)
        return x.reshape(B, C, H, W)
