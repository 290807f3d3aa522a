"""Drop-in for `src.torch_utils.ops.upfirdn2d` (reference: src/torch_utils/ops/upfirdn2d.py).

Same public surface — setup_filter, upfirdn2d, filter2d, upsample2d, downsample2d and the helpers other
reference modules import by name (_parse_scaling, _parse_padding, _get_filter_size; conv2d_resample.py:16-17)
— with the CUDA work done by libsgv_b200 (stylegan_v_b200/csrc/upfirdn2d.cu) through the C ABI.

Semantics kept from the reference:
  * `impl='cuda'` on a CUDA tensor runs the native kernel; a CPU tensor, or `impl='ref'`, runs the
    standard-PyTorch-ops formulation (upfirdn2d.py:162-164).  A CUDA tensor never falls back: if the
    native library is missing the call raises.
  * separable (1-D) filters run as two passes with sqrt(gain) each (upfirdn2d.py:236-240);
  * gradients of any order: the backward of a pass is another pass with up/down swapped, the filter
    flipped and padding from upfirdn2d.py:251-256.
"""
import numpy as np
import torch

from .. import plugin as _plugin


# --------------------------------------------------------------------------- argument helpers
def _parse_scaling(scaling):
    if isinstance(scaling, int):
        scaling = [scaling, scaling]
    assert isinstance(scaling, (list, tuple)) and all(isinstance(v, int) for v in scaling)
    sx, sy = scaling
    assert sx >= 1 and sy >= 1
    return sx, sy


def _parse_padding(padding):
    if isinstance(padding, int):
        padding = [padding, padding]
    assert isinstance(padding, (list, tuple)) and all(isinstance(v, int) for v in padding)
    if len(padding) == 2:
        px, py = padding
        padding = [px, px, py, py]
    px0, px1, py0, py1 = padding
    return px0, px1, py0, py1


def _get_filter_size(f):
    if f is None:
        return 1, 1
    assert isinstance(f, torch.Tensor) and f.ndim in (1, 2)
    fw, fh = int(f.shape[-1]), int(f.shape[0])
    assert fw >= 1 and fh >= 1
    return fw, fh


def setup_filter(f, device=torch.device('cpu'), normalize=True, flip_filter=False, gain=1, separable=None):
    """Builds the float32 FIR tensor `upfirdn2d()` expects (reference: upfirdn2d.py:72-116)."""
    taps = torch.as_tensor(1 if f is None else f, dtype=torch.float32)
    assert taps.ndim in (0, 1, 2) and taps.numel() > 0
    if taps.ndim == 0:
        taps = taps.reshape(1)
    if separable is None:
        separable = taps.ndim == 1 and taps.numel() >= 8
    if taps.ndim == 1 and not separable:
        taps = torch.outer(taps, taps)
    assert taps.ndim == (1 if separable else 2)
    if normalize:
        taps = taps / taps.sum()
    if flip_filter:
        taps = taps.flip(list(range(taps.ndim)))
    taps = taps * (gain ** (taps.ndim / 2))
    return taps.to(device=device)


# --------------------------------------------------------------------------- native path
class _FirPass(torch.autograd.Function):
    """x -> upfirdn2d(x, f) for one (possibly separable) filter; differentiable to any order."""

    @staticmethod
    def forward(ctx, x, f, up, down, pad, flip, gain):
        assert isinstance(x, torch.Tensor) and x.ndim == 4
        if f is None:
            f = torch.ones([1, 1], dtype=torch.float32, device=x.device)
        assert isinstance(f, torch.Tensor) and f.ndim in (1, 2)
        (ux, uy), (dx, dy), (px0, px1, py0, py1) = up, down, pad
        if f.ndim == 2:
            y = _plugin.upfirdn2d(x, f, ux, uy, dx, dy, px0, px1, py0, py1, flip, gain)
        else:
            g = float(np.sqrt(gain))
            y = _plugin.upfirdn2d(x, f.unsqueeze(0), ux, 1, dx, 1, px0, px1, 0, 0, flip, g)
            y = _plugin.upfirdn2d(y, f.unsqueeze(1), 1, uy, 1, dy, 0, 0, py0, py1, flip, g)
        ctx.save_for_backward(f)
        ctx.cfg = (up, down, pad, flip, gain, tuple(x.shape))
        return y

    @staticmethod
    def backward(ctx, dy):
        f, = ctx.saved_tensors
        (ux, uy), (dx_, dy_), (px0, _px1, py0, _py1), flip, gain, (_, _, ih, iw) = ctx.cfg
        _, _, oh, ow = dy.shape
        fw, fh = _get_filter_size(f)
        gpad = (fw - px0 - 1, iw * ux - ow * dx_ + px0 - ux + 1,
                fh - py0 - 1, ih * uy - oh * dy_ + py0 - uy + 1)
        gx = None
        if ctx.needs_input_grad[0]:
            gx = _FirPass.apply(dy, f, (dx_, dy_), (ux, uy), gpad, not flip, gain)
        return gx, None, None, None, None, None, None


# --------------------------------------------------------------------------- standard-ops path
def _upfirdn2d_ref(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1):
    """Formulation with stock PyTorch ops (CPU tensors / impl='ref'); mirrors upfirdn2d.py:169-208."""
    assert isinstance(x, torch.Tensor) and x.ndim == 4
    if f is None:
        f = torch.ones([1, 1], dtype=torch.float32, device=x.device)
    assert isinstance(f, torch.Tensor) and f.ndim in (1, 2) and f.dtype == torch.float32 and not f.requires_grad
    n, c, h, w = x.shape
    ux, uy = _parse_scaling(up)
    dx, dy = _parse_scaling(down)
    px0, px1, py0, py1 = _parse_padding(padding)
    F = torch.nn.functional
    # zero insertion: one real sample followed by (up-1) zeros
    x = F.pad(x.reshape(n, c, h, 1, w, 1), [0, ux - 1, 0, 0, 0, uy - 1]).reshape(n, c, h * uy, w * ux)
    # positive padding pads, negative padding crops
    x = F.pad(x, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    x = x[:, :, max(-py0, 0): x.shape[2] - max(-py1, 0), max(-px0, 0): x.shape[3] - max(-px1, 0)]
    k = (f * (gain ** (f.ndim / 2))).to(x.dtype)
    if not flip_filter:
        k = k.flip(list(range(k.ndim)))
    k = k[None, None].repeat([c, 1] + [1] * k.ndim)
    if k.ndim == 4:
        x = F.conv2d(x, k, groups=c)
    else:
        x = F.conv2d(x, k.unsqueeze(2), groups=c)
        x = F.conv2d(x, k.unsqueeze(3), groups=c)
    return x[:, :, ::dy, ::dx]


# --------------------------------------------------------------------------- public API
def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """Pad, upsample (zero insertion), FIR-filter and downsample a batch of 2-D images [N,C,H,W]."""
    assert isinstance(x, torch.Tensor)
    assert impl in ('ref', 'cuda')
    if impl == 'cuda' and x.device.type == 'cuda':
        return _FirPass.apply(x, f, _parse_scaling(up), _parse_scaling(down), _parse_padding(padding), bool(flip_filter), gain)
    return _upfirdn2d_ref(x, f, up=up, down=down, padding=padding, flip_filter=flip_filter, gain=gain)


def filter2d(x, f, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """FIR-filter keeping the input extent (plus user padding)."""
    px0, px1, py0, py1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    pad = [px0 + fw // 2, px1 + (fw - 1) // 2, py0 + fh // 2, py1 + (fh - 1) // 2]
    return upfirdn2d(x, f, padding=pad, flip_filter=flip_filter, gain=gain, impl=impl)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """Upsample by `up` so that the output extent is a multiple of the input's."""
    ux, uy = _parse_scaling(up)
    px0, px1, py0, py1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    pad = [px0 + (fw + ux - 1) // 2, px1 + (fw - ux) // 2, py0 + (fh + uy - 1) // 2, py1 + (fh - uy) // 2]
    return upfirdn2d(x, f, up=up, padding=pad, flip_filter=flip_filter, gain=gain * ux * uy, impl=impl)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """Downsample by `down` so that the output extent is a fraction of the input's."""
    dx, dy = _parse_scaling(down)
    px0, px1, py0, py1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    pad = [px0 + (fw - dx + 1) // 2, px1 + (fw - dx) // 2, py0 + (fh - dy + 1) // 2, py1 + (fh - dy) // 2]
    return upfirdn2d(x, f, down=down, padding=pad, flip_filter=flip_filter, gain=gain, impl=impl)
