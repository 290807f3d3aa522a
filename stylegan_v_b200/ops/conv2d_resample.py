"""Drop-in for `src.torch_utils.ops.conv2d_resample` (reference: src/torch_utils/ops/conv2d_resample.py).

`conv2d_resample(x, w, f, up, down, padding, groups, flip_weight, flip_filter)` lowers a k x k convolution with
optional 2x resampling onto the dense contraction (conv2d_gradfix) plus FIR passes (upfirdn2d).  The padding
arithmetic (conv2d_resample.py:95-104) and the choice between the six execution plans (:106-154) are the
reference's; they are expressed here as an explicit plan (`_plan`) that is built once per call and then run.
"""
import torch

from . import conv2d_gradfix
from . import upfirdn2d
from .upfirdn2d import _parse_padding, _get_filter_size


def _get_weight_shape(w):
    return [int(s) for s in w.shape]


def _conv2d_wrapper(x, w, stride=1, padding=0, groups=1, transpose=False, flip_weight=True):
    """Dense contraction.  F.conv2d is a correlation, so `flip_weight=False` (true convolution) flips w first."""
    oc, icpg, kh, kw = _get_weight_shape(w)
    if not flip_weight:
        w = w.flip([2, 3])
    # channels_last 1x1 with few channels: the reference sidesteps a cuDNN 8.0.5 pitfall here
    # (conv2d_resample.py:38-50); kept because it also defines the output memory format callers see.
    if kw == 1 and kh == 1 and stride == 1 and padding in (0, [0, 0], (0, 0)) and not transpose:
        if x.stride(1) == 1 and min(oc, icpg) < 64:
            if oc <= 4 and groups == 1:
                shp = x.shape
                y = w.squeeze(3).squeeze(2) @ x.reshape(shp[0], icpg, -1)
                y = y.reshape(shp[0], oc, shp[2], shp[3])
            else:
                y = conv2d_gradfix.conv2d(x.contiguous(), w.contiguous(), groups=groups)
            return y.to(memory_format=torch.channels_last)
    op = conv2d_gradfix.conv_transpose2d if transpose else conv2d_gradfix.conv2d
    return op(x, w, stride=stride, padding=padding, groups=groups)


def _plan(kw, kh, fw, fh, up, down, padding):
    """Returns (kind, fir padding / conv padding) for the reference's fast paths."""
    px0, px1, py0, py1 = _parse_padding(padding)
    if up > 1:
        px0 += (fw + up - 1) // 2; px1 += (fw - up) // 2
        py0 += (fh + up - 1) // 2; py1 += (fh - up) // 2
    if down > 1:
        px0 += (fw - down + 1) // 2; px1 += (fw - down) // 2
        py0 += (fh - down + 1) // 2; py1 += (fh - down) // 2
    one = (kw == 1 and kh == 1)
    if one and down > 1 and up == 1:
        return 'down_then_1x1', (px0, px1, py0, py1)
    if one and up > 1 and down == 1:
        return '1x1_then_up', (px0, px1, py0, py1)
    if down > 1 and up == 1:
        return 'blur_then_strided', (px0, px1, py0, py1)
    if up > 1:
        return 'transposed_then_blur', (px0 - (kw - 1), px1 - (kw - up), py0 - (kh - 1), py1 - (kh - up))
    if px0 == px1 and py0 == py1 and px0 >= 0 and py0 >= 0:
        return 'plain', (px0, px1, py0, py1)
    return 'generic', (px0, px1, py0, py1)


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    assert isinstance(x, torch.Tensor) and x.ndim == 4
    assert isinstance(w, torch.Tensor) and w.ndim == 4 and w.dtype == x.dtype
    assert f is None or (isinstance(f, torch.Tensor) and f.ndim in (1, 2) and f.dtype == torch.float32)
    assert isinstance(up, int) and up >= 1 and isinstance(down, int) and down >= 1
    assert isinstance(groups, int) and groups >= 1
    oc, icpg, kh, kw = _get_weight_shape(w)
    fw, fh = _get_filter_size(f)
    kind, (px0, px1, py0, py1) = _plan(kw, kh, fw, fh, up, down, padding)
    fir = upfirdn2d.upfirdn2d

    if kind == 'down_then_1x1':
        x = fir(x=x, f=f, down=down, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return _conv2d_wrapper(x=x, w=w, groups=groups, flip_weight=flip_weight)
    if kind == '1x1_then_up':
        x = _conv2d_wrapper(x=x, w=w, groups=groups, flip_weight=flip_weight)
        return fir(x=x, f=f, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    if kind == 'blur_then_strided':
        x = fir(x=x, f=f, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return _conv2d_wrapper(x=x, w=w, stride=down, groups=groups, flip_weight=flip_weight)
    if kind == 'transposed_then_blur':
        if groups == 1:
            w = w.transpose(0, 1)
        else:
            w = w.reshape(groups, oc // groups, icpg, kh, kw).transpose(1, 2).reshape(groups * icpg, oc // groups, kh, kw)
        pxt = max(min(-px0, -px1), 0)
        pyt = max(min(-py0, -py1), 0)
        x = _conv2d_wrapper(x=x, w=w, stride=up, padding=[pyt, pxt], groups=groups, transpose=True, flip_weight=(not flip_weight))
        x = fir(x=x, f=f, padding=[px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt], gain=up ** 2, flip_filter=flip_filter)
        if down > 1:
            x = fir(x=x, f=f, down=down, flip_filter=flip_filter)
        return x
    if kind == 'plain':
        return _conv2d_wrapper(x=x, w=w, padding=[py0, px0], groups=groups, flip_weight=flip_weight)
    x = fir(x=x, f=(f if up > 1 else None), up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    x = _conv2d_wrapper(x=x, w=w, groups=groups, flip_weight=flip_weight)
    if down > 1:
        x = fir(x=x, f=f, down=down, flip_filter=flip_filter)
    return x
