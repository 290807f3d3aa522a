"""Drop-in for `src.torch_utils.ops.bias_act` (reference: src/torch_utils/ops/bias_act.py).

`bias_act(x, b, dim, act, alpha, gain, clamp, impl)` and the `activation_funcs` table keep the reference's
names, defaults and indices (bias_act.py:23-33,55).  CUDA tensors run stylegan_v_b200/csrc/bias_act.cu through
the C ABI and raise if the native library is missing; CPU tensors / impl='ref' use stock PyTorch ops
(bias_act.py:87-89).  First- and second-order gradients are supported exactly as in the reference
(forward saves x/b/y according to each activation's `ref` field, bias_act.py:154-158).

Beyond the reference: the bias gradient is reduced inside the gradient kernel (`db_accum`) instead of a
separate `dx.sum(...)` pass (bias_act.py:173) when the layout allows it.
"""
import numpy as np
import torch

from .. import plugin as _plugin


class _Spec(dict):
    __getattr__ = dict.__getitem__


def _mk(func, def_alpha, def_gain, cuda_idx, ref, has_2nd_grad):
    return _Spec(func=func, def_alpha=def_alpha, def_gain=def_gain, cuda_idx=cuda_idx, ref=ref, has_2nd_grad=has_2nd_grad)


_F = torch.nn.functional
_SQRT2 = float(np.sqrt(2))
activation_funcs = {
    'linear':   _mk(lambda x, **_: x,                            0,   1,      1, '',  False),
    'relu':     _mk(lambda x, **_: _F.relu(x),                   0,   _SQRT2, 2, 'y', False),
    'lrelu':    _mk(lambda x, alpha, **_: _F.leaky_relu(x, alpha), 0.2, _SQRT2, 3, 'y', False),
    'tanh':     _mk(lambda x, **_: torch.tanh(x),                0,   1,      4, 'y', True),
    'sigmoid':  _mk(lambda x, **_: torch.sigmoid(x),             0,   1,      5, 'y', True),
    'elu':      _mk(lambda x, **_: _F.elu(x),                    0,   1,      6, 'y', True),
    'selu':     _mk(lambda x, **_: _F.selu(x),                   0,   1,      7, 'y', True),
    'softplus': _mk(lambda x, **_: _F.softplus(x),               0,   1,      8, 'y', True),
    'swish':    _mk(lambda x, **_: torch.sigmoid(x) * x,         0,   _SQRT2, 9, 'x', True),
}

_EMPTY = torch.empty([0])


def _none(device, dtype):
    return _EMPTY.to(device=device, dtype=dtype)


def _fmt(x):
    return torch.channels_last if x.ndim > 2 and x.stride(1) == 1 else torch.contiguous_format


def _resolve(act, alpha, gain, clamp):
    spec = activation_funcs[act]
    assert clamp is None or clamp >= 0
    return (spec,
            float(spec.def_alpha if alpha is None else alpha),
            float(spec.def_gain if gain is None else gain),
            float(-1 if clamp is None else clamp))


class _BiasActFwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, b, dim, act, alpha, gain, clamp):
        spec = activation_funcs[act]
        ctx.fmt = _fmt(x)
        x = x.contiguous(memory_format=ctx.fmt)
        nil = _none(x.device, x.dtype)
        bb = b.contiguous() if b is not None else nil
        y = x
        if act != 'linear' or gain != 1 or clamp >= 0 or b is not None:
            y = _plugin.bias_act(x, bb, nil, nil, nil, 0, dim, spec.cuda_idx, alpha, gain, clamp)
        keep_x = 'x' in spec.ref or spec.has_2nd_grad
        ctx.save_for_backward(x if keep_x else nil, bb if keep_x else nil, y if 'y' in spec.ref else nil)
        ctx.cfg = (dim, act, alpha, gain, clamp, b is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        dim, act, alpha, gain, clamp, has_b = ctx.cfg
        dy = dy.contiguous(memory_format=ctx.fmt)
        x, b, y = ctx.saved_tensors
        dx = db = None
        if ctx.needs_input_grad[0] or (has_b and ctx.needs_input_grad[1]):
            want_db = has_b and ctx.needs_input_grad[1]
            if act != 'linear' or gain != 1 or clamp >= 0:
                fuse = want_db and dy.dtype == torch.float32 and not torch.is_grad_enabled()
                dx, db = _BiasActGrad.apply(dy, x, b, y, dim, act, alpha, gain, clamp, fuse)
                if not fuse:
                    db = None
            else:
                dx = dy
            if want_db and db is None:
                db = dx.sum([i for i in range(dx.ndim) if i != dim])
        return dx, db, None, None, None, None, None


class _BiasActGrad(torch.autograd.Function):
    """dx = d bias_act / dx applied to dy; optionally also returns the fused per-channel sum of dx."""

    @staticmethod
    def forward(ctx, dy, x, b, y, dim, act, alpha, gain, clamp, fuse_db):
        spec = activation_funcs[act]
        ctx.fmt = _fmt(dy)
        nil = _none(dy.device, dy.dtype)
        acc = torch.zeros([dy.shape[dim]], dtype=torch.float32, device=dy.device) if fuse_db else None
        dx = _plugin.bias_act(dy, b, x, y, nil, 1, dim, spec.cuda_idx, alpha, gain, clamp, db_accum=acc)
        ctx.save_for_backward(dy if spec.has_2nd_grad else nil, x, b, y)
        ctx.cfg = (dim, act, alpha, gain, clamp)
        if acc is None:
            acc = nil
        ctx.mark_non_differentiable(acc)
        return dx, acc

    @staticmethod
    def backward(ctx, d_dx, _d_acc):
        dim, act, alpha, gain, clamp = ctx.cfg
        spec = activation_funcs[act]
        d_dx = d_dx.contiguous(memory_format=ctx.fmt)
        dy, x, b, y = ctx.saved_tensors
        nil = _none(d_dx.device, d_dx.dtype)
        d_dy = d_x = d_b = None
        if ctx.needs_input_grad[0]:
            d_dy, _ = _BiasActGrad.apply(d_dx, x, b, y, dim, act, alpha, gain, clamp, False)
        if spec.has_2nd_grad and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]):
            d_x = _plugin.bias_act(d_dx, b, x, y, dy, 2, dim, spec.cuda_idx, alpha, gain, clamp)
        if spec.has_2nd_grad and ctx.needs_input_grad[2]:
            d_b = d_x.sum([i for i in range(d_x.ndim) if i != dim])
        return d_dy, d_x, d_b, None, None, None, None, None, None, None


def _bias_act_ref(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None):
    """Stock-PyTorch formulation (CPU tensors / impl='ref'); mirrors bias_act.py:94-123."""
    assert isinstance(x, torch.Tensor)
    spec, alpha, gain, clamp = _resolve(act, alpha, gain, clamp)
    if b is not None:
        assert isinstance(b, torch.Tensor) and b.ndim == 1
        assert 0 <= dim < x.ndim and b.shape[0] == x.shape[dim]
        x = x + b.reshape([-1 if i == dim else 1 for i in range(x.ndim)])
    x = spec.func(x, alpha=alpha)
    if gain != 1:
        x = x * gain
    if clamp >= 0:
        x = x.clamp(-clamp, clamp)
    return x


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None, impl='cuda'):
    """y = clamp(act(x + b) * gain).  Any shape; `b` is a vector along `dim`."""
    assert isinstance(x, torch.Tensor)
    assert impl in ('ref', 'cuda')
    if impl == 'cuda' and x.device.type == 'cuda':
        _, alpha, gain, clamp = _resolve(act, alpha, gain, clamp)
        return _BiasActFwd.apply(x, b, dim, act, alpha, gain, clamp)
    return _bias_act_ref(x=x, b=b, dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp)
