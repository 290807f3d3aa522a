"""Drop-in for `src.torch_utils.ops.fma` (reference: src/torch_utils/ops/fma.py): fma(a, b, c) = a * b + c with
broadcasting, and a backward that reduces each gradient straight to its operand's shape (fma.py:29-58)."""
import torch


def fma(a, b, c):
    return _MulAdd.apply(a, b, c)


def _reduce_to(g, shape):
    """Sums a broadcast gradient back to `shape` (the inverse of broadcasting)."""
    lead = g.ndim - len(shape)
    assert lead >= 0
    dims = [i for i in range(g.ndim) if g.shape[i] > 1 and (i < lead or shape[i - lead] == 1)]
    if dims:
        g = g.sum(dim=dims, keepdim=True)
    if lead:
        g = g.reshape(-1, *g.shape[lead + 1:])
    assert g.shape == shape
    return g


class _MulAdd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, c):
        ctx.save_for_backward(a, b)
        ctx.c_shape = c.shape
        return torch.addcmul(c, a, b)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        ga = _reduce_to(g * b, a.shape) if ctx.needs_input_grad[0] else None
        gb = _reduce_to(g * a, b.shape) if ctx.needs_input_grad[1] else None
        gc = _reduce_to(g, ctx.c_shape) if ctx.needs_input_grad[2] else None
        return ga, gb, gc
