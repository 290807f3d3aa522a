"""CUDA upfirdn2d (libsgv_b200 through the C ABI) vs the oracle.  fp32 results must be BIT-EXACT: the index
arithmetic is integer work and the tap accumulation order + FMA contraction equal the reference kernel's."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import ops_ref
from stylegan_v_b200 import plugin
from stylegan_v_b200.ops import upfirdn2d as U

pytestmark = pytest.mark.gpu


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _run_cuda(x, f, **kw):
    return U.upfirdn2d(x.cuda(), None if f is None else f.cuda(), **kw)


def test_golden_cases_fp64_and_fp32_bitexact():
    g, meta = load_golden('upfirdn2d_cases.npz')
    for i, m in enumerate(meta):
        x = _t(g[f'c{i}_x'])
        if m.get('channels_last'):
            x = x.contiguous(memory_format=torch.channels_last)
        f = _t(g[f'c{i}_f']) if m['has_f'] else None
        kw = dict(up=m['up'], down=m['down'], padding=m['padding'], flip_filter=m['flip'], gain=m['gain'])
        # fp64 vs the reference-generated golden
        y = _run_cuda(x, f, **kw)
        assert y.shape == g[f'c{i}_y'].shape
        assert rel_err(y, _t(g[f'c{i}_y'])) < 2e-7, (i, m)
        # fp32: bit-exact vs the scalar C port of the reference kernel
        x32 = x.float()
        y32 = _run_cuda(x32, f, **kw).cpu()
        o32 = ops_ref.upfirdn2d_ref(x32, f, **kw)
        assert torch.equal(y32, o32), (i, m, (y32 - o32).abs().max())
        # gradient (another pass, upfirdn2d.py:246-261)
        xg = x.cuda().requires_grad_(True)
        yg = U.upfirdn2d(xg, None if f is None else f.cuda(), **kw)
        dx, = torch.autograd.grad(yg, xg, _t(g[f'c{i}_dy']).cuda())
        assert rel_err(dx, _t(g[f'c{i}_dx'])) < 2e-7, (i, m)
        # fp16 goes through the generic kernel with fp32 accumulation
        y16 = _run_cuda(x.half(), f, **kw)
        assert rel_err(y16.float(), _t(g[f'c{i}_y'])) < 5e-3, (i, m)


SHAPES = [
    # (N, C, H, W, up, down, padding, flip, gain)  — hot-path geometries at reduced batch
    (2, 16, 65, 65, 1, 1, [1, 1, 1, 1], False, 4),       # G up-layer FIR: 2h+1 -> 2h (odd row pitch)
    (2, 8, 129, 129, 1, 1, [1, 1, 1, 1], False, 4),
    (1, 4, 257, 257, 1, 1, [1, 1, 1, 1], False, 4),      # the b256 geometry
    (2, 16, 64, 64, 1, 1, [2, 2, 2, 2], True, 4),        # its backward: 2h -> 2h+1
    (2, 3, 32, 32, 2, 1, [2, 1, 2, 1], False, 4),        # img upsample2d
    (2, 3, 64, 64, 1, 2, [2, 1, 2, 1], True, 4),         # backward of the img upsample
    (2, 8, 64, 64, 1, 2, [1, 1, 1, 1], False, 1),        # D skip: down=2
    (2, 8, 64, 64, 1, 1, [2, 2, 2, 2], False, 1),        # D blur before the stride-2 conv
    (1, 2, 5, 3, 1, 1, [1, 1, 1, 1], False, 1),          # tiny
    (1, 3, 33, 47, 2, 2, [3, 1, 0, 2], False, 2),        # up and down together, ragged
    (1, 2, 40, 40, 3, 1, [2, 2, 2, 2], False, 9),        # factor not covered by the tiled kernel -> generic
    (1, 2, 31, 29, [1, 2], [2, 1], [0, 1, 2, 0], True, 1),
    (2, 8, 128, 128, 1, 1, [2, 2, 2, 2], True, 4),       # backward of the b128 FIR: 128 -> 129 (ragged last tile), pipelined kernel
    (1, 3, 200, 150, 1, 1, [1, 1, 1, 1], False, 4),      # wide, not a multiple of the tile
    (1, 2, 100, 517, 1, 1, [2, 1, 1, 2], False, 1),      # 5 tiles across, asymmetric padding
    (2, 64, 65, 65, 1, 1, [1, 1, 1, 1], False, 4),       # channels_last + C % 32 == 0: the TMA-fed persistent kernel, 2 channel blocks
    (3, 32, 40, 52, 1, 1, [2, 2, 2, 2], True, 4),        # its backward geometry, ragged tiles in both directions
    (1, 96, 33, 100, 1, 1, [2, 1, 1, 2], False, 1),      # 3 channel blocks, asymmetric padding
    (2, 8, 32, 32, 2, 1, [2, 1, 2, 1], True, 1),         # backward of the D skip FIR: zero insertion x2, channels_last quad kernel
    (1, 12, 17, 9, 2, 1, [2, 1, 2, 1], False, 4),        # same kernel, odd extents
    (1, 4, 1, 1, 2, 1, [2, 1, 2, 1], True, 1),           # 1 x 1 input: every tap row / column clipped somewhere
    (1, 4, 6, 5, 2, 1, [4, 0, 0, 3], False, 1),          # even leading pads 4 / 0, odd output extents
    (1, 4, 6, 5, 2, 1, [3, 2, 1, 2], False, 1),          # odd leading pads: stays on the general channels_last kernel
]


@pytest.mark.parametrize('channels_last', [False, True])
def test_shapes_bitexact_vs_oracle(channels_last):
    f = U.setup_filter([1, 3, 3, 1])
    for i, (N, C, H, W, up, down, pad, flip, gain) in enumerate(SHAPES):
        gen = torch.Generator().manual_seed(i)
        x = torch.randn(N, C, H, W, generator=gen)
        if channels_last:
            x = x.contiguous(memory_format=torch.channels_last)
        kw = dict(up=up, down=down, padding=pad, flip_filter=flip, gain=gain)
        y = _run_cuda(x, f, **kw)
        o = ops_ref.upfirdn2d_ref(x, f, **kw)
        assert y.shape == o.shape
        if channels_last and C > 1 and H * W > 1:          # (a 1 x 1 plane is NCHW- and NHWC-contiguous at once: torch reports NCHW)
            assert y.is_contiguous(memory_format=torch.channels_last)
        assert torch.equal(y.cpu(), o), (i, channels_last, (y.cpu() - o).abs().max().item())


def test_other_filters_and_separable():
    for taps in ([1, 2, 1], [1, 1], [1, 4, 6, 4, 1], list(range(1, 13))):
        f = U.setup_filter(taps)
        x = torch.randn(2, 3, 37, 41, generator=torch.Generator().manual_seed(len(taps)))
        for kw in (dict(up=2, padding=[3, 2, 3, 2], gain=4), dict(down=2, padding=[2, 2, 2, 2]), dict(padding=len(taps) // 2)):
            y = _run_cuda(x, f, **kw).cpu()
            o = ops_ref.upfirdn2d_ref(x, f, **kw)
            assert torch.equal(y, o), (taps, kw)
    # asymmetric 2-D filter, both flips
    f = torch.arange(1, 13, dtype=torch.float32).reshape(3, 4) / 78
    x = torch.randn(1, 2, 20, 23)
    for flip in (False, True):
        assert torch.equal(_run_cuda(x, f, padding=[2, 1, 1, 1], flip_filter=flip).cpu(), ops_ref.upfirdn2d_ref(x, f, padding=[2, 1, 1, 1], flip_filter=flip))


def test_strided_views_and_unaligned_pointers():
    f = U.setup_filter([1, 3, 3, 1])
    base = torch.randn(2, 6, 40, 43)
    for x in (base[:, 1:5], base[:, :, 3:, 2:], base.transpose(2, 3), base.flatten()[1:1 + 2 * 5 * 40 * 43].reshape(2, 5, 40, 43)):
        y = _run_cuda(x, f, padding=1, gain=4).cpu()
        assert torch.equal(y, ops_ref.upfirdn2d_ref(x.contiguous(), f, padding=1, gain=4))


def test_full_size_properties():
    """BASELINE config-2 geometry at full width, checked through size-independent properties."""
    dev = torch.device('cuda')
    f = U.setup_filter([1, 3, 3, 1], device=dev)
    N, C, H = 4, 64, 257
    x = torch.randn(N, C, H, H, device=dev)
    y = U.upfirdn2d(x, f, padding=1, gain=4)
    assert y.shape == (N, C, 256, 256)
    # DC gain: constant interior -> gain * sum(f) = 4
    ones = torch.ones(1, 1, H, H, device=dev)
    yo = U.upfirdn2d(ones, f, padding=1, gain=4)
    assert torch.allclose(yo[:, :, 2:-2, 2:-2], torch.full_like(yo[:, :, 2:-2, 2:-2], 4.0), atol=1e-6)
    # linearity
    x2 = torch.randn_like(x)
    y2 = U.upfirdn2d(x2, f, padding=1, gain=4)
    ys = U.upfirdn2d(x + 2 * x2, f, padding=1, gain=4)
    assert rel_err(ys, y + 2 * y2) < 1e-5
    # a random sub-block against the oracle, bit-exact
    sub = x[1:2, 5:9].cpu()
    assert torch.equal(y[1:2, 5:9].cpu(), ops_ref.upfirdn2d_ref(sub, f.cpu(), padding=1, gain=4))
    # adjointness <A x, z> == <x, A^T z>  (backward pass is the adjoint)
    xg = x.clone().requires_grad_(True)
    z = torch.randn_like(y)
    (U.upfirdn2d(xg, f, padding=1, gain=4) * z).sum().backward()
    lhs = (y.double() * z.double()).sum()
    rhs = (x.double() * xg.grad.double()).sum()
    assert abs(lhs - rhs) / abs(lhs) < 1e-5


def test_fused_epilogue_matches_unfused_sequence():
    dev = torch.device('cuda')
    f = U.setup_filter([1, 3, 3, 1], device=dev)
    from stylegan_v_b200.ops import bias_act as B
    for cl, C, H in ((False, 8, 33), (True, 8, 33), (True, 64, 47)):     # the last one runs on the TMA-fed channels_last kernel
        x = torch.randn(3, C, H, H, device=dev)
        if cl:
            x = x.contiguous(memory_format=torch.channels_last)
        scale = torch.rand(3, C, device=dev) + 0.5
        bias = torch.randn(C, device=dev)
        y = plugin.upfirdn2d(x, f, 1, 1, 1, 1, 1, 1, 1, 1, False, 4.0,
                             epilogue=dict(scale=scale, bias=bias, act='lrelu', alpha=0.2, gain=np.sqrt(2), clamp=None))
        ref = U.upfirdn2d(x, f, padding=1, gain=4) * scale[:, :, None, None]
        ref = B.bias_act(ref, bias, act='lrelu')
        assert torch.equal(y, ref)


def test_error_behaviour():
    dev = torch.device('cuda')
    x = torch.randn(1, 1, 4, 4, device=dev)
    with pytest.raises(RuntimeError):
        plugin.upfirdn2d(x, torch.ones(8, 8, device=dev), 1, 1, 1, 1, 0, 0, 0, 0, False, 1.0)   # output < 1x1
    with pytest.raises(RuntimeError):
        plugin.upfirdn2d(x, torch.ones(2, 2, device=dev, dtype=torch.float64), 1, 1, 1, 1, 0, 0, 0, 0, False, 1.0)
    with pytest.raises(AssertionError):
        U.upfirdn2d(x, None, up=0)
