"""CUDA bias_act (libsgv_b200 through the C ABI) vs the oracle (C port of bias_act.cu) and the goldens."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import ops_ref
from stylegan_v_b200.ops import bias_act as B

pytestmark = pytest.mark.gpu


def _t(a):
    return torch.from_numpy(np.asarray(a))


def test_golden_cases_all_orders():
    g, meta = load_golden('bias_act_cases.npz')
    for i, m in enumerate(meta):
        x = _t(g[f'c{i}_x']).cuda().requires_grad_(True)
        b = _t(g[f'c{i}_b']).cuda().requires_grad_(True) if m['use_b'] else None
        dy = _t(g[f'c{i}_dy']).cuda().requires_grad_(True)
        kw = dict(dim=m['dim'], act=m['act'], alpha=m['alpha'], gain=m['gain'], clamp=m['clamp'])
        y = B.bias_act(x, b, **kw)
        assert rel_err(y, _t(g[f'c{i}_y'])) < 2e-7, (i, m)
        ins = [x] + ([b] if m['use_b'] else [])
        grads = torch.autograd.grad(y, ins, dy, create_graph=True)
        if m['act'] == 'linear' and m['clamp'] is not None:
            continue    # reference CUDA quirk (no clamp mask for 'linear'), see tests/test_oracle_vs_golden.py
        assert rel_err(grads[0], _t(g[f'c{i}_dx'])) < 5e-7, (i, m)
        if m['use_b']:
            assert rel_err(grads[1], _t(g[f'c{i}_db'])) < 5e-7, (i, m)
        g2 = torch.autograd.grad(grads[0], [dy, x], _t(g[f'c{i}_ddx']).cuda(), allow_unused=True)
        assert rel_err(g2[0], _t(g[f'c{i}_g2_dy'])) < 5e-7, (i, m)
        ref_g2x = _t(g[f'c{i}_g2_x'])
        got = g2[1] if g2[1] is not None else torch.zeros_like(x)
        assert (got.cpu() - ref_g2x).abs().max() < 5e-7 * max(1.0, float(ref_g2x.abs().max())), (i, m)


@pytest.mark.parametrize('act', ['linear', 'lrelu', 'relu'])
def test_fp32_bitexact_vs_c_port(act):
    gen = torch.Generator().manual_seed(0)
    for shape, dim, cl in (([4, 16, 32, 32], 1, False), ([4, 16, 32, 32], 1, True), ([5, 7, 9, 3], 1, False), ([33, 512], 1, False), ([3, 5], 0, False), ([2, 6, 5, 5], 1, True)):
        x = torch.randn(shape, generator=gen)
        if cl:
            x = x.contiguous(memory_format=torch.channels_last)
        b = torch.randn(shape[dim], generator=gen)
        for gain, clamp in ((None, None), (0.5, 0.8)):
            kw = dict(dim=dim, act=act, gain=gain, clamp=clamp)
            y = B.bias_act(x.cuda(), b.cuda(), **kw).cpu()
            o = ops_ref.bias_act_kernel_ref(x, b, **kw)
            assert y.stride() == x.stride()
            assert torch.equal(y, o), (shape, act, gain, clamp)


def test_transcendental_acts_within_tolerance_fp32():
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(4, 8, 16, 16, generator=gen) * 3
    b = torch.randn(8, generator=gen)
    for act in ('tanh', 'sigmoid', 'elu', 'selu', 'softplus', 'swish'):
        y = B.bias_act(x.cuda(), b.cuda(), act=act).cpu()
        o = ops_ref.bias_act_kernel_ref(x.double(), b.double(), act=act)
        assert rel_err(y, o) < 1e-5, act    # tolerance: 1e-3 relative fp32 is the contract; we are far inside


def test_fused_bias_gradient_reduction():
    dev = torch.device('cuda')
    for shape, cl in (([8, 64, 32, 32], False), ([8, 64, 32, 32], True), ([6, 12, 7, 5], False), ([64, 512], False)):
        x = torch.randn(shape, device=dev)
        if cl:
            x = x.contiguous(memory_format=torch.channels_last)
        x.requires_grad_(True)
        b = torch.randn(shape[1], device=dev, requires_grad=True)
        y = B.bias_act(x, b, act='lrelu')
        dy = torch.randn_like(y)
        gx, gb = torch.autograd.grad(y, [x, b], dy)           # fused reduction path (no create_graph)
        gx2, gb2 = torch.autograd.grad(B.bias_act(x, b, act='lrelu'), [x, b], dy, create_graph=True)   # unfused path
        assert torch.equal(gx, gx2.detach())
        assert rel_err(gb, gb2) < 1e-5
        expect = gx.double().sum([i for i in range(x.ndim) if i != 1])
        assert rel_err(gb, expect) < 1e-5


def test_fp16_and_fp64_dispatch():
    x = torch.randn(2, 4, 8, 8)
    b = torch.randn(4)
    y64 = B.bias_act(x.double().cuda(), b.double().cuda(), act='lrelu').cpu()
    assert torch.equal(y64, ops_ref.bias_act_kernel_ref(x.double(), b.double(), act='lrelu'))
    y16 = B.bias_act(x.half().cuda(), b.half().cuda(), act='lrelu', clamp=256).cpu()
    assert rel_err(y16.float(), ops_ref.bias_act_kernel_ref(x.half().float(), b.half().float(), act='lrelu', clamp=256)) < 2e-3


def test_error_behaviour():
    from stylegan_v_b200 import plugin
    dev = torch.device('cuda')
    x = torch.randn(2, 4, 3, 3, device=dev)
    e = torch.empty(0, device=dev)
    with pytest.raises(RuntimeError):
        plugin.bias_act(x, torch.zeros(5, device=dev), e, e, e, 0, 1, 3, 0.2, 1.0, -1.0)       # wrong bias length
    with pytest.raises(RuntimeError):
        plugin.bias_act(x, e, e, e, e, 0, 1, 42, 0.2, 1.0, -1.0)                              # unknown activation
    with pytest.raises(RuntimeError):
        plugin.bias_act(x[:, :, ::2], e, e, e, e, 0, 1, 3, 0.2, 1.0, -1.0)                     # not dense
    assert B.bias_act(torch.empty(0, 4, device=dev), torch.zeros(4, device=dev)).shape == (0, 4)
