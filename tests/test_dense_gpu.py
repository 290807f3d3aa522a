"""FullyConnectedLayer on the tcgen05 contraction kernels (stylegan_v_b200/dense.py) against fp64 torch on the GPU: forward and all
gradients (the kernels run these layers in tf32x3 arithmetic whatever the global precision mode is), and the module-level switches:
first-order callers get the kernel route, any-order callers the reference's addmm formulation.
Bars by contraction length K: the tensor core's fp32 accumulation is not round-to-nearest, so the tf32x3 error grows with K — measured
7e-6 at K = 512, 4.7e-5 at K = 8192 (call B of round 2); 2e-5 / 1e-4 asserted."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err
from stylegan_v_b200 import _lib, dense
from stylegan_v_b200.networks import FullyConnectedLayer, MappingNetwork

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('M,K,O,act,bias', [(32, 512, 512, 'linear', True), (48, 512, 1536, 'linear', True), (7, 64, 64, 'lrelu', True),
                                            (96, 8192, 512, 'lrelu', True), (33, 512, 1024, 'linear', False), (2, 512, 32, 'linear', True)])
def test_linear_vs_fp64(M, K, O, act, bias):
    g = torch.Generator().manual_seed(M + K)
    x = torch.randn(M, K, generator=g).cuda().requires_grad_(True)
    w = torch.randn(O, K, generator=g).cuda().requires_grad_(True)
    b = torch.randn(O, generator=g).cuda().requires_grad_(True) if bias else None
    wg, bg, gain = 1 / np.sqrt(K), 0.7, float(np.sqrt(2)) if act == 'lrelu' else 1.0
    assert dense.supported(x, w, act)
    n0 = _lib.launch_count()
    y = dense.linear(x, w, b, wg, bg, act=act, gain=gain)
    dy = torch.randn(M, O, generator=g).cuda()
    got = torch.autograd.grad(y, [x, w] + ([b] if bias else []), dy)
    assert _lib.launch_count() - n0 >= 5            # weight pass + contraction forward, two weight passes + two contractions backward
    xd, wd = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    bd = b.detach().double().requires_grad_(True) if bias else None
    r = xd @ (wd * wg).t() + (bd * bg if bias else 0)
    r = (F.leaky_relu(r, 0.2) if act == 'lrelu' else r) * gain
    ref = torch.autograd.grad(r, [xd, wd] + ([bd] if bias else []), dy.double())
    bar = 2e-5 if K <= 1024 else 1e-4
    assert rel_err(y, r) < bar
    for a, e, n in zip(got, ref, ('dx', 'dw', 'db')):
        assert rel_err(a, e) < bar, (n, rel_err(a, e))


def test_module_routes():
    """First-order callers take the kernels; the any-order formulation (default) stays on torch ops and is twice differentiable."""
    fc = FullyConnectedLayer(512, 512, activation='lrelu', lr_multiplier=0.01).cuda()
    x = torch.randn(16, 512, device='cuda', requires_grad=True)
    n0 = _lib.launch_count()
    a = fc(x)                                        # any-order route: no contraction launch of ours (bias_act kernel only)
    n1 = _lib.launch_count()
    b = fc(x, fused=True)
    assert _lib.launch_count() - n1 >= 2 and n1 - n0 <= 1
    assert rel_err(b, a) < 1e-5
    gx, = torch.autograd.grad(a.square().sum(), x, create_graph=True)
    gx.sum().backward()                              # second order works on the default route
    assert fc.weight.grad is not None
    mp = MappingNetwork(z_dim=512, c_dim=0, w_dim=512, num_ws=14, num_layers=2).cuda().eval()
    n2 = _lib.launch_count()
    ws = mp(torch.randn(8, 512, device='cuda'), torch.zeros(8, 0, device='cuda'))
    assert ws.shape == (8, 14, 512) and _lib.launch_count() - n2 >= 4
