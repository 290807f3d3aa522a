"""FullyConnectedLayer / EqLRConv1d on the tcgen05 contraction kernels (stylegan_v_b200/dense.py) against fp64 torch on the GPU: forward
and all gradients, fp32-grade bars (the kernels run these layers in tf32x3 arithmetic whatever the global precision mode is), and the
module-level switches: first-order callers get the kernel route, any-order callers the reference's addmm formulation."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err
from stylegan_v_b200 import _lib, dense
from stylegan_v_b200.networks import FullyConnectedLayer, MappingNetwork
from stylegan_v_b200.time_encoder import EqualizedLinear, MotionMappingNetwork

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
    assert rel_err(y, r) < 1e-5
    for a, e, n in zip(got, ref, ('dx', 'dw', 'db')):
        assert rel_err(a, e) < 2e-5, (n, rel_err(a, e))


@pytest.mark.parametrize('B,L,Ci,Co,k', [(4, 86, 512, 512, 11), (3, 30, 32, 64, 5), (16, 86, 512, 512, 11)])
def test_conv1d_lines_vs_fp64(B, L, Ci, Co, k):
    g = torch.Generator().manual_seed(B + L)
    x = torch.randn(B, L, Ci, generator=g).cuda().requires_grad_(True)
    w = (torch.randn(Co, Ci, k, generator=g) / 0.01).cuda().requires_grad_(True)          # lr_multiplier 0.01 parametrisation (motion.py:55-58)
    b = torch.randn(Co, generator=g).cuda().requires_grad_(True)
    wg, bg = 0.01 / np.sqrt(Ci * k), 0.01
    y = dense.conv1d_lines(x, w, b, wg, bg, act='lrelu')
    assert y.shape == (B, L - k + 1, Co)
    dy = torch.randn(y.shape, generator=g).cuda()
    got = torch.autograd.grad(y, [x, w, b], dy)
    xd, wd, bd = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    r = F.leaky_relu(F.conv1d(xd.permute(0, 2, 1), wd * wg, bd * bg), 0.2).permute(0, 2, 1)
    ref = torch.autograd.grad(r, [xd, wd, bd], dy.double())
    assert rel_err(y, r) < 1e-5
    for a, e, n in zip(got, ref, ('dx', 'dw', 'db')):
        assert rel_err(a, e) < 2e-5, (n, rel_err(a, e))


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
    aff = EqualizedLinear(512, 256, bias_init=1).cuda()
    assert rel_err(aff(x, fused=True), aff(x)) < 1e-5
    mp = MappingNetwork(z_dim=512, c_dim=0, w_dim=512, num_ws=14, num_layers=2).cuda().eval()
    n2 = _lib.launch_count()
    ws = mp(torch.randn(8, 512, device='cuda'), torch.zeros(8, 0, device='cuda'))
    assert ws.shape == (8, 14, 512) and _lib.launch_count() - n2 >= 4


def test_motion_encoder_on_kernels_vs_cpu():
    """The whole motion mapping network (two conv1d layers on the contraction kernel, stacked predictor GEMM, fused Fourier tail) on CUDA
    against its own CPU evaluation (pinned to the reference golden, tests/test_networks_cpu.py), forward and parameter gradients."""
    torch.manual_seed(0)
    enc = MotionMappingNetwork(z_dim=512, v_dim=512, time_enc_dim=256)
    t = torch.tensor([[0.0, 5.25, 9.0], [100.5, 101.0, 130.75], [640.0, 650.5, 700.0]])
    mz = torch.randn(3, enc.traj_len(), 512)
    dv = torch.randn(9, 512)
    v_c = enc(t, motion_z=mz)['motion_v']
    g_c = torch.autograd.grad(v_c, list(enc.parameters()), dv)
    enc = enc.cuda()
    n0 = _lib.launch_count()
    v_g = enc(t.cuda(), motion_z=mz.cuda())['motion_v']
    g_g = torch.autograd.grad(v_g, list(enc.parameters()), dv.cuda())
    assert _lib.launch_count() - n0 >= 12
    assert rel_err(v_g, v_c) < 2e-4                  # fp32 sin / cos of arguments up to ~800 rad: one ulp of the argument is ~5e-5
    for (n, _), a, e in zip(enc.named_parameters(), g_g, g_c):
        assert rel_err(a, e) < 2e-3, (n, rel_err(a, e))
