"""Discriminator-side kernels (csrc/disc_ops.cu, stylegan_v_b200/dconv.py) against torch fp64 on the GPU: the streaming `fromrgb` layer, the
minibatch-std + concat + channel-padding kernel."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err
from stylegan_v_b200 import dconv
from stylegan_v_b200.networks import MinibatchStdLayer

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('N,J,O,H,act', [(6, 3, 64, 33, 'lrelu'), (2, 3, 128, 16, 'lrelu'), (3, 1, 32, 20, 'linear'), (48, 3, 64, 64, 'lrelu')])
def test_fromrgb_forward_backward(N, J, O, H, act):
    g = torch.Generator().manual_seed(N + O)
    img = torch.randn(N, J, H, H, generator=g).cuda().requires_grad_(True)
    w = torch.randn(O, J, 1, 1, generator=g).cuda().requires_grad_(True)
    b = torch.randn(O, generator=g).cuda().requires_grad_(True)
    wg, gain = 1 / np.sqrt(J), float(np.sqrt(2)) if act == 'lrelu' else 1.0
    assert dconv.fromrgb_supported(img, w)
    y = dconv.fromrgb(img, w, b, act=act, gain=gain, weight_gain=wg)
    assert y.shape == (N, O, H, H) and y.stride(1) == 1                              # NHWC out of NCHW frames
    dy = torch.randn(N, O, H, H, generator=g).cuda()
    got = torch.autograd.grad(y, [img, w, b], dy)
    i64, w64, b64 = (t.detach().double().requires_grad_(True) for t in (img, w, b))
    r = F.conv2d(i64, w64 * wg) + b64[None, :, None, None]
    r = (F.leaky_relu(r, 0.2) if act == 'lrelu' else r) * gain
    ref = torch.autograd.grad(r, [i64, w64, b64], dy.double())
    assert rel_err(y, r) < 1e-6
    for a, e, n in zip(got, ref, ('dimg', 'dw', 'db')):
        assert rel_err(a, e) < 2e-5, (n, rel_err(a, e))


@pytest.mark.parametrize('N,C,G,nhwc', [(8, 512, 4, False), (6, 64, 2, True), (4, 128, 4, True), (3, 32, 8, False)])
def test_minibatch_std_concat_kernel(N, C, G, nhwc):
    g = torch.Generator().manual_seed(N * C)
    x = torch.randn(N, C, 4, 4, generator=g).cuda()
    if nhwc:
        x = x.contiguous(memory_format=torch.channels_last)
    x.requires_grad_(True)
    y = dconv.minibatch_std_concat(x, G, 1)
    cpad = (C + 1 + 63) // 64 * 64
    assert y.shape == (N, cpad, 4, 4) and y.stride(1) == 1
    x64 = x.detach().double().requires_grad_(True)
    ref = MinibatchStdLayer(G, 1)(x64)                                              # the reference arithmetic (networks.py:499-514) in fp64
    assert rel_err(y[:, :C + 1], ref) < 1e-6 and not y[:, C + 1:].any()
    dy = torch.randn(y.shape, generator=g).cuda()
    gx, = torch.autograd.grad(y, x, dy)
    rx, = torch.autograd.grad(ref, x64, dy[:, :C + 1].double())
    assert rel_err(gx, rx) < 1e-5
