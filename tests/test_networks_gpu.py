"""The native Generator / Discriminator modules on CUDA (drop-in ops -> libsgv_b200 kernels) against the reference goldens.
Contractions run TF32 products / fp32 accumulation on the tcgen05 kernels for channel counts % 32, so the bars are
the north_star's 1e-3-class tolerance for values and a looser one for gradients of these tiny, leaky-ReLU networks (a pre-activation
within TF32 round-off of zero flips its slope)."""
import pytest
import torch

from conftest import load_golden, rel_err
from stylegan_v_b200 import _lib
from stylegan_v_b200.ops import conv2d_gradfix
from test_networks_cpu import _t, discriminator_checks, make_discriminator, make_synthesis, path_length_checks

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _gradfix_enabled(monkeypatch):
    monkeypatch.setattr(conv2d_gradfix, 'enabled', True)          # what the reference's training loop does (training_loop.py:143)


def test_discriminator_cuda_vs_reference_golden(cuda):
    g, meta = load_golden('discriminator_tiny.npz')
    D = make_discriminator(g, meta).to(cuda)
    n0 = _lib.launch_count()
    discriminator_checks(D, g, cuda, 5e-3, 5e-2)
    assert _lib.launch_count() > n0


def test_synthesis_unfused_cuda_vs_reference_golden(cuda):
    g, meta = load_golden('synthesis_tiny.npz')
    net, _ = make_synthesis(g, meta)
    net = net.to(cuda).train()
    ws, t, mz = _t(g['ws']).to(cuda), _t(g['t']).to(cuda), _t(g['motion_z']).to(cuda)
    img_unfused = net(ws, t, motion_z=mz, unfused=True)
    img_fused = net(ws, t, motion_z=mz)
    assert rel_err(img_unfused, _t(g['img_train'])) < 3e-3
    assert rel_err(img_fused, _t(g['img_train'])) < 3e-3
    assert rel_err(img_fused, img_unfused) < 3e-3


def test_path_length_cuda_vs_reference_golden(cuda):
    g, meta = load_golden('path_length_tiny.npz')
    net, _ = make_synthesis(g, meta)
    path_length_checks(net.to(cuda), g, cuda, 2e-2)
