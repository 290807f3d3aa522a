"""Training phases on CUDA: the loss terms against the reference goldens (TF32 contractions => looser bars than on CPU), and the
flat-state plumbing of TrainingPhases (one fused update launch per phase, EMA, lazy-regularisation schedule)."""
import pytest
import torch

from conftest import load_golden, rel_err
from stylegan_v_b200 import _lib
from stylegan_v_b200 import train_step as ts
from stylegan_v_b200.ops import conv2d_gradfix
from test_train_step_cpu import _t, check_phase, make_gd, run_phase

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _gradfix_enabled(monkeypatch):
    monkeypatch.setattr(conv2d_gradfix, 'enabled', True)


def _main_phase_grads(phase, G, D, g, dev, mz):
    """Gmain / Dmain gradients with the motion noise fed explicitly (the loss functions forward synthesis kwargs to G), so that the result
    does not depend on which device's generator would have drawn it."""
    real = _t(g['real']).to(dev)
    real = real.view(-1, *real.shape[2:])
    real_t, gen_t, z = _t(g['real_t']).to(dev), _t(g['gen_t']).to(dev), _t(g['z']).to(dev)
    c = torch.zeros(len(z), 0, device=dev)
    module = G if phase == 'Gmain' else D
    G.requires_grad_(module is G)
    D.requires_grad_(module is D)
    for p in module.parameters():
        p.grad = None
    if phase == 'Gmain':
        loss = ts.generator_main_loss(G, D, z, c, gen_t, motion_z=mz.to(dev))
    else:
        a, b = ts.discriminator_main_loss(G, D, real, c, real_t, z, c, gen_t, motion_z=mz.to(dev))
        loss = a + b
    loss.backward()
    return float(loss), {n: p.grad.detach().double().cpu() for n, p in module.named_parameters() if p.grad is not None}


@pytest.mark.parametrize('phase', ['Gmain', 'Dmain'])
def test_main_phase_gradients_cuda_vs_cpu_evaluation(cuda, phase):
    """The CPU evaluation of the native modules is pinned to the reference's StyleGAN2Loss (tests/test_train_step_cpu.py, 1e-4); the CUDA
    evaluation (fused synthesis layers, fused discriminator nodes, tcgen05 contractions) of the SAME modules and inputs must agree with it:
    fp32-grade in tf32x3 mode (every gradient <= 2e-3 normwise), direction + norm in the default TF32 mode."""
    import torch.nn.functional as F
    from stylegan_v_b200 import precision
    g, meta = load_golden('loss_phases_tiny.npz')
    G, D = make_gd(g, meta)
    mz = torch.randn(len(_t(g['z'])), G.synthesis.motion_encoder.traj_len(), G.synthesis.motion_encoder.z_dim, generator=torch.Generator().manual_seed(5))
    loss_c, ref = _main_phase_grads(phase, G, D, g, torch.device('cpu'), mz)
    assert len(ref) > 10
    G, D = make_gd(g, meta)
    G, D = G.to(cuda), D.to(cuda)
    n0 = _lib.launch_count()
    with precision.precision('tf32x3'):
        loss_g, got = _main_phase_grads(phase, G, D, g, cuda, mz)
    assert _lib.launch_count() > n0
    assert abs(loss_g - loss_c) < 1e-4 * max(1.0, abs(loss_c))
    assert set(got) == set(ref)
    worst = max((rel_err(got[n], ref[n]), n) for n in ref)
    assert worst[0] < 1e-2, worst              # measured 2.5e-3 ... 8e-3 (affine weights / biases of this tiny network: few-hundred-element sums)
    loss_t, got = _main_phase_grads(phase, G, D, g, cuda, mz)            # default TF32 mode
    assert abs(loss_t - loss_c) < 5e-3 * max(1.0, abs(loss_c))
    for n in ref:
        if ref[n].ndim >= 2 and float(ref[n].norm()) > 0:
            cos = float(F.cosine_similarity(got[n].flatten(), ref[n].flatten(), dim=0))
            assert cos > 0.99 and abs(float(got[n].norm() / ref[n].norm()) - 1) < 5e-2, (n, cos)


def test_dreg_gradients_cuda_vs_reference_loss(cuda):
    g, meta = load_golden('loss_phases_tiny.npz')
    G, D = make_gd(g, meta)
    G, D = G.to(cuda), D.to(cuda)
    module = run_phase('Dreg', G, D, g, cuda, meta['r1_gamma'])
    check_phase('Dreg', module, g, 5e-2, 1e-1, weights_only=True)


@pytest.mark.parametrize('fused_d', [False, True])
def test_training_phases_step(cuda, fused_d, monkeypatch):
    monkeypatch.setattr(ts, 'FUSED_DISCRIMINATOR', fused_d)
    g, meta = load_golden('loss_phases_tiny.npz')
    G, D = make_gd(g, meta)
    G, D = G.to(cuda), D.to(cuda)
    tp = ts.TrainingPhases(G, D, lr=0.0025, r1_gamma=0.5, pl_weight=2.0, G_reg_interval=4, D_reg_interval=16, batch_size=2, ema_kimg=0.01)
    assert all(p.data_ptr() % 256 == 0 for p in G.parameters())
    real = _t(g['real']).to(cuda)
    real = real.view(-1, *real.shape[2:])
    real_t, gen_t, z = _t(g['real_t']).to(cuda), _t(g['gen_t']).to(cuda), _t(g['z']).to(cuda)
    p0 = tp.G_state.param.clone()
    d0 = tp.D_state.param.clone()
    e0 = tp.G_state.ema.clone()
    out = tp.step(real, real_t, z, gen_t)                         # iteration 0: all four phases are due
    assert set(out) == {'Gmain', 'Greg', 'Dmain', 'Dreg'}
    assert all(torch.isfinite(v).all() for v in out.values())
    assert not torch.equal(tp.G_state.param, p0) and not torch.equal(tp.D_state.param, d0) and not torch.equal(tp.G_state.ema, e0)
    assert not tp.G_state.grad.any() and not tp.D_state.grad.any()                   # re-zeroed by the update kernel
    assert tp.G_opt.t == 2 and tp.D_opt.t == 2
    # EMA moved towards the new parameters by (1 - beta)
    beta = tp.ema_beta()
    want = tp.G_state.param + (e0 - tp.G_state.param) * beta
    assert rel_err(tp.G_state.ema, want) < 1e-4
    out = tp.step(real, real_t, z, gen_t)                         # iteration 1: main phases only
    assert set(out) == {'Gmain', 'Dmain'}
    assert torch.isfinite(tp.G_state.param).all() and torch.isfinite(tp.D_state.param).all()
    # the EMA generator is a usable module on its flat storage
    with torch.no_grad():
        img = tp.G_ema(z, torch.zeros(2, 0, device=cuda), gen_t)
    assert img.shape == (6, 3, 32, 32) and torch.isfinite(img).all()
