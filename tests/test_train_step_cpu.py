"""Loss terms of the training phases (stylegan_v_b200/train_step.py) against the reference's StyleGAN2Loss.accumulate_gradients
(loss.py:73-173) run on the unmodified reference G / D (goldens: oracle/make_goldens.py::gen_loss_phases): per-parameter gradient
sums / norms for Gmain, Dmain and Dreg (R1, gain 16), the RNG stream of the motion noise included."""
import json

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from stylegan_v_b200 import train_step as ts
from stylegan_v_b200.networks import Discriminator, Generator


def _t(a):
    return torch.from_numpy(np.asarray(a))


def make_gd(g, meta):
    mg, md = meta['G'], meta['D']
    G = Generator(z_dim=mg['w_dim'], w_dim=mg['w_dim'], img_resolution=mg['img_resolution'], channel_base=mg['channel_base'], channel_max=mg['channel_max'],
                  motion_z_dim=mg['motion_z_dim'], motion_v_dim=mg['motion_v_dim'], time_enc_dim=mg['time_enc_dim'], mapping_layers=2)
    D = Discriminator(c_dim=0, img_resolution=md['img_resolution'], channel_base=md['channel_base'], channel_max=md['channel_max'],
                      num_frames_per_video=md['num_frames_per_video'], max_num_frames=md['max_num_frames'], concat_res=md['concat_res'],
                      num_frames_div_factor=md['num_frames_div_factor'], mbstd_group_size=md['mbstd_group_size'], mapping_layers=md['mapping_layers'])
    gsd = {k[2:]: _t(g[k]) for k in g.files if k.startswith('g:')}
    dsd = {k[2:]: _t(g[k]) for k in g.files if k.startswith('d:')}
    assert set(gsd) == set(G.state_dict()) and set(dsd) == set(D.state_dict())
    G.load_state_dict(gsd)
    D.load_state_dict(dsd)
    return G.train(), D.train()


def run_phase(phase, G, D, g, dev, r1_gamma):
    real = _t(g['real']).to(dev)
    real = real.view(-1, *real.shape[2:])
    real_t, gen_t, z = _t(g['real_t']).to(dev), _t(g['gen_t']).to(dev), _t(g['z']).to(dev)
    c = torch.zeros(len(z), 0, device=dev)
    module = G if phase == 'Gmain' else D
    G.requires_grad_(module is G)
    D.requires_grad_(module is D)
    for p in module.parameters():
        p.grad = None
    torch.manual_seed(100)
    if phase == 'Gmain':
        ts.generator_main_loss(G, D, z, c, gen_t).backward()
    elif phase == 'Dmain':
        a, b = ts.discriminator_main_loss(G, D, real, c, real_t, z, c, gen_t)
        (a + b).backward()
    else:
        ts.discriminator_r1_loss(D, real, c, real_t, r1_gamma).mul(16).backward()
    return module


def check_phase(phase, module, g, tol_stat, tol_full, weights_only=False):
    """weights_only (TF32 hardware): bias / embedding gradients of the R1 term are ill-conditioned in this tiny network (see
    test_networks_cpu.discriminator_checks) — compare the norms of the weight gradients and the stored full tensors only."""
    stats = json.loads(bytes(g['stats:' + phase]).decode())
    P = dict(module.named_parameters())
    assert set(stats) == {n for n, p in P.items() if p.grad is not None}
    for n, (s, nrm) in stats.items():
        gr = P[n].grad.double()
        if weights_only:
            if n.endswith('.weight') and 'const_embed' not in n:
                assert abs(float(gr.norm()) - nrm) <= tol_stat * max(nrm, 1e-12), (phase, n, float(gr.norm()), nrm)
            continue
        assert abs(float(gr.norm()) - nrm) <= tol_stat * max(nrm, 1e-12), (phase, n, float(gr.norm()), nrm)
        assert abs(float(gr.sum()) - s) <= tol_stat * max(nrm, 1e-12) * np.sqrt(gr.numel()), (phase, n)
    for k in g.files:
        if k.startswith(f'grad:{phase}:'):
            assert rel_err(P[k.split(':', 2)[2]].grad, _t(g[k])) < tol_full, k


@pytest.mark.parametrize('phase', ['Gmain', 'Dmain', 'Dreg'])
def test_phase_gradients_vs_reference_loss(phase):
    g, meta = load_golden('loss_phases_tiny.npz')
    G, D = make_gd(g, meta)
    module = run_phase(phase, G, D, g, torch.device('cpu'), meta['r1_gamma'])
    check_phase(phase, module, g, 1e-4, 1e-4)


def test_w_avg_tracks_both_generator_calls():
    """Gmain and Dmain each run the mapping network in training mode (loss.py:86,124), i.e. two moving-average updates (layers.py:86-88)."""
    g, meta = load_golden('loss_phases_tiny.npz')
    G, D = make_gd(g, meta)
    run_phase('Gmain', G, D, g, torch.device('cpu'), meta['r1_gamma'])
    run_phase('Dmain', G, D, g, torch.device('cpu'), meta['r1_gamma'])
    assert rel_err(G.mapping.w_avg, _t(g['w_avg_after'])) < 1e-5


def test_dmain_with_video_consistent_augmentation_vs_reference_loss():
    """Dmain with the ADA pipe in front of D (loss.py:58-70, video_consistent_aug): same generator stream (motion noise, pipe draws for
    the generated clips, pipe draws for the real clips) and same discriminator gradients as the reference loss."""
    from stylegan_v_b200.augment import AugmentPipe
    g, meta = load_golden('loss_phases_tiny.npz')
    G, D = make_gd(g, meta)
    pipe = AugmentPipe(**meta['aug'])
    pipe.p.copy_(torch.as_tensor(meta['aug_p']))
    real = _t(g['real'])
    real = real.view(-1, *real.shape[2:])
    real_t, gen_t, z = _t(g['real_t']), _t(g['gen_t']), _t(g['z'])
    c = torch.zeros(len(z), 0)
    G.requires_grad_(False)
    D.requires_grad_(True)
    torch.manual_seed(100)
    a, b = ts.discriminator_main_loss(G, D, real, c, real_t, z, c, gen_t, augment_pipe=pipe, video_consistent_aug=True)
    (a + b).backward()
    check_phase('Dmain_aug', D, g, 1e-4, 1e-4)
