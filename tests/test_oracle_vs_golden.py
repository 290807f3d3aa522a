"""Pins the oracle (oracle/) against the golden vectors minted from the unmodified reference
(oracle/make_goldens.py).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import ops_ref, synthesis_ref as sr


def _t(a):
    return torch.from_numpy(np.asarray(a))


def test_upfirdn2d_c_port_matches_reference_ref_impl():
    g, meta = load_golden('upfirdn2d_cases.npz')
    for i, m in enumerate(meta):
        x = _t(g[f'c{i}_x'])
        if m.get('channels_last'):
            x = x.contiguous(memory_format=torch.channels_last)
        f = _t(g[f'c{i}_f']) if m['has_f'] else None
        kw = dict(up=m['up'], down=m['down'], padding=m['padding'], flip_filter=m['flip'], gain=m['gain'])
        y = ops_ref.upfirdn2d_ref(x, f, **kw)
        assert y.shape == g[f'c{i}_y'].shape, (i, m)
        # fp64 data, but the taps are fp32 and the reference folds `gain` into them in fp32 (upfirdn2d.py:192)
        assert rel_err(y, _t(g[f'c{i}_y'])) < 2e-7, (i, m)
        y32 = ops_ref.upfirdn2d_ref(x.float(), f, **kw)
        assert rel_err(y32, _t(g[f'c{i}_y'])) < 5e-6, (i, m)
        # gradient = another pass with swapped factors (upfirdn2d.py:246-261)
        dy = _t(g[f'c{i}_dy'])
        p = ops_ref.upfirdn2d_backward_padding(x.shape, dy.shape, f, m['up'], m['down'], m['padding'])
        dx = ops_ref.upfirdn2d_ref(dy, f, up=m['down'], down=m['up'], padding=p, flip_filter=not m['flip'], gain=m['gain'])
        assert rel_err(dx, _t(g[f'c{i}_dx'])) < 2e-7, (i, m)
        # torch-op restatement agrees too
        yt = ops_ref.upfirdn2d_ref_torch(x, f, **kw)
        assert rel_err(yt, _t(g[f'c{i}_y'])) < 1e-12, (i, m)


def test_bias_act_c_port_matches_reference_ref_impl():
    # fp64 data; alpha/gain/clamp cross the plugin ABI as float32 (bias_act.cpp:32), hence ~1e-7 not 1e-16
    g, meta = load_golden('bias_act_cases.npz')
    for i, m in enumerate(meta):
        x = _t(g[f'c{i}_x']); dy = _t(g[f'c{i}_dy']); ddx = _t(g[f'c{i}_ddx'])
        b = _t(g[f'c{i}_b']) if m['use_b'] else None
        kw = dict(dim=m['dim'], act=m['act'], alpha=m['alpha'], gain=m['gain'], clamp=m['clamp'])
        y = ops_ref.bias_act_kernel_ref(x, b, **kw)
        assert rel_err(y, _t(g[f'c{i}_y'])) < 2e-7, (i, m)
        _, _, _, ref, has2 = ops_ref.ACTS[m['act']]
        keep_x = 'x' in ref or has2
        xs, bs, ys = (x if keep_x else None), (b if keep_x else None), (y if 'y' in ref else None)
        dx = ops_ref.bias_act_kernel_ref(dy, bs, xref=xs, yref=ys, grad=1, **kw) if (m['act'] != 'linear' or kw['gain'] not in (None, 1) or kw['clamp'] is not None) else dy
        if m['act'] == 'linear' and m['clamp'] is not None:
            # Reference quirk kept on purpose: 'linear' saves no y (bias_act.py:23 ref=''), so the CUDA gradient kernel
            # sees yref == 0 and never applies the clamp mask (bias_act.cu:141), unlike torch.clamp's autograd.
            assert rel_err(dx, dy * np.float32(m['gain'] if m['gain'] is not None else 1.0)) < 5e-7, (i, m)
            continue
        assert rel_err(dx, _t(g[f'c{i}_dx'])) < 5e-7, (i, m)
        if m['use_b']:
            db = dx.sum([d for d in range(x.ndim) if d != m['dim']])
            assert rel_err(db, _t(g[f'c{i}_db'])) < 5e-7, (i, m)
        # second order: d<dx,ddx>/d(dy) is the same grad=1 kernel applied to ddx; d/dx is the grad=2 kernel
        g2dy = ops_ref.bias_act_kernel_ref(ddx, bs, xref=xs, yref=ys, grad=1, **kw) if dx is not dy else ddx
        assert rel_err(g2dy, _t(g[f'c{i}_g2_dy'])) < 5e-7, (i, m)
        if has2:
            g2x = ops_ref.bias_act_kernel_ref(ddx, bs, xref=xs, yref=ys, dy=dy, grad=2, **kw)
            assert (g2x - _t(g[f'c{i}_g2_x'])).abs().max() < 5e-7 * max(1.0, float(_t(g[f'c{i}_g2_x']).abs().max())), (i, m)
        yt = ops_ref.bias_act_ref_torch(x, b, **kw)
        assert rel_err(yt, _t(g[f'c{i}_y'])) < 2e-7


def test_conv2d_resample_restatement():
    g, meta = load_golden('conv2d_resample_cases.npz')
    f = _t(g['f'])
    for i, m in enumerate(meta):
        x = _t(g[f'c{i}_x']).requires_grad_(True); w = _t(g[f'c{i}_w']).requires_grad_(True)
        y = ops_ref.conv2d_resample_ref(x, w, f=f, up=m['up'], down=m['down'], padding=m['k'] // 2, flip_weight=m['flip_weight'])
        dx, dw = torch.autograd.grad(y, [x, w], _t(g[f'c{i}_dy']))
        assert rel_err(y, _t(g[f'c{i}_y'])) < 1e-5, (i, m)
        assert rel_err(dx, _t(g[f'c{i}_dx'])) < 1e-5 and rel_err(dw, _t(g[f'c{i}_dw'])) < 1e-5, (i, m)


def test_modulated_conv2d_restatement():
    g, meta = load_golden('modconv_cases.npz')
    f = _t(g['f'])
    for i, m in enumerate(meta):
        x = _t(g[f'c{i}_x']).requires_grad_(True); w = _t(g[f'c{i}_w']).requires_grad_(True); s = _t(g[f'c{i}_s']).requires_grad_(True)
        y = ops_ref.modulated_conv2d_ref(x, w, s, up=m['up'], padding=m['k'] // 2, resample_filter=f, demodulate=m['demod'],
                                         flip_weight=(m['up'] == 1), fused_modconv=m['fused'])
        dx, dw, ds = torch.autograd.grad(y, [x, w, s], _t(g[f'c{i}_dy']))
        for a, k in ((y, 'y'), (dx, 'dx'), (dw, 'dw'), (ds, 'ds')):
            assert rel_err(a, _t(g[f'c{i}_{k}'])) < 2e-5, (i, m, k)


def test_synthesis_network_restatement():
    g, meta = load_golden('synthesis_tiny.npz')
    cfg = sr.SynthesisConfig(**meta)
    P = {k[2:]: _t(g[k]).clone().requires_grad_(True) for k in g.files if k.startswith('p:') and 'resample_filter' not in k
         and not k.endswith('freqs') and not k.endswith('phase_scales')}
    ws = _t(g['ws']).requires_grad_(True)
    t = _t(g['t']); mz = _t(g['motion_z'])
    mv = sr.motion_encoder(P, cfg, t, mz)
    assert rel_err(mv, _t(g['motion_v'])) < 1e-5
    img = sr.synthesis_forward(P, cfg, ws, t, motion_z=mz, fused_modconv=False)
    assert rel_err(img, _t(g['img_train'])) < 1e-5
    names = sorted(k[2:] for k in g.files if k.startswith('g:'))
    grads = torch.autograd.grad(img, [ws] + [P[n] for n in names], _t(g['dimg']))
    assert rel_err(grads[0], _t(g['d_ws'])) < 1e-4
    for n, gr in zip(names, grads[1:]):
        assert rel_err(gr, _t(g['g:' + n])) < 1e-4, n
    with torch.no_grad():
        img_e = sr.synthesis_forward(P, cfg, ws, t, motion_z=mz, fused_modconv=True)
    assert rel_err(img_e, _t(g['img_eval'])) < 1e-5
    assert cfg.num_ws == ws.shape[1]


def test_synthesis_network_with_noise_restatement_and_native_cpu_path():
    """use_noise = true / noise_mode = 'const' (networks.py:119-121,130-134): the oracle restatement and the native network's
    layer-by-layer CPU formulation against the golden minted from the reference (image, d ws, every parameter gradient incl. noise_strength)."""
    from stylegan_v_b200.synthesis import SynthesisNetwork
    g, meta = load_golden('synthesis_noise_tiny.npz')
    cfg = sr.SynthesisConfig(**meta)
    skip = ('resample_filter', 'freqs', 'phase_scales')
    P = {k[2:]: _t(g[k]).clone().requires_grad_(not k.endswith('noise_const')) for k in g.files if k.startswith('p:') and not k.endswith(skip)}
    ws = _t(g['ws']).requires_grad_(True)
    t = _t(g['t']); mz = _t(g['motion_z'])
    names = sorted(k[2:] for k in g.files if k.startswith('g:'))
    assert sum(n.endswith('noise_strength') for n in names) == 5
    img = sr.synthesis_forward(P, cfg, ws, t, motion_z=mz, fused_modconv=False, noise_mode='const')
    assert rel_err(img, _t(g['img_train'])) < 1e-5
    grads = torch.autograd.grad(img, [ws] + [P[n] for n in names], _t(g['dimg']))
    for n, gr in zip(['ws'] + names, grads):
        assert rel_err(gr, _t(g['d_ws'] if n == 'ws' else g['g:' + n])) < 1e-4, n
    net = SynthesisNetwork.from_config(cfg)
    net.load_state_dict({k[2:]: _t(g[k]) for k in g.files if k.startswith('p:')}, strict=True)
    net.train()
    ws2 = _t(g['ws']).requires_grad_(True)
    img2 = net(ws2, t, motion_z=mz, noise_mode='const')
    assert rel_err(img2, _t(g['img_train'])) < 1e-5
    params = dict(net.named_parameters())
    grads2 = torch.autograd.grad(img2, [params[n] for n in names], _t(g['dimg']))
    for n, gr in zip(names, grads2):
        assert rel_err(gr, _t(g['g:' + n])) < 1e-4, n
    with torch.no_grad():      # 'none' drops the noise, 'random' draws fresh planes
        assert rel_err(net(ws2, t, motion_z=mz, noise_mode='none'), img2) > 1e-3
        assert rel_err(net(ws2, t, motion_z=mz, noise_mode='random'), img2) > 1e-3


def test_flop_model_matches_baseline_md():
    # 29.870 GFLOP/frame at 256^2, 15.336 at 64^2, 148.596 at 1024^2
    assert abs(sr.conv_flops_per_frame(sr.SynthesisConfig(img_resolution=256)) / 1e9 - 29.870) < 0.01
    assert abs(sr.conv_flops_per_frame(sr.SynthesisConfig(img_resolution=64)) / 1e9 - 15.336) < 0.01
    assert abs(sr.conv_flops_per_frame(sr.SynthesisConfig(img_resolution=1024, channel_base=32768)) / 1e9 - 148.596) < 0.01


def test_reference_cuda_plugins_built_and_loadable():
    """oracle/_ref holds the reference's own CUDA plugins compiled for sm_100a (oracle/build_ref.py); they must import without a GPU and
    expose the two pybind entry points the GPU comparison test calls.  Skipped where they were never built (no reference tree)."""
    import pytest
    from oracle import build_ref, ref_loader
    if not all(__import__('os').path.exists(build_ref.plugin_path(n)) for n in build_ref.PLUGINS):
        if not ref_loader.available():
            pytest.skip('reference tree absent and oracle/_ref not built')
        build_ref.build()
    up, ba = build_ref.load_plugin('upfirdn2d_plugin'), build_ref.load_plugin('bias_act_plugin')
    assert callable(up.upfirdn2d) and callable(ba.bias_act)
    import subprocess
    sass = subprocess.run(['cuobjdump', '-lelf', build_ref.plugin_path('upfirdn2d_plugin')], capture_output=True, text=True).stdout
    assert 'sm_100a' in sass                                           # built for the B200, not for a fallback architecture
