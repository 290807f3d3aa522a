"""Native synthesis network (fused sm_100a kernels, through the C ABI) vs the oracle / reference-minted goldens.

Tolerances: every contraction multiplies TF32-rounded operands (unit roundoff 2^-11) with fp32 accumulation; per
layer that is <= 1e-3 normwise (tests/test_conv_gpu.py).  Whole-network OUTPUTS are held to 3e-3 (measured 6e-4).

Gradients need care: a leaky-ReLU gradient is discontinuous in the forward value, so any forward that is not bit-equal
to fp32 (TF32 here, cuDNN-TF32 or fp16 in the reference's own fast modes) flips the slope of the few activations whose
pre-activation is within ~1e-3*sigma of zero.  In this tiny network a channel sums only 1.5k-6k pixels, so 1-3 flipped
elements move a (random-sign) gradient sum by 1-4 %.  That is a property of the forward precision, not of the backward
kernels, so the backward is validated in two ways:
  * test_fused_layer_backward_with_matched_forward: the reference forward is fed the same TF32-rounded operands (masks
    then agree) and every gradient of the fused layer (dx, dW, dstyles, ddcoefs, dbias) must agree to 2e-3;
  * whole-network gradients vs the true-fp32 goldens: cosine similarity >= 0.998 and normwise error <= 6e-2."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import synthesis_ref as sr
from stylegan_v_b200.synthesis import SynthesisNetwork
from stylegan_v_b200 import modconv

pytestmark = pytest.mark.gpu


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _load(g, meta):
    cfg = sr.SynthesisConfig(**meta)
    net = SynthesisNetwork.from_config(cfg)
    sd = {k[2:]: _t(g[k]) for k in g.files if k.startswith('p:')}
    missing, unexpected = net.load_state_dict(sd, strict=True), None
    return cfg, net.cuda()


def test_state_dict_keys_equal_reference():
    g, meta = load_golden('synthesis_tiny.npz')
    cfg = sr.SynthesisConfig(**meta)
    net = SynthesisNetwork.from_config(cfg)
    ref_keys = sorted(k[2:] for k in g.files if k.startswith('p:'))
    assert sorted(net.state_dict().keys()) == ref_keys
    for k, v in net.state_dict().items():
        assert tuple(v.shape) == g['p:' + k].shape, k
    assert net.num_ws == cfg.num_ws


def test_fused_layer_vs_golden_modconv():
    g, meta = load_golden('modconv_cases.npz')
    for i, m in enumerate(meta):
        if m['I'] % 32 or m['O'] % 16 or m['fused']:
            continue
        x = _t(g[f'c{i}_x']).cuda().requires_grad_(True)
        w = _t(g[f'c{i}_w']).cuda().requires_grad_(True)
        s = _t(g[f'c{i}_s']).cuda().requires_grad_(True)
        y = modconv.fused_modulated_conv(x, w, s, None, up=m['up'], demodulate=m['demod'], act='linear', gain=1.0, flip_weight=(m['up'] == 1))
        dx, dw, ds = torch.autograd.grad(y, [x, w, s], _t(g[f'c{i}_dy']).cuda())
        for a, k in ((y, 'y'), (dx, 'dx'), (dw, 'dw'), (ds, 'ds')):
            assert rel_err(a, _t(g[f'c{i}_{k}'])) < 2e-3, (i, m, k, rel_err(a, _t(g[f'c{i}_{k}'])))


def test_network_forward_backward_vs_golden():
    g, meta = load_golden('synthesis_tiny.npz')
    cfg, net = _load(g, meta)
    ws = _t(g['ws']).cuda().requires_grad_(True)
    t = _t(g['t']).cuda(); mz = _t(g['motion_z']).cuda()
    mv = net.motion_encoder(t, motion_z=mz, t_max=float(t.max()))['motion_v']
    assert rel_err(mv, _t(g['motion_v'])) < 1e-4
    img = net(ws, t, motion_z=mz, t_max=float(t.max()))
    assert img.shape == g['img_train'].shape and not img.is_contiguous(memory_format=torch.channels_last) or img.shape[1] == 1
    e = rel_err(img, _t(g['img_train']))
    assert e < 3e-3, e
    names = sorted(k[2:] for k in g.files if k.startswith('g:'))
    params = dict(net.named_parameters())
    grads = torch.autograd.grad(img, [ws] + [params[n] for n in names], _t(g['dimg']).cuda())
    def cos(a, b):
        a, b = a.detach().double().cpu().flatten(), b.double().flatten()
        return float((a @ b) / (a.norm() * b.norm()))
    assert rel_err(grads[0], _t(g['d_ws'])) < 6e-2 and cos(grads[0], _t(g['d_ws'])) > 0.998
    worst = max((rel_err(gr, _t(g['g:' + n])), n) for n, gr in zip(names, grads[1:]))
    assert worst[0] < 6e-2, worst
    worst_cos = min((cos(gr, _t(g['g:' + n])), n) for n, gr in zip(names, grads[1:]))
    assert worst_cos[0] > 0.998, worst_cos


def _tf32_ste(t):
    """TF32 rounding (round-to-nearest, ties away) with a straight-through gradient."""
    r = ((t.detach().float().contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32).to(t.dtype)
    return t + (r - t).detach()


@pytest.mark.parametrize('up', [1, 2])
def test_fused_layer_backward_with_matched_forward(up):
    gen = torch.Generator().manual_seed(17 + up)
    N, I, O, H = 3, 64, 64, 12
    x = torch.randn(N, I, H, H, generator=gen).cuda().requires_grad_(True)
    w = torch.randn(O, I, 3, 3, generator=gen).cuda().requires_grad_(True)
    s = (torch.randn(N, I, generator=gen) + 1).cuda().requires_grad_(True)
    d = (torch.rand(N, O, generator=gen) + 0.5).cuda().requires_grad_(True)
    b = (0.3 * torch.randn(O, generator=gen)).cuda().requires_grad_(True)
    gain = float(np.sqrt(2))
    y = modconv._FusedModConv.apply(x, w, s, d, b, up, 'lrelu', gain, up == 1, None, None, None)
    dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(5)).cuda()
    got = torch.autograd.grad(y, [x, w, s, d, b], dy)
    # reference: same rounded operands, fp64 arithmetic, autograd
    X, W, S, D, B = (t.detach().double().requires_grad_(True) for t in (x, w, s, d, b))
    xs = _tf32_ste((X * S[:, :, None, None]).float()).double() if False else None
    xs32 = (x.detach() * s.detach()[:, :, None, None])
    delta_x = (((xs32.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32) - xs32).double()
    w32 = w.detach()
    delta_w = (((w32.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32) - w32).double()
    xs = X * S[:, :, None, None] + delta_x
    wr = W + delta_w
    F = torch.nn.functional
    if up == 1:
        c = F.conv2d(xs, wr, padding=1)
    else:
        u = F.conv_transpose2d(xs, wr.transpose(0, 1), stride=2)
        f = modconv._fir(x.device).double()
        c = F.conv2d(F.pad(u, [1, 1, 1, 1]), (f * 4).flip([0, 1])[None, None].repeat(O, 1, 1, 1), groups=O)
    yr = F.leaky_relu(c * D[:, :, None, None] + B[None, :, None, None], 0.2) * gain
    assert rel_err(y, yr) < 2e-5
    ref = torch.autograd.grad(yr, [X, W, S, D, B], dy.double())
    for name, a, r in zip(('dx', 'dw', 'dstyles', 'ddcoefs', 'dbias'), got, ref):
        assert rel_err(a, r) < 2e-3, (up, name, rel_err(a, r))


def test_network_vs_oracle_other_config():
    """A second architecture (64x64, 64 channels) against the CPU oracle computed on the fly."""
    cfg = sr.SynthesisConfig(img_resolution=64, w_dim=128, channel_base=2048, channel_max=64, motion_z_dim=64, motion_v_dim=64, time_enc_dim=32)
    P = sr.init_params(cfg, seed=3)
    net = SynthesisNetwork.from_config(cfg)
    sd = net.state_dict()
    for k in sd:
        if k in P:
            sd[k] = P[k]
    net.load_state_dict(sd)
    net = net.cuda()
    gen = torch.Generator().manual_seed(5)
    B, Fr = 2, 2
    ws = torch.randn(B, cfg.num_ws, cfg.w_dim, generator=gen)
    t = torch.tensor([[3.0, 40.5], [700.25, 701.0]])
    mz = torch.randn(B, sr.max_traj_len(cfg, 1023.0), cfg.motion_z_dim, generator=gen)
    ref = sr.synthesis_forward(P, cfg, ws, t, motion_z=mz, fused_modconv=False)
    img = net(ws.cuda(), t.cuda(), motion_z=mz.cuda())
    e = rel_err(img, ref)
    assert e < 3e-3, e


def test_network_vs_oracle_32_channel_ladder():
    """The channel ladder of config 5 (1024x1024, fmaps 1) ends in 64 -> 32 -> 32 channels; the same ladder at 128x128 against the
    CPU oracle: exercises the 32-column conv tiles, the 32-channel weight-gradient tiles and the 32-channel FIR / ToRGB paths."""
    cfg = sr.SynthesisConfig(img_resolution=128, w_dim=128, channel_base=4096, channel_max=128, motion_z_dim=64, motion_v_dim=64, time_enc_dim=32)
    assert [cfg.channels(r) for r in (32, 64, 128)] == [128, 64, 32]
    P = sr.init_params(cfg, seed=4)
    net = SynthesisNetwork.from_config(cfg)
    sd = net.state_dict()
    for k in sd:
        if k in P:
            sd[k] = P[k]
    net.load_state_dict(sd)
    net = net.cuda()
    gen = torch.Generator().manual_seed(6)
    ws = torch.randn(2, cfg.num_ws, cfg.w_dim, generator=gen)
    t = torch.tensor([[1.0], [250.5]])
    mz = torch.randn(2, sr.max_traj_len(cfg, 1023.0), cfg.motion_z_dim, generator=gen)
    ref = sr.synthesis_forward(P, cfg, ws, t, motion_z=mz, fused_modconv=False)
    wsg = ws.cuda().requires_grad_(True)
    img = net(wsg, t.cuda(), motion_z=mz.cuda())
    assert rel_err(img, ref) < 3e-3
    img.square().mean().backward()
    assert torch.isfinite(wsg.grad).all() and all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters() if p.requires_grad)


def test_config5_1024_forward_runs():
    """BASELINE config 5 geometry (1024x1024, channel_base 32768: ... 512:64, 1024:32) at batch 1: shape, finiteness, determinism."""
    net = SynthesisNetwork(img_resolution=1024, channel_base=32768).cuda().eval()
    gen = torch.Generator().manual_seed(7)
    ws = torch.randn(1, net.num_ws, net.w_dim, generator=gen).cuda()
    t = torch.zeros(1, 1).cuda()
    mz = torch.randn(1, net.motion_encoder.traj_len(), 512, generator=gen).cuda()
    with torch.no_grad():
        a = net(ws, t, motion_z=mz)
        b = net(ws, t, motion_z=mz)
    assert a.shape == (1, 3, 1024, 1024) and torch.isfinite(a).all() and torch.equal(a, b)


def test_layer_elementwise_kernels():
    from stylegan_v_b200 import conv as C
    gen = torch.Generator().manual_seed(3)
    for (N, Cc, H) in ((3, 64, 9), (2, 512, 4), (2, 128, 16), (1, 1024, 4), (2, 16, 8), (2, 256, 19), (3, 512, 33)):
        cl = lambda t: t.cuda().contiguous(memory_format=torch.channels_last)
        dy = cl(torch.randn(N, Cc, H, H, generator=gen)); v = torch.randn(N, Cc, H, H, generator=gen)
        bias = torch.randn(Cc, generator=gen).cuda(); gain = float(np.sqrt(2))
        y = cl(torch.nn.functional.leaky_relu(v.cuda() + bias[None, :, None, None], 0.2) * gain)
        dz, db, dd = C.act_bwd(dy, y, bias, 'lrelu', gain, True, True)
        dz_ref = torch.where(y > 0, dy, dy * 0.2) * gain
        assert rel_err(dz, dz_ref) < 1e-6 and rel_err(db, dz_ref.double().sum([0, 2, 3])) < 1e-5
        assert rel_err(dd, (dz_ref.double() * v.cuda().double()).sum([2, 3])) < 1e-4
        x = cl(torch.randn(N, Cc, H, H, generator=gen)); s = torch.randn(N, Cc, generator=gen).cuda()
        dxs = cl(torch.randn(N, Cc, H, H, generator=gen)); keep = dxs.clone()
        dx, ds = C.scale_reduce(dxs, x, s)
        assert rel_err(dx, keep * s[:, :, None, None]) < 1e-6 and rel_err(ds, (keep.double() * x.double()).sum([2, 3])) < 1e-5
        if Cc >= 16:
            wmod = torch.randn(N, 3, Cc, generator=gen).cuda(); b3 = torch.randn(3, generator=gen).cuda()
            yrgb = C.torgb_fwd(x, wmod, b3)
            ref = torch.einsum('nchw,njc->njhw', x.double(), wmod.double()) + b3.double()[None, :, None, None]
            assert yrgb.is_contiguous() and rel_err(yrgb, ref) < 1e-5
            g3 = torch.randn(N, 3, H, H, generator=gen).cuda()
            dxr, dwm = C.torgb_bwd(g3, x, wmod)
            assert rel_err(dxr, torch.einsum('njhw,njc->nchw', g3.double(), wmod.double())) < 1e-5
            assert rel_err(dwm, torch.einsum('njhw,nchw->njc', g3.double(), x.double())) < 1e-5
            # the same ToRGB gradient folded into act_bwd (sgv_modconv_act_bwd_rgb): dy + torgb data gradient, and d(wmod) from y
            for with_dy in (True, False):
                dz2, db2, dd2, dwm2 = C.act_bwd(dy if with_dy else None, y, bias, 'lrelu', gain, True, True, dyimg=g3, wmod=wmod)
                dtot = torch.einsum('njhw,njc->nchw', g3.double(), wmod.double()) + (dy.double() if with_dy else 0)
                dz2_ref = torch.where(y > 0, dtot, dtot * 0.2) * gain
                assert rel_err(dz2, dz2_ref) < 1e-6 and rel_err(db2, dz2_ref.sum([0, 2, 3])) < 1e-5
                assert rel_err(dd2, (dz2_ref * v.cuda().double()).sum([2, 3])) < 1e-4
                assert rel_err(dwm2, torch.einsum('njhw,nchw->njc', g3.double(), y.double())) < 1e-5


def test_torgb_modulated_weight_node():
    """wmod = weight * styles * gain of the ToRGB layers (networks.py:159-160) as one launch + one backward launch vs the torch formula; styles is
    a column slice of the stacked affine output."""
    from stylegan_v_b200.synthesis import _ToRgbWmod
    gen = torch.Generator().manual_seed(5)
    for N, C in ((32, 512), (3, 64), (5, 200)):
        w = torch.randn(3, C, 1, 1, generator=gen).cuda().requires_grad_(True)
        wide = torch.randn(N, C + 24, generator=gen).cuda().requires_grad_(True)
        s = wide[:, 8:8 + C]
        gain = 1 / np.sqrt(C)
        got = _ToRgbWmod.apply(w, s, float(gain))
        ref = w.reshape(1, 3, C) * (s * gain).unsqueeze(1)
        dy = torch.randn(N, 3, C, generator=gen).cuda()
        gw, gs = torch.autograd.grad(got, [w, wide], dy)
        rw, rs = torch.autograd.grad(ref, [w, wide], dy)
        assert rel_err(got, ref) < 1e-6 and rel_err(gw, rw) < 1e-5 and rel_err(gs, rs) < 1e-5
