"""The exact-fp32 dense kernels (csrc/dense_f32.cu through stylegan_v_b200/dense.py) against fp64 torch on the GPU: FullyConnectedLayer
forward + all gradients, the one-launch stacked style affines (column groups), the conv1d-as-windows formulation of the motion trajectory
against F.conv1d, the whole MotionMappingNetwork against its library-op formulation, and the module-level routing switches.
Bar: 2e-6 normwise against fp64 (plain fp32 FMAs with round-to-nearest accumulation; the tcgen05 route this replaced needed 2e-5 ... 1e-4)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err
from stylegan_v_b200 import _lib, dense
from stylegan_v_b200.networks import FullyConnectedLayer, MappingNetwork
from stylegan_v_b200.time_encoder import MotionMappingNetwork

pytestmark = pytest.mark.gpu
BAR = 2e-6


@pytest.mark.parametrize('M,K,O,act,bias', [(32, 512, 512, 'linear', True), (48, 512, 1536, 'linear', True), (7, 64, 64, 'lrelu', True),
                                            (96, 8192, 512, 'lrelu', True), (33, 512, 1024, 'linear', False), (2, 512, 32, 'linear', True),
                                            (5, 512, 1, 'linear', True), (70, 36, 13, 'lrelu', True), (200, 4096, 512, 'lrelu', True)])
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
    assert _lib.launch_count() - n0 == 3            # forward, data gradient, weight (+ bias) gradient
    xd, wd = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    bd = b.detach().double().requires_grad_(True) if bias else None
    r = xd @ (wd * wg).t() + (bd * bg if bias else 0)
    r = (F.leaky_relu(r, 0.2) if act == 'lrelu' else r) * gain
    ref = torch.autograd.grad(r, [xd, wd] + ([bd] if bias else []), dy.double())
    assert rel_err(y, r) < BAR
    for a, e, n in zip(got, ref, ('dx', 'dw', 'db')):
        assert rel_err(a, e) < BAR, (n, rel_err(a, e))


def test_linear_strided_rows():
    """x may be a row-strided view (the motion encoder feeds y2[:, 0] of a [R, 2, C] tensor)."""
    base = torch.randn(24, 2, 256, device='cuda')
    w = torch.randn(128, 256, device='cuda')
    for j in (0, 1):
        x = base[:, j]
        assert rel_err(dense.linear(x, w, None, 0.1), (x.double() @ w.double().t()) * 0.1) < BAR


def test_stacked_affine_vs_per_group():
    """One launch for all style affines: column group g reads ws[:, g, :]; gradients land in the right w rows."""
    g = torch.Generator().manual_seed(3)
    M, G, K = 32, 5, 512
    widths = [[512], [512, 256], [256, 128, 64], [64], [8]]          # layers per w index
    ws = torch.randn(M, G, K, generator=g).cuda().requires_grad_(True)
    col, weights, biases = [0], [], []
    for ws_i in widths:
        for o in ws_i:
            weights.append(torch.randn(o, K, generator=g).cuda().requires_grad_(True))
            biases.append(torch.randn(o, generator=g).cuda().requires_grad_(True))
        col.append(col[-1] + sum(ws_i))
    groups = dense.make_groups(col, list(range(G)), K, ws.device)
    gain = 1 / np.sqrt(K)
    n0 = _lib.launch_count()
    s = dense.stacked_affine(ws, torch.cat(weights), torch.cat(biases), groups, gain)
    dy = torch.randn(s.shape, generator=g).cuda()
    got = torch.autograd.grad(s, [ws] + weights + biases, dy)
    assert _lib.launch_count() - n0 == 3
    wsd = ws.detach().double().requires_grad_(True)
    wd = [w.detach().double().requires_grad_(True) for w in weights]
    bd = [b.detach().double().requires_grad_(True) for b in biases]
    pieces, li = [], 0
    for gi, ws_i in enumerate(widths):
        for _ in ws_i:
            pieces.append(wsd[:, gi] @ (wd[li] * gain).t() + bd[li])
            li += 1
    r = torch.cat(pieces, dim=1)
    ref = torch.autograd.grad(r, [wsd] + wd + bd, dy.double())
    assert rel_err(s, r) < BAR
    for a, e in zip(got, ref):
        assert rel_err(a, e) < BAR


@pytest.mark.parametrize('P', [2, 30])
def test_conv1d_slabs_vs_conv1d(P):
    """Windows of a [B, L, C] sequence (gathered by per-slab offsets), then a second layer on the slab output with the overlap-add gradient."""
    g = torch.Generator().manual_seed(P)
    B, L, C, O, k = 3, 60, 64, 32, 5
    z = torch.randn(B, L, C, generator=g).cuda()
    w1 = torch.randn(C, C, k, generator=g).cuda().requires_grad_(True)
    b1 = torch.randn(C, generator=g).cuda().requires_grad_(True)
    w2 = torch.randn(O, C, k, generator=g).cuda().requires_grad_(True)
    b2 = torch.randn(O, generator=g).cuda().requires_grad_(True)
    starts = torch.tensor([0, 7, L - (P + 2 * (k - 1))], device='cuda')                      # slab start per batch row
    base = (torch.arange(B, device='cuda') * L + starts) * C
    y1 = dense.conv1d_slabs(z, base, P + k - 1, w1, b1, 0.05, 0.5, 'lrelu')
    y2 = dense.conv1d_slabs(y1, None, P, w2, b2, 0.07, 0.3, 'lrelu')
    dy = torch.randn(y2.shape, generator=g).cuda()
    got = torch.autograd.grad(y2, [w1, b1, w2, b2], dy)
    # fp64 reference: full conv1d over the slab of each batch row
    slabs = torch.stack([z[i, int(s):int(s) + P + 2 * (k - 1)] for i, s in enumerate(starts)]).double()          # [B, P + 2(k-1), C]
    wd = [t.detach().double().requires_grad_(True) for t in (w1, b1, w2, b2)]
    r1 = F.leaky_relu(F.conv1d(slabs.permute(0, 2, 1), wd[0] * 0.05, wd[1] * 0.5), 0.2)
    r2 = F.leaky_relu(F.conv1d(r1, wd[2] * 0.07, wd[3] * 0.3), 0.2).permute(0, 2, 1)
    ref = torch.autograd.grad(r2, wd, dy.double())
    assert rel_err(y2, r2) < BAR
    for a, e, n in zip(got, ref, ('dw1', 'db1', 'dw2', 'db2')):
        assert rel_err(a, e) < 5e-6, (n, rel_err(a, e))


@pytest.mark.parametrize('frames', [1, 3, 16])
def test_motion_encoder_windows_vs_library(frames):
    """MotionMappingNetwork on the window GEMMs (few frames per clip) / the full-trajectory GEMMs (many) vs an fp64 CPU evaluation of its
    PyTorch-op formulation (F.conv1d + matmul + elementwise tail): motion_v and every parameter gradient."""
    torch.manual_seed(frames)
    enc = MotionMappingNetwork(z_dim=512, v_dim=512, kernel_size=11, motion_z_distance=16, time_enc_dim=256, max_num_frames=1024).cuda()
    B = 4
    t = torch.randint(0, 1000, (B, frames), device='cuda').float()
    L = enc.traj_len()
    mz = torch.randn(B, L, 512, device='cuda')
    n0 = _lib.launch_count()
    v = enc(t, motion_z=mz)['motion_v']
    dy = torch.randn_like(v)
    params = list(enc.parameters())
    got = torch.autograd.grad(v, params, dy)
    assert _lib.launch_count() - n0 >= 2 + 3 + 3 + 3 + 2      # conv layers (the first has no data gradient), heads, aligners, fused tail
    import copy
    ref_enc = copy.deepcopy(enc).cpu().double()                # fp64 evaluation of the same module on its PyTorch-op formulation
    v_ref = ref_enc(t.cpu().double(), motion_z=mz.cpu().double())['motion_v']
    ref = torch.autograd.grad(v_ref, list(ref_enc.parameters()), dy.cpu().double())
    # sin / cos of arguments up to ~800 rad turn the fp32 rounding of the phase arguments (6e-8 relative) into ~5e-5 absolute
    assert rel_err(v, v_ref) < 5e-4, rel_err(v, v_ref)
    for p, a, e in zip(params, got, ref):
        assert rel_err(a, e) < 1e-3, (tuple(p.shape), rel_err(a, e))


def test_module_routes():
    """First-order callers take the kernels; the any-order formulation (default) stays on torch ops and is twice differentiable."""
    fc = FullyConnectedLayer(512, 512, activation='lrelu', lr_multiplier=0.01).cuda()
    x = torch.randn(16, 512, device='cuda', requires_grad=True)
    n0 = _lib.launch_count()
    a = fc(x)                                        # any-order route: no launch of the dense kernel (bias_act kernel only)
    n1 = _lib.launch_count()
    b = fc(x, fused=True)
    assert _lib.launch_count() - n1 == 1 and n1 - n0 <= 1
    assert rel_err(b, a) < 1e-5
    gx, = torch.autograd.grad(a.square().sum(), x, create_graph=True)
    gx.sum().backward()                              # second order works on the default route
    assert fc.weight.grad is not None
    mp = MappingNetwork(z_dim=512, c_dim=0, w_dim=512, num_ws=14, num_layers=2).cuda().eval()
    n2 = _lib.launch_count()
    ws = mp(torch.randn(8, 512, device='cuda'), torch.zeros(8, 0, device='cuda'))
    assert ws.shape == (8, 14, 512) and _lib.launch_count() - n2 >= 2


@pytest.mark.parametrize('N,O,I', [(32, 512, 512), (5, 64, 64), (3, 12, 20), (33, 256, 128)])
def test_demod_coefs_vs_fp64(N, O, I):
    """Demodulation coefficients (networks.py:57-59) and their gradients from csrc/demod.cu against the materialised fp64 formula; styles are
    a column slice of a wider tensor like the stacked affine output."""
    from stylegan_v_b200.modconv import demod_coefs
    g = torch.Generator().manual_seed(N + O)
    w = torch.randn(O, I, 3, 3, generator=g).cuda().requires_grad_(True)
    wide = (torch.randn(N, I + 16, generator=g) + 1).cuda().requires_grad_(True)
    s = wide[:, 8:8 + I]
    n0 = _lib.launch_count()
    dc = demod_coefs(w, s)
    ddc = torch.randn(N, O, generator=g).cuda()
    gw, gs = torch.autograd.grad(dc, [w, wide], ddc)
    assert _lib.launch_count() - n0 == 3
    wd, sd = w.detach().double().requires_grad_(True), wide.detach().double().requires_grad_(True)
    wm = wd.unsqueeze(0) * sd[:, 8:8 + I].reshape(N, 1, I, 1, 1)
    ref = (wm.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt()
    rw, rs = torch.autograd.grad(ref, [wd, sd], ddc.double())
    assert rel_err(dc, ref) < BAR
    assert rel_err(gw, rw) < 5e-6 and rel_err(gs, rs) < 5e-6, (rel_err(gw, rw), rel_err(gs, rs))
