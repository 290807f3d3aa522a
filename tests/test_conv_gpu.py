"""tcgen05 implicit-GEMM convolution (through the C ABI) vs fp32/fp64 convolution of the same operands.

Tolerance: the kernel multiplies TF32-rounded operands (10-bit mantissa, round-to-nearest) and accumulates in
fp32, so against a true-fp32 contraction we require normwise relative error <= 1e-3 (BASELINE north_star);
against a reference fed the SAME TF32-rounded operands the only difference is accumulation order: <= 2e-5."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err
from stylegan_v_b200 import conv as C

pytestmark = pytest.mark.gpu


def tf32_round(t):
    """round-to-nearest (ties away) to 10 explicit mantissa bits, like cvt.rna.tf32.f32"""
    i = t.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32)


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize('N,Cin,Cout,H,W', [(2, 32, 64, 16, 16), (1, 64, 128, 8, 8), (3, 64, 64, 20, 12), (2, 128, 256, 16, 16),
                                            (8, 32, 64, 4, 4), (1, 32, 512, 8, 8), (2, 96, 64, 33, 17),
                                            (1, 64, 128, 40, 40), (2, 32, 256, 24, 31), (1, 128, 512, 16, 16)])
def test_conv3x3_plain(N, Cin, Cout, H, W):
    g = torch.Generator().manual_seed(N * 1000 + Cin)
    x = torch.randn(N, Cin, H, W, generator=g).cuda()
    w = torch.randn(Cout, Cin, 3, 3, generator=g).cuda()
    taps, offs = C.conv3x3_taps()
    wp = C.prep_weights(w, taps)
    assert torch.equal(wp, tf32_round(w).permute(2, 3, 0, 1).reshape(9, Cout, Cin))
    y = C.igemm_conv(_cl(x), wp, offs)
    assert y.shape == (N, Cout, H, W)
    ref_same = F.conv2d(tf32_round(x).double(), tf32_round(w).double(), padding=1)
    ref_fp32 = F.conv2d(x.double(), w.double(), padding=1)
    assert rel_err(y, ref_same) < 2e-5
    assert rel_err(y, ref_fp32) < 1e-3


def test_conv1x1_and_modulated_epilogue():
    g = torch.Generator().manual_seed(7)
    N, Cin, Cout, H = 4, 64, 64, 16
    x = torch.randn(N, Cin, H, H, generator=g).cuda()
    w = torch.randn(Cout, Cin, 3, 3, generator=g).cuda()
    s = (torch.randn(N, Cin, generator=g) + 1).cuda()
    d = (torch.rand(N, Cout, generator=g) + 0.5).cuda()
    b = torch.randn(Cout, generator=g).cuda()
    taps, offs = C.conv3x3_taps()
    wp = C.prep_weights(w, taps)
    y = C.igemm_conv(_cl(x), wp, offs, a_scale=s, o_scale=d, bias=b, act='lrelu', gain=float(np.sqrt(2)), clamp=None)
    xs = (x * s[:, :, None, None])
    ref = F.conv2d(xs.double(), w.double(), padding=1) * d.double()[:, :, None, None] + b.double()[None, :, None, None]
    ref = F.leaky_relu(ref, 0.2) * np.sqrt(2)
    assert rel_err(y, ref) < 1e-3
    ref_same = F.conv2d(tf32_round(xs).double(), tf32_round(w).double(), padding=1) * d.double()[:, :, None, None] + b.double()[None, :, None, None]
    ref_same = F.leaky_relu(ref_same, 0.2) * np.sqrt(2)
    assert rel_err(y, ref_same) < 2e-5
    # 1x1
    w1 = torch.randn(Cout, Cin, 1, 1, generator=g).cuda()
    y1 = C.igemm_conv(_cl(x), C.prep_weights(w1, [(0, 0)]), [(0, 0)], clamp=2.0)
    ref1 = F.conv2d(tf32_round(x).double(), tf32_round(w1).double()).clamp(-2, 2)
    assert rel_err(y1, ref1) < 2e-5


@pytest.mark.parametrize('N,Cin,Cout,H', [(32, 1024, 512, 4), (32, 512, 512, 8), (4, 512, 128, 8), (2, 256, 64, 6)])
def test_small_planes_split_k_cluster(N, Cin, Cout, H):
    """The 4x4 ... 16x16 layers (b4.conv1: 1024 -> 512 on 4x4 planes) have few output tiles and a long contraction: the per-tap kernel
    splits K over a thread-block cluster and reduces the partial tiles through distributed shared memory.  Same results as the
    one-CTA form (checked against fp64 on the same rounded operands), epilogue and fused d(styles) reduction included."""
    g = torch.Generator().manual_seed(H * 7 + N)
    x = torch.randn(N, Cin, H, H, generator=g).cuda()
    w = torch.randn(Cout, Cin, 3, 3, generator=g).cuda()
    s = (torch.randn(N, Cin, generator=g) + 1).cuda()
    d = (torch.rand(N, Cout, generator=g) + 0.5).cuda() / np.sqrt(Cin * 9)
    b = torch.randn(Cout, generator=g).cuda()
    taps, offs = C.conv3x3_taps()
    wp = C.prep_weights(w, taps, x3=False)
    kw = dict(a_scale=s, o_scale=d, bias=b, act='lrelu', gain=float(np.sqrt(2)))
    v = C.igemm_conv(_cl(x), wp, offs, query=True, **kw)
    assert v['kernel'] == 1 and v['cluster'] >= 2, v
    y = C.igemm_conv(_cl(x), wp, offs, **kw)
    xs = tf32_round(x * s[:, :, None, None])
    ref = F.conv2d(xs.double(), tf32_round(w).double(), padding=1) * d.double()[:, :, None, None] + b.double()[None, :, None, None]
    ref = F.leaky_relu(ref, 0.2) * np.sqrt(2)
    assert rel_err(y, ref) < 2e-5
    # data-gradient form: raw accumulator times o_scale, with the fused reduction against a tensor shaped like the output
    rx = _cl(torch.randn(N, Cout, H, H, generator=g).cuda())
    red = torch.zeros(N, Cout, device='cuda')
    y2 = C.igemm_conv(_cl(x), wp, offs, o_scale=d, red_x=rx, red_out=red)
    raw = F.conv2d(tf32_round(x).double(), tf32_round(w).double(), padding=1)
    assert rel_err(y2, raw * d.double()[:, :, None, None]) < 2e-5
    assert rel_err(red, (raw * rx.double()).sum(dim=[2, 3])) < 1e-4


def test_transposed_conv_as_four_phases():
    """conv_transpose2d(stride 2, pad 0) written as 4 polyphase stride-1 launches into one [2h+1, 2w+1] tensor."""
    g = torch.Generator().manual_seed(3)
    N, Cin, Cout, h = 2, 64, 64, 12
    x = torch.randn(N, Cin, h, h, generator=g).cuda()
    w = torch.randn(Cout, Cin, 3, 3, generator=g).cuda()        # layer weight [O, I, 3, 3]; conv_transpose2d takes w.transpose(0,1)
    u = torch.zeros(N, Cout, 2 * h + 1, 2 * h + 1, device='cuda').contiguous(memory_format=torch.channels_last)
    for a in (0, 1):
        for b in (0, 1):
            kys = [a + 2 * m for m in range(2) if a + 2 * m <= 2]
            kxs = [b + 2 * m for m in range(2) if b + 2 * m <= 2]
            taps = [(ky, kx) for ky in kys for kx in kxs]
            offs = [(-(ky - a) // 2, -(kx - b) // 2) for ky, kx in taps]
            view = u[:, :, a::2, b::2]
            C.igemm_conv(_cl(x), C.prep_weights(w, taps), offs, out_view=view)
    ref = F.conv_transpose2d(tf32_round(x).double(), tf32_round(w).double().transpose(0, 1), stride=2)
    assert rel_err(u, ref) < 2e-5


@pytest.mark.parametrize('N,Cin,Cout,h', [(2, 64, 64, 10), (3, 64, 128, 16), (2, 128, 64, 21), (5, 32, 256, 32)])
def test_stride2_input(N, Cin, Cout, h):
    """data gradient of the transposed conv = stride-2 correlation (TMA element strides); h >= 12 runs on the persistent
    kernel with one every-other-pixel patch per (dy, dx) parity class, smaller outputs on the per-tap kernel."""
    g = torch.Generator().manual_seed(5)
    du = torch.randn(N, Cin, 2 * h + 1, 2 * h + 1, generator=g).cuda()
    w = torch.randn(Cout, Cin, 3, 3, generator=g).cuda()
    taps = C.TAPS_3x3
    y = C.igemm_conv(_cl(du), C.prep_weights(w, taps), taps, out_hw=(h, h), in_stride=2)
    ref = F.conv2d(tf32_round(du).double(), tf32_round(w).double(), stride=2)
    assert ref.shape == y.shape
    assert rel_err(y, ref) < 2e-5


@pytest.mark.parametrize('N,Cin,Cout,H', [(2, 64, 128, 16), (3, 32, 64, 12), (8, 64, 64, 4), (2, 256, 128, 8), (1, 128, 32, 20),
                                        (3, 64, 64, 20), (2, 128, 64, 33), (1, 64, 64, 9), (2, 64, 64, 64)])      # last four: the stacked-M kernel (64 output channels)
def test_wgrad_stride1(N, Cin, Cout, H):
    g_ = torch.Generator().manual_seed(N + Cin)
    x = torch.randn(N, Cin, H, H, generator=g_).cuda()
    gy = torch.randn(N, Cout, H, H, generator=g_).cuda()
    s = (torch.rand(N, Cin, generator=g_) + 0.5).cuda()
    d = (torch.rand(N, Cout, generator=g_) + 0.5).cuda()
    taps = C.TAPS_3x3
    q = C.igemm_wgrad(_cl(gy), _cl(x), [(0, 0)] * 9, [(ky - 1, kx - 1) for ky, kx in taps], (H, H), g_scale=d, x_scale=s, x3=False, query=True)
    assert q['kernel'] == (3 if (Cout == 64 and Cin % 64 == 0 and H >= 8) else 2 if H >= 8 else 1), q
    dw = C.igemm_wgrad(_cl(gy), _cl(x), [(0, 0)] * 9, [(ky - 1, kx - 1) for ky, kx in taps], (H, H), g_scale=d, x_scale=s)
    xs = tf32_round(x * s[:, :, None, None]).double().requires_grad_(False)
    gs = tf32_round(gy * d[:, :, None, None]).double()
    w = torch.zeros(Cout, Cin, 3, 3, dtype=torch.float64, device='cuda', requires_grad=True)
    ref, = torch.autograd.grad(F.conv2d(xs, w, padding=1), w, gs)
    got = dw.reshape(3, 3, Cout, Cin).permute(2, 3, 0, 1)
    assert rel_err(got, ref) < 2e-5
    ref32, = torch.autograd.grad(F.conv2d((x * s[:, :, None, None]).double(), w, padding=1), w, (gy * d[:, :, None, None]).double())
    assert rel_err(got, ref32) < 1e-3


def test_wgrad_transposed_stride2():
    g_ = torch.Generator().manual_seed(11)
    N, Cin, Cout, h = 2, 64, 64, 9
    x = torch.randn(N, Cin, h, h, generator=g_).cuda()
    du = torch.randn(N, Cout, 2 * h + 1, 2 * h + 1, generator=g_).cuda()
    taps = C.TAPS_3x3
    dw = C.igemm_wgrad(_cl(du), _cl(x), taps, [(0, 0)] * 9, (h, h), g_stride=2)
    wT = torch.zeros(Cin, Cout, 3, 3, dtype=torch.float64, device='cuda', requires_grad=True)
    ref, = torch.autograd.grad(F.conv_transpose2d(tf32_round(x).double(), wT, stride=2), wT, tf32_round(du).double())
    got = dw.reshape(3, 3, Cout, Cin).permute(3, 2, 0, 1)       # [Cin, Cout, ky, kx] like the conv_transpose2d weight
    assert rel_err(got, ref) < 2e-5


def test_argument_errors():
    x = torch.randn(1, 24, 8, 8).cuda().contiguous(memory_format=torch.channels_last)
    with pytest.raises(RuntimeError):
        C.igemm_conv(x, torch.zeros(1, 64, 24, device='cuda'), [(0, 0)])        # cin % 32 != 0
    x = torch.randn(1, 32, 8, 8).cuda()
    with pytest.raises(RuntimeError):
        C.igemm_conv(x, torch.zeros(1, 64, 32, device='cuda'), [(0, 0)])        # not channels_last


def test_strided_view_input_and_accumulate():
    """Polyphase views of a larger tensor as conv input (TMA strides) and y += accumulation across launches."""
    g = torch.Generator().manual_seed(21)
    N, Cin, Cout, h = 2, 64, 64, 16
    du = _cl(torch.randn(N, Cin, 2 * h + 1, 2 * h + 1, generator=g).cuda())
    w = torch.randn(Cout, Cin, 3, 3, generator=g).cuda()
    out = None
    for a in (0, 1):
        for b in (0, 1):
            taps = [(ky, kx) for ky in range(a, 3, 2) for kx in range(b, 3, 2)]
            offs = [(ky // 2, kx // 2) for ky, kx in taps]
            wp = C.prep_weights(w, taps)
            view = du[:, :, a::2, b::2]
            if out is None:
                out = C.igemm_conv(view, wp, offs, out_hw=(h, h))
            else:
                C.igemm_conv(view, wp, offs, out_view=out, accumulate=True)
    ref = F.conv2d(tf32_round(du).double(), tf32_round(w).double(), stride=2)
    assert rel_err(out, ref) < 2e-5


def test_prep_weights_pair_equals_gather():
    """sgv_conv_prep_weights_pair (one coalesced pass) == two sgv_conv_prep_weights gathers, bit for bit, incl. mirrored taps."""
    g = torch.Generator().manual_seed(21)
    for O, I, k in ((64, 96, 3), (128, 32, 3), (32, 64, 1)):
        w = torch.randn(O, I, k, k, generator=g).cuda()
        base = [(ky, kx) for ky in range(k) for kx in range(k)]
        fwd = base[::-1] if k == 3 else base                                     # an arbitrary (here: mirrored) order
        a, b = C.prep_weights_pair(w, fwd, base)
        assert torch.equal(a, C.prep_weights(w, fwd)) and torch.equal(b, C.prep_weights(w, base, rows_dim=1, cols_dim=0))


def test_wgrad_into_slots():
    g = torch.Generator().manual_seed(22)
    N, Ci, Co, H = 2, 64, 32, 12
    x = _cl(torch.randn(N, Ci, H, H, generator=g).cuda()); gy = _cl(torch.randn(N, Co, H, H, generator=g).cuda())
    taps, offs = C.conv3x3_taps()
    ref = C.igemm_wgrad(gy, x, [(0, 0)] * 9, offs, (H, W) if (W := H) else None)
    out = torch.zeros(9, Co, Ci, device='cuda')
    perm = [8, 0, 3, 1, 2, 7, 6, 4, 5]
    C.igemm_wgrad(gy, x, [(0, 0)] * 4, offs[:4], (H, H), out=out, slots=perm[:4])
    C.igemm_wgrad(gy, x, [(0, 0)] * 5, offs[4:], (H, H), out=out, slots=perm[4:])
    for t in range(9):
        assert rel_err(out[perm[t]], ref[t]) < 1e-6
