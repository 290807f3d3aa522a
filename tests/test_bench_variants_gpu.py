"""Parity of the EXACT kernel variants `bench.py` launches on BASELINE configs[1] (256^2 synthesis, 32 frames), at extents that select
them — asserted through the variant-query entry points (sgv_conv2d_tf32_variant / sgv_conv2d_wgrad_tf32_variant), not assumed.

  conv_tf32_v3_kernel<256,2,2,4,2,pair>   512->512 @ 32^2, N = 32   (256-column N tile; CTA pair: tcgen05 cta_group::2 MMAs of M = 256)
  conv_tf32_v3_kernel<128,2,3,6,2,pair>   128->128 @ 128^2, 256->256 @ 64^2
  conv_tf32_v3_kernel<64,4,2,6,2,pair>    64->64 @ 256^2            (16 x 32-pixel tiles, the roofline entry of bench.py)
  the stride-2 data gradient          64 -> 128 channels, 257^2 -> 128^2
  wgrad_tf32_v2_kernel<128,5>, wgrad_tf32_s64_kernel (64 output channels)
  fir_nhwc_tma44 with the fused epilogue on [N,257,257,64]
and the whole 256^2 network (forward + backward, N = 2 frames) and the 1024^2 network (forward, N = 1) against the CPU oracle.

Bars: vs an fp64 contraction of the SAME TF32-rounded operands 2e-5 (accumulation order only); vs the true fp32 operands 1e-3
(north_star); in tf32x3 mode 5e-5 per contraction (7e-6 ... 3e-5 measured, growing with K) and 5e-4 for a whole network image (2.3e-4 measured)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err
from oracle import ops_ref, synthesis_ref as sr
from stylegan_v_b200 import conv as C
from stylegan_v_b200 import plugin, precision
from stylegan_v_b200.ops import upfirdn2d as U
from stylegan_v_b200.synthesis import SynthesisNetwork

pytestmark = pytest.mark.gpu


def tf32_round(t):
    i = t.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32)


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def _modulated_case(N, Cin, Cout, H, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, Cin, H, H, generator=g).cuda()
    w = torch.randn(Cout, Cin, 3, 3, generator=g).cuda()
    s = (torch.randn(N, Cin, generator=g) + 1).cuda()
    d = (torch.rand(N, Cout, generator=g) + 0.5).cuda() / np.sqrt(Cin * 9)
    b = torch.randn(Cout, generator=g).cuda()
    return x, w, s, d, b


def _ref_layer(x, w, s, d, b, rounded):
    xs = x * s[:, :, None, None]
    if rounded:
        xs, w = tf32_round(xs), tf32_round(w)
    y = F.conv2d(xs.double(), w.double(), padding=1) * d.double()[:, :, None, None] + b.double()[None, :, None, None]
    return F.leaky_relu(y, 0.2) * np.sqrt(2)


# (N, Cin, Cout, H) -> expected (bn, mh, cluster)
BENCH_FWD = [((32, 512, 512, 32), (256, 2, 2)), ((8, 128, 128, 128), (128, 2, 2)), ((8, 64, 64, 256), (64, 4, 2)), ((32, 256, 256, 64), (128, 2, 2))]


@pytest.mark.parametrize('shape,variant', BENCH_FWD)
def test_forward_variants_of_the_benchmark(shape, variant):
    N, Cin, Cout, H = shape
    x, w, s, d, b = _modulated_case(N, Cin, Cout, H, seed=H)
    taps, offs = C.conv3x3_taps()
    wp = C.prep_weights(w, taps, x3=False)
    kw = dict(a_scale=s, o_scale=d, bias=b, act='lrelu', gain=float(np.sqrt(2)))
    v = C.igemm_conv(_cl(x), wp, offs, query=True, **kw)
    assert (v['kernel'], v['bn'], v['mh'], v['cluster'], v['cta_pair']) == (3,) + variant + (1,), v       # clusters of 2 issue tcgen05 cta_group::2 MMAs
    y = C.igemm_conv(_cl(x), wp, offs, **kw)
    assert rel_err(y, _ref_layer(x, w, s, d, b, rounded=True)) < 2e-5
    ref = _ref_layer(x, w, s, d, b, rounded=False)
    assert rel_err(y, ref) < 1e-3
    y3 = C.igemm_conv(_cl(x), C.prep_weights(w, taps, x3=True), offs, **kw)
    # tf32x3: products exact to ~2^-22; the remaining error is the tensor core's fp32 accumulation (not round-to-nearest), growing with the
    # contraction length: measured 7e-6 (K = 576), 1.1e-5 (1152), 1.8e-5 (2304), 3.0e-5 (4608) on the B200
    assert rel_err(y3, ref) < 5e-5


def test_stride2_data_gradient_variant_of_the_benchmark():
    """b256.conv0 backward: gradient of the (2h+1)^2 transposed-conv output [N,64,257,257] -> dx [N,128,128,128] (TMA element strides,
    four parity classes) with the styles epilogue and the fused d(styles) reduction."""
    g = torch.Generator().manual_seed(2)
    N, Cg, Cx, h = 8, 64, 128, 128
    du = tf32_round(torch.randn(N, Cg, 2 * h + 1, 2 * h + 1, generator=g).cuda())        # pre-rounded like the FIR adjoint's output (a_ready)
    w = torch.randn(Cg, Cx, 3, 3, generator=g).cuda()                                    # layer weight [O = Cg, I = Cx]
    s = (torch.randn(N, Cx, generator=g) + 1).cuda()
    x = torch.randn(N, Cx, h, h, generator=g).cuda()
    wp = C.prep_weights(w, C.TAPS_3x3, rows_dim=1, cols_dim=0, x3=False)
    ds = torch.zeros(N, Cx, device='cuda')
    v = C.igemm_conv(_cl(du), wp, C.TAPS_3x3, out_hw=(h, h), in_stride=2, o_scale=s, a_ready=True, query=True)
    assert v['kernel'] == 3 and v['bn'] == 128 and v['cluster'] == 2 and v['cta_pair'] == 1, v
    dx = C.igemm_conv(_cl(du), wp, C.TAPS_3x3, out_hw=(h, h), in_stride=2, o_scale=s, a_ready=True, red_x=_cl(x), red_out=ds)
    raw = F.conv2d(du.double(), tf32_round(w).double().transpose(0, 1), stride=2)
    assert rel_err(dx, raw * s.double()[:, :, None, None]) < 2e-5
    assert rel_err(ds, (raw * x.double()).sum(dim=[2, 3])) < 1e-4


@pytest.mark.parametrize('shape,nt', [((8, 128, 128, 128), 128), ((4, 64, 64, 256), 64), ((32, 512, 512, 32), 128)])
def test_weight_gradient_variants_of_the_benchmark(shape, nt):
    N, Cin, Cout, H = shape
    g_ = torch.Generator().manual_seed(H)
    x = torch.randn(N, Cin, H, H, generator=g_).cuda()
    gy = tf32_round(torch.randn(N, Cout, H, H, generator=g_).cuda())       # the bench path hands the gradient over pre-scaled and pre-rounded (g_ready)
    s = (torch.rand(N, Cin, generator=g_) + 0.5).cuda()
    offs = [(ky - 1, kx - 1) for ky, kx in C.TAPS_3x3]
    v = C.igemm_wgrad(_cl(gy), _cl(x), [(0, 0)] * 9, offs, (H, H), x_scale=s, g_ready=True, x3=False, query=True)
    # 64 output channels: the stacked-M kernel (wgrad_tf32_s64.cu, kernel 3); otherwise the grouped-tap kernel (kernel 2)
    # ... with CTA pairs (tcgen05 cta_group::2, 6 stages) when there are two output-channel tiles to pair
    assert (v['kernel'], v['nt'], v['stages']) == ((3, 64, 5) if Cout == 64 else (2, nt, 6 if Cout % 256 == 0 else 5)), v
    dw = C.igemm_wgrad(_cl(gy), _cl(x), [(0, 0)] * 9, offs, (H, H), x_scale=s, g_ready=True, x3=False)
    got = dw.reshape(3, 3, Cout, Cin).permute(2, 3, 0, 1)
    w = torch.zeros(Cout, Cin, 3, 3, dtype=torch.float64, device='cuda', requires_grad=True)
    xs = x * s[:, :, None, None]
    ref_same, = torch.autograd.grad(F.conv2d(tf32_round(xs).double(), w, padding=1), w, gy.double())
    assert rel_err(got, ref_same) < 5e-5          # split-K partial sums are added with fp32 atomics in arbitrary order
    ref, = torch.autograd.grad(F.conv2d(xs.double(), w, padding=1), w, gy.double())
    assert rel_err(got, ref) < 1e-3
    dw3 = C.igemm_wgrad(_cl(gy), _cl(x), [(0, 0)] * 9, offs, (H, H), x_scale=s, x3=True)
    # K = N*H*W = 32k ... 524k products per output, accumulated in fp32 (TMEM, then fp32 atomics across the K splits): the products are exact to
    # ~2^-22, the fp32 SUM is what remains (measured 1.4e-5 / 2.4e-5; cuDNN's fp32 kernels accumulate in fp32 too).  Small-K cases hold 1e-5
    # (tests/test_precision_gpu.py).
    assert rel_err(dw3.reshape(3, 3, Cout, Cin).permute(2, 3, 0, 1), ref) < 5e-5


def test_fir_tma_kernel_with_epilogue_at_benchmark_extent():
    """[N,257,257,64] -> [N,256,256,64] channels_last, the FIR of b256.conv0 with its dcoefs / noise / bias / lrelu epilogue:
    bit-exact against the C port of the reference kernel followed by the unfused fp32 op sequence."""
    g = torch.Generator().manual_seed(4)
    N, Cc, H = 2, 64, 257
    f = U.setup_filter([1, 3, 3, 1])
    x = torch.randn(N, Cc, H, H, generator=g)
    scale = torch.rand(N, Cc, generator=g) + 0.5
    bias = torch.randn(Cc, generator=g)
    noise = torch.randn(N, 1, H - 1, H - 1, generator=g)
    y = plugin.upfirdn2d(_cl(x.cuda()), f.cuda(), 1, 1, 1, 1, 1, 1, 1, 1, False, 4.0,
                         epilogue=dict(scale=scale.cuda(), noise=noise.cuda(), bias=bias.cuda(), act='lrelu', alpha=0.2, gain=float(np.sqrt(2))))
    o = ops_ref.upfirdn2d_ref(x, f, padding=1, gain=4)
    o = o * scale[:, :, None, None]
    o = o + noise
    o = o + bias[None, :, None, None]
    o = torch.where(o > 0, o, o * 0.2) * np.float32(np.sqrt(2))
    assert y.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(y.cpu(), o), float((y.cpu() - o).abs().max())


def _net_and_inputs(res, channel_base, B, seed=0):
    cfg = sr.SynthesisConfig(img_resolution=res, channel_base=channel_base)
    P = sr.init_params(cfg, seed=seed)
    net = SynthesisNetwork.from_config(cfg)
    sd = net.state_dict()
    sd.update({k: v for k, v in P.items() if k in sd})
    net.load_state_dict(sd)
    g = torch.Generator().manual_seed(seed + 1)
    ws = torch.randn(B, cfg.num_ws, cfg.w_dim, generator=g)
    t = torch.tensor([[0.0], [37.25]])[:B]
    mz = torch.randn(B, sr.max_traj_len(cfg, 1023.0), cfg.motion_z_dim, generator=g)
    return cfg, P, net.cuda(), ws, t, mz


def test_256_network_forward_backward_vs_cpu_oracle():
    """BASELINE configs[1]'s network (256^2, fmaps 0.5, random init) on 2 frames: image and parameter gradients of the fused path against
    the CPU oracle (torch fp32).  Default mode: TF32-class bars (image measured ~1.5e-3: 14 conv layers x <= 1e-3 each);
    tf32x3 mode: fp32-grade bars."""
    cfg, P, net, ws, t, mz = _net_and_inputs(256, 16384, 2)
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    # The motion codes are computed ONCE (CPU oracle) and fed to both sides: the Fourier features take sin / cos of arguments of several hundred
    # radians, where one fp32 ulp of the ARGUMENT (FMA contraction on the GPU vs separate roundings on the CPU) already moves the feature by ~5e-5 —
    # a property of the fp32 conditioning of motion.py:198-212 that would mask the contraction error this test is about.  The time encoder itself is
    # compared in tests/test_train_aux_gpu.py and, below, as a whole against the oracle at its own (2e-4) bar.
    with torch.no_grad():
        mv = sr.motion_encoder(P, cfg, t, mz)
        mv_gpu = net.motion_encoder(t.cuda(), motion_z=mz.cuda())['motion_v']
    assert rel_err(mv_gpu, mv) < 5e-4
    ref = sr.synthesis_forward(Pg, cfg, ws, t, motion_v=mv, fused_modconv=False)
    gen = torch.Generator().manual_seed(9)
    dimg = torch.randn(ref.shape, generator=gen)
    names = ['b256.conv1.weight', 'b256.conv0.weight', 'b128.conv1.weight', 'b64.conv0.weight', 'b32.conv1.weight', 'b8.conv0.weight', 'b4.conv1.weight',
             'b64.conv1.affine.weight', 'b256.torgb.weight', 'b256.conv1.bias']
    gref = torch.autograd.grad(ref, [Pg[n] for n in names], dimg)
    params = dict(net.named_parameters())
    for mode, bar_img, bar_grad in (('tf32', 3e-3, None), ('tf32x3', 5e-4, 1e-2)):      # measured: image 1.5e-3 / 2.3e-4, weight gradients (tf32x3) 2.5e-3 ... 5.5e-3: the backward chain of 14 layers compounds the per-contraction error
        with precision.precision(mode):
            img = net(ws.cuda(), t.cuda(), motion_v=mv.cuda())
            grads = torch.autograd.grad(img, [params[n] for n in names], dimg.cuda())
        e = rel_err(img, ref)
        assert e < bar_img, (mode, e)
        for n, a, r in zip(names, grads, gref):
            ge = rel_err(a, r)
            cos = float(F.cosine_similarity(a.flatten().double().cpu(), r.flatten().double(), dim=0))
            if bar_grad is not None:
                # tf32x3: 20x closer than the TF32 mode, but not bit-level fp32 — per-contraction errors (~1e-5, tensor-core accumulation)
                # compound along the 14-layer backward chain and the test batch is 2 frames, so a low-resolution layer's weight gradient is a sum
                # of only 32-512 terms per element: measured 2.5e-3 (b256) ... 5.5e-3 (b32) ... 3.6e-2 (b8.conv0).  Bars: 1e-2 down to 32^2, 6e-2 +
                # cosine 0.999 below.
                low_res = any(n.startswith(f'b{r}.') for r in (4, 8, 16))
                assert ge < (6e-2 if low_res else 1e-2) and cos > 0.999, (mode, n, ge, cos)
            else:       # TF32 forward flips a few leaky-ReLU slopes (tests/test_synthesis_gpu.py docstring): direction + coarse norm bar
                assert cos > 0.998 and ge < 6e-2, (mode, n, ge, cos)


def test_1024_network_forward_vs_cpu_oracle():
    """BASELINE configs[4]'s network (1024^2, fmaps 1) on 1 frame against the CPU oracle, both precision modes."""
    cfg, P, net, ws, t, mz = _net_and_inputs(1024, 32768, 1, seed=3)
    with torch.no_grad():
        mv = sr.motion_encoder(P, cfg, t, mz)          # shared motion codes: see test_256_network_forward_backward_vs_cpu_oracle
        ref = sr.synthesis_forward(P, cfg, ws, t, motion_v=mv, fused_modconv=False)
        for mode, bar in (('tf32', 3e-3), ('tf32x3', 5e-4)):                    # measured 2.3e-4 in tf32x3 mode
            with precision.precision(mode):
                img = net(ws.cuda(), t.cuda(), motion_v=mv.cuda())
            e = rel_err(img, ref)
            assert e < bar, (mode, e)
