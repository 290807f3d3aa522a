"""GPU parity of the two auxiliary kernels either side of the contraction stack (include/sgv_b200_aux.h):
the Fourier time-encoder tail (fwd + gradient) against the oracle restatement of motion.py:111-115,198-212, and the fused
nan_to_num -> Adam -> EMA step against torch.optim.Adam driven the way training_loop.py:381-400 drives it."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import synthesis_ref as sr, train_ref
from stylegan_v_b200 import _lib
from stylegan_v_b200.optim import FlatModuleState, FusedAdamEMA
from stylegan_v_b200.time_encoder import MotionMappingNetwork, _TimeEncoderTail

pytestmark = pytest.mark.gpu


def _tail_inputs(m, nf, seed, tmax):
    g = torch.Generator().manual_seed(seed)
    hl = torch.randn(m, 4 * nf, generator=g)
    ar = torch.randn(m, 2 * nf, generator=g)
    t = torch.rand(m, generator=g) * tmax
    t[:6] = torch.tensor([0.0, 16.0, 15.75, min(tmax, 1023.0), 32.0, 7.5])
    freqs = sr.linspaced_frequencies(nf, 16, 1024).reshape(-1)
    ps = (1024 / (2 * np.pi / freqs)).float()
    return hl, ar, t, freqs, ps


@pytest.mark.parametrize('m,nf', [(48, 256), (7, 16), (96, 100)])
def test_time_encoder_tail_forward(cuda, m, nf):
    hl, ar, t, freqs, ps = _tail_inputs(m, nf, 0, 1023.0)
    n0 = _lib.launch_count()
    out = _TimeEncoderTail.apply(hl.to(cuda), ar.to(cuda), t.to(cuda), freqs.to(cuda), ps.to(cuda), 16.0)
    assert _lib.launch_count() == n0 + 1
    ref = train_ref.time_encoder_tail_ref(hl, ar, t, freqs, ps, 16.0)
    # same operation order; libdevice vs host sin/cos/tanh differ by a few ulp and one ulp of tanh moves an 800-rad phase by ~1e-4
    assert float((out.cpu() - ref).abs().max()) < 5e-4


def test_time_encoder_tail_backward(cuda):
    hl, ar, t, freqs, ps = _tail_inputs(48, 64, 1, 200.0)
    dout = torch.randn(48, 128, generator=torch.Generator().manual_seed(5))
    hl64, ar64 = hl.double().requires_grad_(True), ar.double().requires_grad_(True)
    ref = train_ref.time_encoder_tail_ref(hl64, ar64, t.double(), freqs.double(), ps.double(), 16.0)
    ghl, gar = torch.autograd.grad(ref, [hl64, ar64], dout.double())
    a, b = hl.to(cuda).requires_grad_(True), ar.to(cuda).requires_grad_(True)
    out = _TimeEncoderTail.apply(a, b, t.to(cuda), freqs.to(cuda), ps.to(cuda), 16.0)
    out.backward(dout.to(cuda))
    assert float((b.grad.cpu().double() - gar).abs().max()) < 1e-6
    assert rel_err(a.grad, ghl) < 1e-4


def test_motion_encoder_vs_reference_golden(cuda):
    """Whole motion encoder (conv1d trajectory + gather + fused tail) against motion_v minted from the unmodified reference."""
    g, meta = load_golden('synthesis_tiny.npz')
    cfg = sr.SynthesisConfig(**meta)
    net = MotionMappingNetwork(cfg.motion_z_dim, cfg.motion_v_dim, cfg.motion_kernel_size, cfg.motion_z_distance, cfg.time_enc_dim,
                               cfg.min_period_len, cfg.max_period_len, cfg.max_num_frames)
    sd = {k[len('p:motion_encoder.'):]: torch.from_numpy(g[k]) for k in g.files if k.startswith('p:motion_encoder.')}
    net.load_state_dict(sd)
    net = net.to(cuda)
    v = net(torch.from_numpy(g['t']).to(cuda), motion_z=torch.from_numpy(g['motion_z']).to(cuda))['motion_v']
    assert float((v.cpu() - torch.from_numpy(g['motion_v'])).abs().max()) < 1e-3


def _make(shapes, seed, dev):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter(torch.randn(s, generator=g).to(dev)) for s in shapes]


@pytest.mark.parametrize('ema', [False, True])
def test_fused_adam_ema_vs_torch_adam(cuda, ema):
    shapes = [(64, 32, 3, 3), (33,), (7, 5), (128, 128), (1,)]           # odd sizes: padded offsets + scalar tail
    params = _make(shapes, 0, cuda)
    emas = [torch.nn.Parameter(p.detach().clone() + 0.01) for p in params] if ema else None
    ref = train_ref.OptimizerRef([p.detach().cpu() for p in params], [p.detach().cpu() for p in emas] if ema else None,
                                 lr=0.0025, betas=(0.0, 0.99), eps=1e-8)
    st = FlatModuleState(params, emas)
    assert all(p.data_ptr() % 256 == 0 and p.grad.data_ptr() % 256 == 0 for p in params)
    opt = FusedAdamEMA(st, lr=0.0025, betas=(0.0, 0.99), eps=1e-8)
    g = torch.Generator().manual_seed(1)
    for step in range(1, 6):
        grads = [torch.randn(s, generator=g) * (10.0 ** (step - 3)) for s in shapes]
        grads[1][0] = float('nan')
        grads[1][1] = float('inf')
        grads[1][2] = -3e5
        for p, gr in zip(params, grads):
            p.grad.copy_(gr.to(cuda))
        n0 = _lib.launch_count()
        opt.step(ema_beta=0.998 if ema else None, zero_grad=True, grad_scale=0.5)
        assert _lib.launch_count() == n0 + 1
        ref.step(grads, ema_beta=0.998 if ema else None, grad_scale=0.5)
        assert not st.grad.any()
        for p, q in zip(params, ref.params):
            assert float((p.detach().cpu() - q.detach()).abs().max()) < 2e-6, step
        if ema:
            for p, q in zip(emas, ref.ema):
                assert float((p.detach().cpu() - q).abs().max()) < 1e-6, step


def test_fused_adam_device_step_counter_in_cuda_graph(cuda):
    """The launch pair (advance counter, update) captured once is the next optimiser step on every replay."""
    shapes = [(256, 256), (100,)]
    params = _make(shapes, 2, cuda)
    ref = train_ref.OptimizerRef([p.detach().cpu() for p in params], None, lr=0.01, betas=(0.5, 0.99), eps=1e-8)
    st = FlatModuleState(params)
    opt = FusedAdamEMA(st, lr=0.01, betas=(0.5, 0.99), eps=1e-8, device_step=True)
    # first launches of these kernels (module load, occupancy query) must not happen inside a stream capture: warm up on a throwaway state
    FusedAdamEMA(FlatModuleState(_make([(8,)], 3, cuda)), device_step=True).step()
    torch.cuda.synchronize()
    gstat = [torch.randn(s, generator=torch.Generator().manual_seed(7)) for s in shapes]
    for p, gr in zip(params, gstat):
        p.grad.copy_(gr.to(cuda))
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(graph, stream=s):
            opt.step()
    for _ in range(4):
        graph.replay()
        ref.step(gstat)
    torch.cuda.synchronize()
    assert int(opt.step_count.item()) == 4
    for p, q in zip(params, ref.params):
        assert float((p.detach().cpu() - q.detach()).abs().max()) < 5e-6


def test_fused_adam_rejects_cpu_and_misaligned(cuda):
    with pytest.raises(AssertionError):
        FusedAdamEMA(FlatModuleState([torch.nn.Parameter(torch.zeros(4))]))
    q = _lib.AdamParams()
    buf = torch.zeros(64, device=cuda)
    q.param, q.grad, q.exp_avg, q.exp_avg_sq = buf.data_ptr() + 4, buf.data_ptr(), buf.data_ptr(), buf.data_ptr()
    q.numel, q.step, q.beta2 = 8, 1, 0.99
    assert _lib.lib().sgv_adam_ema_step(q, None) == 1           # SGV_ERR_INVALID


def test_fused_adam_bandwidth_shape(cuda):
    """Synthesis-network-sized flat state (31.5 M parameters): one launch, finite results; prints the achieved GB/s."""
    n = 31_500_000
    p = torch.nn.Parameter(torch.randn(n, device=cuda))
    e = torch.nn.Parameter(p.detach().clone())
    st = FlatModuleState([p], [e])
    st.grad.normal_()
    opt = FusedAdamEMA(st, lr=0.0025)
    for _ in range(3):
        opt.step(ema_beta=0.998)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        opt.step(ema_beta=0.998)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f'fused adam+ema: {ms:.3f} ms, {opt.algorithmic_bytes(True, False) / ms / 1e6:.0f} GB/s')
    assert torch.isfinite(st.param).all() and torch.isfinite(st.ema).all()
