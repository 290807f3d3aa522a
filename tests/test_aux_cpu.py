"""CPU checks of the time-encoder tail and of the fused optimiser step:
  * the oracle restatement of the tail agrees with the full motion-encoder oracle (itself pinned to the reference golden motion_v);
  * the per-element arithmetic the CUDA kernels execute (csrc/aux_math.cuh, compiled here with g++ by tests/host_emul) agrees with
    the oracle — bit-exact where the operation order is fixed, a few ulp of sin/cos/tanh otherwise.
The CUDA launches themselves are checked in tests/test_aux_gpu.py."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from oracle import synthesis_ref as sr, train_ref

EMUL_DIR = os.path.join(ROOT, 'tests', 'host_emul')


@pytest.fixture(scope='module')
def emul():
    out = os.path.join(EMUL_DIR, '_emul.so')
    src = os.path.join(EMUL_DIR, 'aux_emul.cpp')
    hdr = os.path.join(ROOT, 'stylegan_v_b200', 'csrc', 'aux_math.cuh')
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(['g++', '-O1', '-ffp-contract=off', '-shared', '-fPIC', '-x', 'c++', src, '-o', out, '-lm'], check=True)
    return ctypes.CDLL(out)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _tail_inputs(m=37, nf=24, seed=0, tmax=1023.0):
    g = torch.Generator().manual_seed(seed)
    hl = torch.randn(m, 4 * nf, generator=g)
    ar = torch.randn(m, 2 * nf, generator=g)
    t = torch.rand(m, generator=g) * tmax
    t[:6] = torch.tensor([0.0, 16.0, 15.75, 1023.0, 32.0, 7.5])        # block boundaries: remainder 0 and the last frame
    freqs = sr.linspaced_frequencies(nf, 16, 1024).reshape(-1)
    ps = (1024 / (2 * np.pi / freqs)).float()
    return hl, ar, t, freqs, ps


def test_tail_restatement_matches_motion_encoder_oracle():
    g, meta = load_golden('synthesis_tiny.npz')
    cfg = sr.SynthesisConfig(**meta)
    P = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('p:')}
    t, mz = torch.from_numpy(g['t']), torch.from_numpy(g['motion_z'])
    # stacked predictor outputs exactly as the product computes them
    B, Fr = t.shape
    L = sr.max_traj_len(cfg, float(t.max()))
    h = mz[:B, :L, :cfg.motion_z_dim].permute(0, 2, 1)
    h = sr.eqlr_conv1d(h, P['motion_encoder.conv.0.weight'], P['motion_encoder.conv.0.bias'], lr_multiplier=0.01)
    h = sr.eqlr_conv1d(h, P['motion_encoder.conv.1.weight'], P['motion_encoder.conv.1.bias'], lr_multiplier=0.01)
    trajs = h.permute(0, 2, 1)
    d = cfg.motion_z_distance
    li = (t / d).floor().long()
    bi = torch.arange(B).unsqueeze(1).repeat(1, Fr)
    uL, uR = trajs[bi, li].reshape(B * Fr, -1), trajs[bi, li + 1].reshape(B * Fr, -1)
    te = 'motion_encoder.time_encoder.'
    heads = torch.cat([sr.fully_connected(uL, P[te + k + '.weight']) for k in ('periods_predictor', 'phase_predictor', 'aligners_predictor')], dim=1)
    ar = sr.fully_connected(uR, P[te + 'aligners_predictor.weight'])
    freqs = sr.linspaced_frequencies(cfg.time_enc_dim, cfg.min_period_len, cfg.max_period_len)
    ps = cfg.max_period_len / (2 * np.pi / freqs)
    v = train_ref.time_encoder_tail_ref(heads, ar, t.reshape(-1), freqs, ps, d)
    assert torch.equal(v, sr.motion_encoder(P, cfg, t, mz))
    assert float((v - torch.from_numpy(g['motion_v'])).abs().max()) < 1e-5         # the reference's own output


def test_kernel_arithmetic_time_encoder_fwd(emul):
    hl, ar, t, freqs, ps = _tail_inputs()
    m, nf = t.numel(), freqs.numel()
    out = np.zeros([m, 2 * nf], np.float32)
    a = [x.numpy().copy() for x in (hl, ar, t, freqs, ps)]
    emul.emul_time_encoder_fwd(*[_ptr(x) for x in a], _ptr(out), m, nf, ctypes.c_float(16.0))
    ref = train_ref.time_encoder_tail_ref(hl, ar, t, freqs, ps, 16.0)
    # operation order is identical; sin/cos/tanh implementations (glibc vs torch's vectorised) differ by a few ulp, and a 1-ulp tanh
    # difference moves a 800-rad phase by ~1e-4
    assert float((torch.from_numpy(out) - ref).abs().max()) < 5e-4


def test_c_port_of_the_tail_and_kernel_arithmetic_bit_exact(emul):
    """Three statements of the same arithmetic: the torch restatement (pinned to the reference golden above), the plain-C port, and the
    kernel's per-element code compiled for the host.  The C port and the kernel arithmetic share libm, so they must agree BIT FOR BIT
    (operation order, separate roundings, remainder semantics); the torch version differs only by its vectorised sin / cos / tanh."""
    hl, ar, t, freqs, ps = _tail_inputs(m=53, nf=40, seed=7)
    t[6:10] = torch.tensor([1008.0, 1022.99, 0.001, 511.5])
    c_out = train_ref.time_encoder_tail_c(hl, ar, t, freqs, ps, 16.0)
    m, nf = t.numel(), freqs.numel()
    k_out = np.zeros([m, 2 * nf], np.float32)
    a = [x.numpy().copy() for x in (hl, ar, t, freqs, ps)]
    emul.emul_time_encoder_fwd(*[_ptr(x) for x in a], _ptr(k_out), m, nf, ctypes.c_float(16.0))
    assert np.array_equal(c_out.numpy(), k_out)
    ref = train_ref.time_encoder_tail_ref(hl, ar, t, freqs, ps, 16.0)
    assert float((c_out - ref).abs().max()) < 5e-4


def test_kernel_arithmetic_time_encoder_bwd(emul):
    hl, ar, t, freqs, ps = _tail_inputs(seed=1, tmax=200.0)
    m, nf = t.numel(), freqs.numel()
    dout = torch.randn(m, 2 * nf, generator=torch.Generator().manual_seed(5))
    hl64, ar64 = hl.double().requires_grad_(True), ar.double().requires_grad_(True)
    ref = train_ref.time_encoder_tail_ref(hl64, ar64, t.double(), freqs.double(), ps.double(), 16.0)
    # (the fp64 evaluation casts t to float32 inside emb() like the reference does; t values here are fp32-exact anyway)
    ghl, gar = torch.autograd.grad(ref, [hl64, ar64], dout.double())
    dhl = np.zeros([m, 4 * nf], np.float32)
    dar = np.zeros([m, 2 * nf], np.float32)
    a = [x.numpy().copy() for x in (dout, hl, t, freqs, ps)]
    emul.emul_time_encoder_bwd(*[_ptr(x) for x in a], _ptr(dhl), _ptr(dar), m, nf, ctypes.c_float(16.0))
    assert float((torch.from_numpy(dar).double() - gar).abs().max()) < 1e-6
    err = (torch.from_numpy(dhl).double() - ghl).abs().max() / ghl.abs().max()
    assert float(err) < 1e-4, float(err)


@pytest.mark.parametrize('ema', [False, True])
def test_kernel_arithmetic_adam(emul, ema):
    g = torch.Generator().manual_seed(3)
    shapes = [(7, 5), (33,), (4, 3, 3, 3)]
    params = [torch.randn(s, generator=g) for s in shapes]
    emas = [p.clone() + 0.01 for p in params] if ema else None
    ref = train_ref.OptimizerRef(params, emas, lr=0.0025, betas=(0.0, 0.99), eps=1e-8)
    n = sum(p.numel() for p in params)
    p = torch.cat([x.reshape(-1) for x in params]).numpy().copy()
    pe = torch.cat([x.reshape(-1) for x in emas]).numpy().copy() if ema else None
    m, v = np.zeros(n, np.float32), np.zeros(n, np.float32)
    for step in range(1, 6):
        grads = [torch.randn(s, generator=g) * (10.0 ** (step - 3)) for s in shapes]
        grads[0][0, 0] = float('nan')
        grads[1][1] = float('inf')
        grads[1][2] = -3e5                                                       # finite but beyond the clamp
        ref.step(grads, ema_beta=0.998 if ema else None, grad_scale=0.5)
        gf = torch.cat([x.reshape(-1) for x in grads]).numpy().copy()
        emul.emul_adam_ema(_ptr(p), _ptr(gf), _ptr(m), _ptr(v), _ptr(pe), ctypes.c_int64(n), *[ctypes.c_float(x) for x in
                           (0.0025, 0.0, 0.99, 1e-8, 0.998 if ema else 0.0, 0.5, 1e5)], step, 1)
        assert not gf.any()                                                      # zero_grad
        want = torch.cat([x.detach().reshape(-1) for x in ref.params])
        assert float((torch.from_numpy(p) - want).abs().max()) < 2e-6, step
        if ema:
            want_e = torch.cat([x.reshape(-1) for x in ref.ema])
            assert float((torch.from_numpy(pe) - want_e).abs().max()) < 1e-6, step
