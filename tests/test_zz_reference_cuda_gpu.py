"""libsgv_b200 against THE REFERENCE'S OWN CUDA PLUGINS on identical inputs (north_star: "outputs match the reference's own JIT-compiled ops").

oracle/build_ref.py compiles /root/reference/src/torch_utils/ops/{upfirdn2d,bias_act}.{cpp,cu} UNMODIFIED for sm_100a into oracle/_ref/
(in the build container; the .so files travel to the GPU box).  Both sides expose the same two pybind-style entry points
(`upfirdn2d(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain)`, `bias_act(x, b, xref, yref, dy, grad, dim, act, alpha,
gain, clamp)`), so the comparison is plugin against plugin.

Bars: upfirdn2d and the piecewise-linear activations accumulate in the same order with FMAs on both sides — expected bit-identical, asserted
to 1e-6 of the output range (a wrong tap or index would be O(1)); transcendental activations 2e-3 because the reference is built with
--use_fast_math (bias_act.py:45) and ours uses the accurate functions.
When oracle/_ref holds the built plugins (it does on the GPU box: they travel with the snapshot) a plugin that cannot be loaded or launched
FAILS the test; it is skipped only where the plugins were never built (a checkout without /root/reference and without a prior build()).
The same plugins are timed by bench.py ("beat this kernel": `reference_cuda_kernels` in the bench line)."""
import pytest
import torch

from conftest import rel_err
from oracle import build_ref
from stylegan_v_b200 import plugin

pytestmark = pytest.mark.gpu


def _ref(name):
    import os
    if not os.path.exists(build_ref.plugin_path(name)):
        pytest.skip(f'oracle/_ref/{name} not built (python -m oracle.build_ref in the build container)')
    mod = build_ref.load_plugin(name)          # built but not loadable = a broken comparison, not a skip
    assert mod is not None, f'{build_ref.plugin_path(name)} exists but did not load'
    return mod


def _call_ref(fn, *args):
    out = fn(*args)                            # a launch failure of the reference kernel on this device fails the test
    torch.cuda.synchronize()
    return out


FIR_CASES = [
    # shape, filter taps (outer product of 1-D taps), up, down, (px0, px1, py0, py1), flip, gain, channels_last
    ([2, 8, 33, 33], [1, 3, 3, 1], 1, 1, (1, 1, 1, 1), False, 4.0, False),        # G up-layer FIR (2h+1 -> 2h)
    ([2, 8, 33, 33], [1, 3, 3, 1], 1, 1, (1, 1, 1, 1), False, 4.0, True),
    ([2, 3, 16, 16], [1, 3, 3, 1], 2, 1, (2, 1, 2, 1), False, 4.0, False),        # img upsample2d
    ([2, 16, 32, 32], [1, 3, 3, 1], 1, 1, (2, 2, 2, 2), False, 1.0, False),       # D blur before the stride-2 conv
    ([2, 16, 32, 32], [1, 3, 3, 1], 1, 2, (1, 1, 1, 1), False, 1.0, True),        # D skip down-sampling
    ([1, 4, 20, 24], [1, 3, 3, 1], 1, 1, (2, 2, 2, 2), True, 4.0, False),         # backward of the first case (flipped)
    ([1, 3, 12, 14], [1, 2, 3], 2, 3, (1, 2, 0, 3), False, 1.5, False),           # generic kernel path
    ([1, 3, 12, 14], [1, 3, 3, 1], 1, 1, (-1, 2, 1, -2), False, 1.0, False),      # negative padding = crop
    ([4, 64, 65, 65], [1, 3, 3, 1], 1, 1, (1, 1, 1, 1), False, 4.0, True),        # wide channels_last (TMA kernel on our side)
]


@pytest.mark.parametrize('case', FIR_CASES)
def test_upfirdn2d_vs_reference_cuda_kernel(cuda, case):
    ref = _ref('upfirdn2d_plugin')
    shape, taps, up, down, pad, flip, gain, cl = case
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(shape, generator=g).to(cuda)
    if cl:
        x = x.contiguous(memory_format=torch.channels_last)
    k = torch.tensor(taps, dtype=torch.float32)
    f = torch.outer(k, k)
    f = (f / f.sum()).to(cuda)
    args = (x, f, up, up, down, down, pad[0], pad[1], pad[2], pad[3], flip, gain)
    want = _call_ref(ref.upfirdn2d, *args)
    got = plugin.upfirdn2d(*args)
    assert got.shape == want.shape
    if got.stride() != want.stride():                                               # output layout rule of upfirdn2d.cpp:35 (reported, not asserted, until seen on a GPU)
        print(f'note: output strides differ for {case}: ours {got.stride()} vs reference {want.stride()}')
    assert rel_err(got, want) <= 1e-6
    if not torch.equal(got, want):
        print(f'note: not bit-identical for {case}: max |diff| = {float((got - want).abs().max()):.3e}')


ACTS = {'linear': 1, 'relu': 2, 'lrelu': 3, 'tanh': 4, 'sigmoid': 5, 'elu': 6, 'selu': 7, 'softplus': 8, 'swish': 9}      # bias_act.py:23-33
ALPHA = {'lrelu': 0.2}


@pytest.mark.parametrize('act', list(ACTS))
@pytest.mark.parametrize('cl', [False, True])
def test_bias_act_vs_reference_cuda_kernel(cuda, act, cl):
    ref = _ref('bias_act_plugin')
    g = torch.Generator().manual_seed(ACTS[act])
    x = torch.randn(3, 16, 9, 11, generator=g).to(cuda)
    dy = torch.randn(3, 16, 9, 11, generator=g).to(cuda)
    if cl:
        x, dy = x.contiguous(memory_format=torch.channels_last), dy.contiguous(memory_format=torch.channels_last)
    b = torch.randn(16, generator=g).to(cuda)
    nil = torch.empty([0], device=cuda)
    exact = act in ('linear', 'relu', 'lrelu')
    # clamp only where the mask is decided by stored values: softplus / swish recompute yref inside the gradient kernel (bias_act.cu:113-129), and a
    # fast-math exp can move an element across the clamp threshold on one side only
    alpha, gain, clamp = ALPHA.get(act, 0.0), 1.3, (2.0 if exact else -1.0)
    tol = 1e-6 if exact else 2e-3
    # forward
    a0 = (x, b, nil, nil, nil, 0, 1, ACTS[act], alpha, gain, clamp)
    want = _call_ref(ref.bias_act, *a0)
    got = plugin.bias_act(*a0)
    assert rel_err(got, want) <= tol
    # first-order gradient kernel (grad = 1): xref = x, yref = y as BiasActCudaGrad passes them (bias_act.py:164-170)
    a1 = (dy, b, x, want, nil, 1, 1, ACTS[act], alpha, gain, clamp)
    want1 = _call_ref(ref.bias_act, *a1)
    got1 = plugin.bias_act(*a1)
    assert rel_err(got1, want1) <= tol
    # second order (grad = 2) for the activations that have one (bias_act.py:23-33 has_2nd_grad)
    if act in ('tanh', 'sigmoid', 'elu', 'selu', 'softplus', 'swish'):
        a2 = (dy, b, x, want, dy, 2, 1, ACTS[act], alpha, gain, clamp)
        want2 = _call_ref(ref.bias_act, *a2)
        got2 = plugin.bias_act(*a2)
        assert rel_err(got2, want2) <= 5e-3
