"""AugmentPipe (stylegan_v_b200/augment.py) against outputs of the unmodified reference pipe (augment.py:117-436) for fixed seeds:
same buffers, same random-number stream, same images and input gradients — blit / geometric / colour ('bgc', the reference default),
clip-shaped inputs (video-consistent augmentation), grayscale, image-space filtering, noise, cutout, and the debug-percentile mode."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from stylegan_v_b200.augment import AugmentPipe


def _t(a):
    return torch.from_numpy(np.asarray(a))


def test_buffers_equal_reference():
    g, _ = load_golden('augment_cases.npz')
    pipe = AugmentPipe(xflip=1)
    sd = pipe.state_dict()
    assert set(sd) == {k[4:] for k in g.files if k.startswith('buf:')}
    assert torch.equal(sd['Hz_geom'], _t(g['buf:Hz_geom']))                  # sym6 low-pass through setup_filter
    assert rel_err(sd['Hz_fbank'], _t(g['buf:Hz_fbank'])) < 1e-7             # sym2 octave filter bank


@pytest.mark.parametrize('idx', range(6))
def test_pipe_matches_reference_stream_and_values(idx):
    g, meta = load_golden('augment_cases.npz')
    case = meta[idx]
    pipe = AugmentPipe(**case['kw'])
    pipe.p.copy_(torch.as_tensor(case['p']))
    x = _t(g[case['name'] + ':x']).requires_grad_(True)
    torch.manual_seed(1234)
    y = pipe(x, debug_percentile=case.get('debug_percentile'))
    assert y.shape == x.shape
    assert rel_err(y, _t(g[case['name'] + ':y'])) < 1e-5, case['name']
    dx, = torch.autograd.grad(y, [x], _t(g[case['name'] + ':dy']))
    assert rel_err(dx, _t(g[case['name'] + ':dx'])) < 1e-5, case['name']


def test_identity_when_disabled():
    pipe = AugmentPipe()                                                      # every multiplier 0: nothing is drawn, nothing changes
    x = torch.randn(2, 3, 8, 8)
    state = torch.get_rng_state()
    assert pipe(x) is x and torch.equal(torch.get_rng_state(), state)
