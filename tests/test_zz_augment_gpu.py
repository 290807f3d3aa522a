"""AugmentPipe on CUDA (FIR kernels of libsgv_b200 for the 12-tap up / down passes, library sampler and grouped convs) against the same
module on CPU, in the deterministic debug-percentile mode (random draws would come from different generators on the two devices).
(Round 1 shipped this test opt-in because it had never run on a GPU; it is a normal member of the GPU suite now.)"""
import pytest
import torch

from conftest import rel_err
from stylegan_v_b200 import _lib
from stylegan_v_b200.augment import AugmentPipe

pytestmark = pytest.mark.gpu

BGC = dict(xflip=1, rotate90=1, xint=1, scale=1, rotate=1, aniso=1, xfrac=1, brightness=1, contrast=1, lumaflip=1, hue=1, saturation=1)


@pytest.mark.parametrize('shape,pct,extra', [([3, 3, 32, 32], 0.7, {}), ([2, 9, 24, 40], 0.3, {}), ([2, 3, 32, 32], 0.6, dict(imgfilter=1, cutout=1))])
def test_pipe_cuda_matches_cpu_in_debug_mode(cuda, shape, pct, extra):
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False                       # the grouped band-filter convolutions take the library path
    try:
        pipe = AugmentPipe(**BGC, **extra)
        x = torch.randn(shape, generator=torch.Generator().manual_seed(3))
        dy = torch.randn(shape, generator=torch.Generator().manual_seed(4))
        xc = x.clone().requires_grad_(True)
        yc = pipe(xc, debug_percentile=pct)
        gc, = torch.autograd.grad(yc, [xc], dy)
        n0 = _lib.launch_count()
        xg = x.to(cuda).requires_grad_(True)
        yg = pipe.to(cuda)(xg, debug_percentile=pct)
        gg, = torch.autograd.grad(yg, [xg], dy.to(cuda))
        assert _lib.launch_count() > n0                            # the FIR passes ran on our kernels
        assert rel_err(yg, yc) < 2e-4 and rel_err(gg, gc) < 2e-4
    finally:
        torch.backends.cudnn.allow_tf32 = old
