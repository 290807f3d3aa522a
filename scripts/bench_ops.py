"""Micro-benchmark of the HBM-bound kernels at BASELINE config-2 / config-5 shapes (CUDA events, L2 flushed by
cycling through buffers larger than L2).  Prints one JSON object per kernel."""
import json
import sys
import os
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stylegan_v_b200.ops import upfirdn2d as U, bias_act as B
from stylegan_v_b200 import plugin


def timeit(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def main():
    dev = torch.device('cuda')
    peak = 6567.7
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs']
    except Exception:
        pass
    f = U.setup_filter([1, 3, 3, 1], device=dev)
    out = []
    cases = [('fir_b256_nchw', (32, 64, 257, 257), False), ('fir_b256_nhwc', (32, 64, 257, 257), True),
             ('fir_b128_nchw', (32, 128, 129, 129), False), ('fir_b128_nhwc', (32, 128, 129, 129), True),
             ('fir_1024_nchw', (8, 32, 1025, 1025), False)]
    for name, shape, cl in cases:
        x = torch.randn(shape, device=dev)
        if cl:
            x = x.contiguous(memory_format=torch.channels_last)
        y = U.upfirdn2d(x, f, padding=1, gain=4)
        nbytes = (x.numel() + y.numel()) * 4
        med, best = timeit(lambda: plugin.upfirdn2d(x, f, 1, 1, 1, 1, 1, 1, 1, 1, False, 4.0))
        out.append(dict(kernel=name, shape=list(shape), ms=med, ms_best=best, gbs=nbytes / med / 1e6, frac_of_measured_peak=nbytes / med / 1e6 / peak))
        del x, y
    # D-side geometries
    for name, shape, kw in [('fir_down2_nchw', (32, 64, 256, 256), dict(down=2, padding=1)), ('fir_blur_pad2_nchw', (32, 64, 256, 256), dict(padding=2)),
                            ('fir_up2_img', (32, 3, 128, 128), dict(up=2, padding=[2, 1, 2, 1], gain=4))]:
        x = torch.randn(shape, device=dev)
        y = U.upfirdn2d(x, f, **kw)
        nbytes = (x.numel() + y.numel()) * 4
        med, best = timeit(lambda: U.upfirdn2d(x, f, **kw))
        out.append(dict(kernel=name, shape=list(shape), ms=med, ms_best=best, gbs=nbytes / med / 1e6, frac_of_measured_peak=nbytes / med / 1e6 / peak))
        del x, y
    # the same D-side geometries channels_last (what the fused discriminator runs), and the adjoint of the decimating FIR (zero insertion x2)
    for name, shape, kw in [('fir_down2_nhwc', (32, 64, 256, 256), dict(down=2, padding=1)), ('fir_blur_pad2_nhwc', (32, 64, 256, 256), dict(padding=2)),
                            ('fir_up2_adjoint_nhwc', (32, 64, 128, 128), dict(up=2, padding=[2, 1, 2, 1], flip_filter=True)),
                            ('fir_up2_adjoint_nhwc_128ch', (32, 128, 64, 64), dict(up=2, padding=[2, 1, 2, 1], flip_filter=True))]:
        x = torch.randn(shape, device=dev).contiguous(memory_format=torch.channels_last)
        y = U.upfirdn2d(x, f, **kw)
        nbytes = (x.numel() + y.numel()) * 4
        med, best = timeit(lambda: U.upfirdn2d(x, f, **kw))
        out.append(dict(kernel=name, shape=list(shape), ms=med, ms_best=best, gbs=nbytes / med / 1e6, frac_of_measured_peak=nbytes / med / 1e6 / peak))
        del x, y
    for name, shape, cl in [('bias_act_b256_nchw', (32, 64, 256, 256), False), ('bias_act_b256_nhwc', (32, 64, 256, 256), True)]:
        x = torch.randn(shape, device=dev)
        if cl:
            x = x.contiguous(memory_format=torch.channels_last)
        b = torch.randn(shape[1], device=dev)
        nbytes = 2 * x.numel() * 4
        med, best = timeit(lambda: B.bias_act(x, b, act='lrelu'))
        out.append(dict(kernel=name, shape=list(shape), ms=med, ms_best=best, gbs=nbytes / med / 1e6, frac_of_measured_peak=nbytes / med / 1e6 / peak))
    # torch copy as the in-process ceiling
    a = torch.empty(32 * 64 * 256 * 256, device=dev); c = torch.empty_like(a)
    med, best = timeit(lambda: c.copy_(a))
    out.append(dict(kernel='torch_copy_537MB', ms=med, gbs=2 * a.numel() * 4 / med / 1e6))
    for o in out:
        print(json.dumps(o))


if __name__ == '__main__':
    main()
