"""Micro-benchmark of the tcgen05 implicit-GEMM convolution at the config-2 layer shapes (N=32 frames)."""
import json
import os
import sys
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stylegan_v_b200 import conv as C


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else None
    N = 32
    shapes = [('b32.conv1', 512, 512, 32), ('b64.conv1', 256, 256, 64), ('b128.conv1', 128, 128, 128), ('b256.conv1', 64, 64, 256),
              ('b16.conv1', 512, 512, 16), ('b8.conv1', 512, 512, 8), ('b4.conv1', 1024, 512, 4)]
    taps, offs = C.conv3x3_taps()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    for name, cin, cout, res in shapes:
        if only and only != name and not (only == 'main4' and res >= 32) and not (only == 'small' and res <= 16):
            continue
        x = torch.randn(N, cin, res, res, device='cuda').contiguous(memory_format=torch.channels_last)
        w = torch.randn(cout, cin, 3, 3, device='cuda')
        s = torch.rand(N, cin, device='cuda') + 0.5
        d = torch.rand(N, cout, device='cuda') + 0.5
        b = torch.randn(cout, device='cuda')
        wp = C.prep_weights(w, taps)
        fn = lambda: C.igemm_conv(x, wp, offs, a_scale=s, o_scale=d, bias=b, act='lrelu', gain=1.414)
        ms = timeit(fn)
        flops = 2.0 * N * res * res * cin * cout * 9
        g = torch.randn(N, cout, res, res, device='cuda').contiguous(memory_format=torch.channels_last)
        ms_w = timeit(lambda: C.igemm_wgrad(g, x, [(0, 0)] * 9, offs, (res, res), g_scale=d, x_scale=s))
        del g
        out = dict(kernel=name, cin=cin, cout=cout, res=res, ms=ms, tflops=flops / ms / 1e9, wgrad_ms=ms_w, wgrad_tflops=flops / ms_w / 1e9)
        if not only:
            wcl = w.contiguous(memory_format=torch.channels_last)
            ms_cudnn = timeit(lambda: F.conv2d(x, wcl, padding=1))
            torch.backends.cudnn.allow_tf32 = True
            ms_cudnn_tf32 = timeit(lambda: F.conv2d(x, wcl, padding=1))
            torch.backends.cudnn.allow_tf32 = False
            out.update(cudnn_fp32_ms=ms_cudnn, cudnn_tf32_ms=ms_cudnn_tf32)
        print(json.dumps(out), flush=True)


if __name__ == '__main__':
    main()
