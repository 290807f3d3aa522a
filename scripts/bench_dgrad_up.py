"""Times the stride-2 data-gradient conv of the up layers (config-2 shapes) — compare SGV_CONV_NO_V3=1 (per-tap kernel) with the default."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stylegan_v_b200 import conv as C
from bench_conv import timeit
N = 32
for name, cg, cx, h in (('b256.conv0.dgrad', 64, 128, 128), ('b128.conv0.dgrad', 128, 256, 64), ('b64.conv0.dgrad', 256, 512, 32), ('b32.conv0.dgrad', 512, 512, 16)):
    g = torch.randn(N, cg, 2 * h + 1, 2 * h + 1, device='cuda').contiguous(memory_format=torch.channels_last)
    w = torch.randn(cx, cg, 3, 3, device='cuda')
    wp = C.prep_weights(w, C.TAPS_3x3)
    d = torch.rand(N, cg, device='cuda') + 0.5; s = torch.rand(N, cx, device='cuda') + 0.5
    ms = timeit(lambda: C.igemm_conv(g, wp, C.TAPS_3x3, out_hw=(h, h), in_stride=2, a_scale=d, o_scale=s))
    flops = 2.0 * N * h * h * cg * cx * 9
    print(json.dumps(dict(kernel=name, ms=round(ms, 4), tflops=round(flops / ms / 1e9, 1))), flush=True)
    del g
