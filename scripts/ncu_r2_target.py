"""Target for `ncu --set full` (round 2): the hot kernels at their BASELINE configs[1] shapes, in a FIXED order so that scripts/ncu_traffic.py
can name each captured launch.  Each kernel is launched twice (the first warms caches / attributes); capture with
  ncu --set full --clock-control none --import-source on -k regex:'conv_tf32_v3|wgrad_tf32_v2|wgrad_tf32_s64|fir_nhwc_tma44' -o gpurun_out/ncu_r2 python scripts/ncu_r2_target.py
Order of the captured names (second launch of each pair is the one summarised): see ORDER below."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stylegan_v_b200 import conv as C, plugin
from stylegan_v_b200.ops import upfirdn2d as U

ORDER = ['b32.conv1', 'b64.conv1', 'b128.conv1', 'b256.conv1', 'wgrad.b128.conv1', 'wgrad.b256.conv1', 'wgrad.b32.conv1', 'fir_nhwc_tma44']
if __name__ == '__main__':
    N, dev = 32, 'cuda'
    cl = lambda t: t.contiguous(memory_format=torch.channels_last)
    taps, offs = C.conv3x3_taps()
    for (ci, co, r) in ((512, 512, 32), (256, 256, 64), (128, 128, 128), (64, 64, 256)):
        x = cl(torch.randn(N, ci, r, r, device=dev))
        wf = C.prep_weights(torch.randn(co, ci, 3, 3, device=dev), taps, x3=False)
        s = torch.rand(N, ci, device=dev) + 0.5; d = torch.rand(N, co, device=dev) + 0.5; b = torch.zeros(co, device=dev)
        for _ in range(2):
            C.igemm_conv(x, wf, offs, a_scale=s, o_scale=d, bias=b, act='lrelu', gain=1.4)
        torch.cuda.synchronize()
        del x
    for (ci, co, r) in ((128, 128, 128), (64, 64, 256), (512, 512, 32)):
        x = cl(torch.randn(N, ci, r, r, device=dev)); g = cl(torch.randn(N, co, r, r, device=dev)); s = torch.rand(N, ci, device=dev) + 0.5
        for _ in range(2):
            C.igemm_wgrad(g, x, [(0, 0)] * 9, offs, (r, r), x_scale=s, g_ready=True, x3=False)
        torch.cuda.synchronize()
        del x, g
    f = U.setup_filter([1, 3, 3, 1], device=dev)
    u = cl(torch.randn(N, 64, 257, 257, device=dev)); sc = torch.rand(N, 64, device=dev) + 0.5; bi = torch.randn(64, device=dev)
    for _ in range(2):
        plugin.upfirdn2d(u, f, 1, 1, 1, 1, 1, 1, 1, 1, False, 4.0, epilogue=dict(scale=sc, bias=bi, act='lrelu', alpha=0.2, gain=1.414, clamp=None))
    torch.cuda.synchronize()
