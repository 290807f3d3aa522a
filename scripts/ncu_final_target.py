"""One launch of each hot kernel at its config-2 shape (target for `ncu --set full`, see scripts/run_r1aw.sh)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stylegan_v_b200 import conv as C, plugin
from stylegan_v_b200.ops import upfirdn2d as U
N = 32
dev = 'cuda'
cl = lambda t: t.contiguous(memory_format=torch.channels_last)
taps, offs = C.conv3x3_taps()
f = U.setup_filter([1, 3, 3, 1], device=dev)
# FIR of the b256 up layer with the fused demod/bias/lrelu epilogue (channels_last, TMA-fed)
u = cl(torch.randn(N, 64, 257, 257, device=dev)); sc = torch.rand(N, 64, device=dev) + 0.5; bi = torch.randn(64, device=dev)
for _ in range(2):
    plugin.upfirdn2d(u, f, 1, 1, 1, 1, 1, 1, 1, 1, False, 4.0, epilogue=dict(scale=sc, bias=bi, act='lrelu', alpha=0.2, gain=1.414, clamp=None))
del u
for (ci, co, r) in ((128, 128, 128), (512, 512, 32), (64, 64, 256)):
    x = cl(torch.randn(N, ci, r, r, device=dev)); g = cl(torch.randn(N, co, r, r, device=dev))
    w = torch.randn(co, ci, 3, 3, device=dev); s = torch.rand(N, ci, device=dev) + 0.5; d = torch.rand(N, co, device=dev) + 0.5
    wf, wd = C.prep_weights_pair(w, taps, taps)
    b = torch.zeros(co, device=dev)
    for _ in range(2):
        y = C.igemm_conv(x, wf, offs, a_scale=s, o_scale=d, bias=b, act='lrelu', gain=1.4)
        C.igemm_wgrad(g, x, [(0, 0)] * 9, offs, (r, r), x_scale=s, g_ready=True)
        C.act_bwd(g, y, b, 'lrelu', 1.4, True, True, oscale=d)
    torch.cuda.synchronize()
    del x, g, y
