"""Runs the tensor-core kernels once each at config-2 layer shapes (target for `ncu --set full -k regex:...`)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stylegan_v_b200 import conv as C
N = 32
taps, offs = C.conv3x3_taps()
cl = lambda t: t.contiguous(memory_format=torch.channels_last)
for (ci, co, r) in ((512, 512, 32), (64, 64, 256)):
    x = cl(torch.randn(N, ci, r, r, device='cuda')); g = cl(torch.randn(N, co, r, r, device='cuda'))
    w = torch.randn(co, ci, 3, 3, device='cuda'); s = torch.rand(N, ci, device='cuda') + 0.5; d = torch.rand(N, co, device='cuda') + 0.5
    wp = C.prep_weights(w, taps)
    for _ in range(2):
        C.igemm_conv(x, wp, offs, a_scale=s, o_scale=d, bias=torch.zeros(co, device='cuda'), act='lrelu', gain=1.4)
        C.igemm_wgrad(g, x, [(0, 0)] * 9, offs, (r, r), g_scale=d, x_scale=s)
    torch.cuda.synchronize()
    del x, g
