"""Runs the persistent conv kernel once per high-resolution config-2 layer shape (target for `ncu --set full -k regex:conv_tf32_v3`)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stylegan_v_b200 import conv as C
N = 32
taps, offs = C.conv3x3_taps()
cl = lambda t: t.contiguous(memory_format=torch.channels_last)
for (ci, co, r) in ((64, 64, 256), (128, 128, 128), (256, 256, 64)):
    x = cl(torch.randn(N, ci, r, r, device='cuda'))
    w = torch.randn(co, ci, 3, 3, device='cuda'); s = torch.rand(N, ci, device='cuda') + 0.5; d = torch.rand(N, co, device='cuda') + 0.5
    wp = C.prep_weights(w, taps)
    b = torch.zeros(co, device='cuda')
    for _ in range(2):
        C.igemm_conv(x, wp, offs, a_scale=s, o_scale=d, bias=b, act='lrelu', gain=1.4)
    torch.cuda.synchronize()
    del x
