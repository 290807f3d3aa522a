"""Runs each HBM-bound kernel a few times at the b256 shape (target for `ncu --set full -k regex:...`)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stylegan_v_b200.ops import upfirdn2d as U, bias_act as B
f = U.setup_filter([1, 3, 3, 1], device='cuda')
x = torch.randn(32, 64, 257, 257, device='cuda')
xcl = x.contiguous(memory_format=torch.channels_last)
z = torch.randn(32, 64, 256, 256, device='cuda')
b = torch.randn(64, device='cuda')
for _ in range(3):
    U.upfirdn2d(x, f, padding=1, gain=4)
    U.upfirdn2d(xcl, f, padding=1, gain=4)
    B.bias_act(z, b, act='lrelu')
torch.cuda.synchronize()
