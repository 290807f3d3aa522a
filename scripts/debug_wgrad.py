import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stylegan_v_b200 import conv as C
torch.manual_seed(0)
N, Cin, Cout, H = 1, 32, 32, 8
cl = lambda t: t.contiguous(memory_format=torch.channels_last)
# g[n,o,y,x] = o+1 ; x[n,i,y,x] = 1  -> dw[t][o][i] = (o+1) * (#valid pixels for tap t)
g = (torch.arange(Cout, dtype=torch.float32) + 1).reshape(1, Cout, 1, 1).expand(N, Cout, H, H).cuda()
x = torch.ones(N, Cin, H, H).cuda()
dw = C.igemm_wgrad(cl(g), cl(x), [(0, 0)], [(0, 0)], (H, H))
print('1 tap, const g (o+1), x=1: dw[0][:4,:4]=\n', dw[0][:4, :4].cpu(), ' expected rows', [(o + 1) * H * H for o in range(4)])
x2 = (torch.arange(Cin, dtype=torch.float32) + 1).reshape(1, Cin, 1, 1).expand(N, Cin, H, H).cuda()
dw = C.igemm_wgrad(cl(torch.ones(N, Cout, H, H).cuda()), cl(x2), [(0, 0)], [(0, 0)], (H, H))
print('g=1, x=(i+1): dw[0][:4,:6]=\n', dw[0][:4, :6].cpu(), ' expected cols', [(i + 1) * H * H for i in range(6)])
# pixel-dependent pattern: g = delta at pixel (2,3) channel o -> picks x at that pixel
gp = torch.zeros(N, Cout, H, H); gp[0, :, 2, 3] = 1
xp = torch.arange(H * H, dtype=torch.float32).reshape(1, 1, H, H).expand(N, Cin, H, H).clone()
dw = C.igemm_wgrad(cl(gp.cuda()), cl(xp.cuda()), [(0, 0)], [(0, 0)], (H, H))
print('delta g at (2,3), x = pixel index: dw[0][0,:4]=', dw[0][0, :4].cpu(), ' expected', 2 * H + 3)
