"""One forward+backward step of the bench workload between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stylegan_v_b200.synthesis import SynthesisNetwork
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device('cuda')
N = int(os.environ.get('SGV_FRAMES', '32'))
net = SynthesisNetwork(img_resolution=256).to(dev).train()
ws = torch.randn(N, net.num_ws, net.w_dim, device=dev)
t = torch.zeros(N, 1, device=dev)
mz = torch.randn(N, net.motion_encoder.traj_len(), 512, device=dev)
dimg = torch.randn(N, 3, 256, 256, device=dev)


def step():
    for p in net.parameters():
        p.grad = None
    w = ws.clone().requires_grad_(True)
    img = net(w, t, motion_z=mz)
    (img * dimg).sum().backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
