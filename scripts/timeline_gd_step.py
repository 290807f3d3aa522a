"""Kernel time breakdown of ONE G+D training step (BASELINE configs[2] shapes, eager launches) with torch.profiler / CUPTI:
total kernel time and the top kernels by summed duration.  Text summary on stdout."""
import os, sys, collections, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch.profiler import profile, ProfilerActivity
from stylegan_v_b200.networks import Generator, Discriminator
from stylegan_v_b200.train_step import TrainingPhases
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device('cuda')
B, Fr, RES = 16, 3, 256
G = Generator(img_resolution=RES).to(dev).train()
D = Discriminator(img_resolution=RES, mbstd_group_size=4).to(dev).train()
tp = TrainingPhases(G, D, r1_gamma=0.0, pl_weight=0.0, batch_size=B)
real = torch.randn(B * Fr, 3, RES, RES, device=dev).clamp_(-1, 1)
z = torch.randn(B, 512, device=dev)
t = (torch.randint(0, 900, (B, 1)).float() + torch.tensor([[0.0, 5.0, 9.0]])).to(dev)
for _ in range(3):
    tp.step(real, t, z, t)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tp.step(real, t, z, t)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
by = collections.defaultdict(float); cnt = collections.Counter()
for e in evs:
    short = e.name.split('(')[0][:90]
    by[short] += e.time_range.end - e.time_range.start; cnt[short] += 1
tot = sum(by.values())
print(f'{len(evs)} device activities, summed duration {tot / 1e3:.3f} ms')
for k, v in sorted(by.items(), key=lambda kv: -kv[1])[:45]:
    print(f'  {v / 1e3:8.3f} ms x{cnt[k]:<4d} {k}')
