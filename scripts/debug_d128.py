"""Localises disagreements between the fused and the unfused route of the native Discriminator (128-channel test network of
tests/test_networks_gpu.py): per-tensor error of every gradient, fused vs unfused vs the fp32 CPU evaluation."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from conftest import rel_err
from test_networks_gpu import _d128, cos_sim

cuda = torch.device('cuda')
D = _d128(cuda)
g = torch.Generator().manual_seed(1)
img = torch.randn(6, 3, 32, 32, generator=g)
t = torch.tensor([[0.0, 5.0, 9.0], [100.0, 101.0, 131.0]])
c = torch.zeros(2, 0)
names = ['img'] + [n for n, _ in D.named_parameters()]


def run(dev, fused):
    Dd = D.to(dev).train()
    x = img.to(dev).requires_grad_(True)
    logits = Dd(x, c.to(dev), t.to(dev), fused=fused)['image_logits']
    grads = torch.autograd.grad(torch.nn.functional.softplus(-logits).mean(), [x] + list(Dd.parameters()), allow_unused=True)
    return logits.detach().cpu(), [None if a is None else a.detach().cpu() for a in grads]


l_cpu, g_cpu = run(torch.device('cpu'), False)
l_unf, g_unf = run(cuda, False)
l_fus, g_fus = run(cuda, True)
print('logits', rel_err(l_unf, l_cpu), rel_err(l_fus, l_cpu), rel_err(l_fus, l_unf))
for n, a, b, r in zip(names, g_fus, g_unf, g_cpu):
    if a is None:
        continue
    print(f'{n:28s} fused-vs-unfused {rel_err(a, b):.2e}  fused-vs-cpu {rel_err(a, r):.2e} cos {cos_sim(a, r):.5f}  unfused-vs-cpu {rel_err(b, r):.2e} cos {cos_sim(b, r):.5f}')
