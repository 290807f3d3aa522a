"""Localises CPU-vs-CUDA differences of the native Discriminator module by module (forward hooks), on the tiny golden config."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from conftest import load_golden, rel_err
from test_networks_cpu import make_discriminator, _t
from stylegan_v_b200.ops import conv2d_gradfix
from stylegan_v_b200 import native_conv

g, meta = load_golden('discriminator_tiny.npz')


def run(dev, tag):
    D = make_discriminator(g, meta).to(dev).train()
    outs = []
    hooks = []
    for name, m in D.named_modules():
        if name and not list(m.children()) or name.count('.') == 0 and name:
            hooks.append(m.register_forward_hook(lambda mod, i, o, name=name: outs.append((name, (o[0] if isinstance(o, tuple) else o))) if isinstance(o, (tuple, torch.Tensor)) else None))
    img = _t(g['img']).to(dev); t = _t(g['t']).to(dev)
    logits = D(img, torch.zeros(len(t), 0, device=dev), t)['image_logits']
    return [(n, o.detach().float().cpu()) for n, o in outs if isinstance(o, torch.Tensor)], logits.detach().cpu()


cpu, lc = run(torch.device('cpu'), 'cpu')
print('cpu logits', lc, 'golden', g['logits'])
for native in (True, False):
    conv2d_gradfix.enabled = True
    native_conv.enabled = native
    gpu, lg = run(torch.device('cuda'), 'cuda')
    print(f'--- native_conv.enabled={native}: cuda logits', lg)
    for (n, a), (n2, b) in zip(gpu, cpu):
        assert n == n2
        print(f'{n:28s} {tuple(b.shape)} rel_err {rel_err(a, b):.3e}')
