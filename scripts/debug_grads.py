import os, sys, json, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from conftest import load_golden, rel_err
from oracle import synthesis_ref as sr
from stylegan_v_b200.synthesis import SynthesisNetwork
g, meta = load_golden('synthesis_tiny.npz')
_t = lambda a: torch.from_numpy(np.asarray(a))
cfg = sr.SynthesisConfig(**meta)
net = SynthesisNetwork.from_config(cfg)
net.load_state_dict({k[2:]: _t(g[k]) for k in g.files if k.startswith('p:')})
net = net.cuda()
ws = _t(g['ws']).cuda().requires_grad_(True)
t = _t(g['t']).cuda(); mz = _t(g['motion_z']).cuda()
img = net(ws, t, motion_z=mz, t_max=float(t.max()))
print('img', rel_err(img, _t(g['img_train'])))
names = sorted(k[2:] for k in g.files if k.startswith('g:'))
params = dict(net.named_parameters())
grads = torch.autograd.grad(img, [ws] + [params[n] for n in names], _t(g['dimg']).cuda())
print('d_ws', rel_err(grads[0], _t(g['d_ws'])))
for n, gr in zip(names, grads[1:]):
    e = rel_err(gr, _t(g['g:' + n]))
    if e > 2e-3:
        print(f'{e:.3e}', n, tuple(gr.shape))
