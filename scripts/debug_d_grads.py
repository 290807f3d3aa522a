"""Per-parameter gradient agreement (rel err of max, cosine) of the native Discriminator on CUDA against the reference golden, for
(a) tcgen05 kernels, (b) library convs with TF32, (c) library convs in true fp32 — separates contraction rounding from everything else."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from conftest import load_golden, rel_err
from test_networks_cpu import make_discriminator, _t, cos_sim
from stylegan_v_b200.ops import conv2d_gradfix
from stylegan_v_b200 import native_conv

g, meta = load_golden('discriminator_tiny.npz')
dev = torch.device('cuda')
conv2d_gradfix.enabled = True
for tag, native, tf32 in (('tcgen05', True, True), ('cudnn-tf32', False, True), ('cudnn-fp32', False, False)):
    native_conv.enabled = native
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    D = make_discriminator(g, meta).to(dev).train()
    names = [k[2:] for k in g.files if k.startswith('g:')]
    img = _t(g['img']).to(dev).requires_grad_(True); t = _t(g['t']).to(dev)
    logits = D(img, torch.zeros(2, 0, device=dev), t)['image_logits']
    P = dict(D.named_parameters())
    grads = torch.autograd.grad(torch.nn.functional.softplus(-logits).mean(), [P[n] for n in names], retain_graph=True)
    print(f'=== {tag}: logits rel {rel_err(logits, _t(g["logits"])):.2e}')
    for n, a in zip(names, grads):
        print(f'{tag} G1 {n:36s} rel {rel_err(a, _t(g["g:" + n])):.2e} cos {cos_sim(a, _t(g["g:" + n])):.5f}')
    with conv2d_gradfix.no_weight_gradients():
        r1, = torch.autograd.grad(logits.sum(), [img], create_graph=True)
    print(f'{tag} r1_grads rel {rel_err(r1, _t(g["r1_grads"])):.2e} cos {cos_sim(r1, _t(g["r1_grads"])):.5f}')
    loss_r1 = (r1.square().sum([1, 2, 3]) * 0.5).view(-1, 3).mean(1).mean()
    names2 = [k[3:] for k in g.files if k.startswith('r1:')]
    g2 = torch.autograd.grad(loss_r1, [P[n] for n in names2], allow_unused=True)
    for n, a in zip(names2, g2):
        print(f'{tag} R1 {n:36s} rel {rel_err(a, _t(g["r1:" + n])):.2e} cos {cos_sim(a, _t(g["r1:" + n])):.5f}')
