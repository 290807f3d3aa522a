"""Micro-benchmark of the exact-fp32 dense kernels (csrc/dense_f32.cu) at the shapes of the 256^2 synthesis step (32 frames) against the
library calls they replace (cuBLAS fp32 addmm / matmul, cuDNN fp32 conv1d): forward and forward+backward, median of CUDA-event timings."""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stylegan_v_b200 import dense
from stylegan_v_b200.time_encoder import MotionMappingNetwork
from bench_conv import timeit


def fwd_bwd(make, params):
    def fn():
        y = make()
        torch.autograd.grad(y, params, torch.ones_like(y))
    return fn


def main():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = 'cuda'
    # FullyConnectedLayer shapes: mapping network layer, discriminator epilogue fc, time-encoder heads
    for name, M, K, O, act in [('mapping.fc', 32, 512, 512, 'lrelu'), ('D.epilogue.fc', 48, 8192, 512, 'lrelu'), ('time_enc.heads', 32, 512, 1024, 'linear')]:
        x = torch.randn(M, K, device=dev, requires_grad=True)
        w = torch.randn(O, K, device=dev, requires_grad=True)
        b = torch.randn(O, device=dev, requires_grad=True)
        g = 1 / np.sqrt(K)
        ours = lambda: dense.linear(x, w, b, g, 1.0, act=act, gain=1.0)
        lib = (lambda: F.leaky_relu(torch.addmm(b.unsqueeze(0), x, (w * g).t()), 0.2)) if act == 'lrelu' else (lambda: torch.addmm(b.unsqueeze(0), x, (w * g).t()))
        with torch.no_grad():
            f_ours, f_lib = timeit(ours), timeit(lib)
        print(json.dumps(dict(op=name, m=M, k=K, n=O, fwd_ms=f_ours, lib_fwd_ms=f_lib, fwd_bwd_ms=timeit(fwd_bwd(ours, [x, w, b])),
                              lib_fwd_bwd_ms=timeit(fwd_bwd(lib, [x, w, b])))), flush=True)
    # all style affines of the 256^2 synthesis network (26 layers, 14 w rows): one grouped launch vs 14 addmm calls
    M, K = 32, 512
    widths = [[512], [512, 512]] + [[512, 512]] * 3 * 2 + [[512, 512], [256, 256], [256, 256], [128, 128], [128, 128], [64, 64]]
    widths = widths[:14]
    ws = torch.randn(M, len(widths), K, device=dev, requires_grad=True)
    weights = [torch.randn(o, K, device=dev, requires_grad=True) for grp in widths for o in grp]
    biases = [torch.randn(o, device=dev, requires_grad=True) for grp in widths for o in grp]
    col = [0]
    for grp in widths:
        col.append(col[-1] + sum(grp))
    groups = dense.make_groups(col, list(range(len(widths))), K, ws.device)
    g = 1 / np.sqrt(K)
    ours = lambda: dense.stacked_affine(ws, torch.cat(weights), torch.cat(biases), groups, g)

    def lib():
        out, li = [], 0
        for gi, grp in enumerate(widths):
            wc = torch.cat(weights[li:li + len(grp)]) * g
            bc = torch.cat(biases[li:li + len(grp)])
            out.append(torch.addmm(bc.unsqueeze(0), ws[:, gi], wc.t()))
            li += len(grp)
        return torch.cat(out, dim=1)
    with torch.no_grad():
        f_ours, f_lib = timeit(ours), timeit(lib)
    pr = [ws] + weights + biases
    print(json.dumps(dict(op='style_affines(all layers)', m=M, k=K, n=col[-1], fwd_ms=f_ours, lib_fwd_ms=f_lib, fwd_bwd_ms=timeit(fwd_bwd(ours, pr)),
                          lib_fwd_bwd_ms=timeit(fwd_bwd(lib, pr)))), flush=True)
    # motion encoder: windows (1 and 3 frames per clip) and the full trajectory (16 frames) vs the cuDNN fp32 conv1d formulation
    for B, Fr in [(32, 1), (16, 3), (4, 16)]:
        enc = MotionMappingNetwork().to(dev)
        t = torch.randint(0, 1000, (B, Fr), device=dev).float()
        mz = torch.randn(B, enc.traj_len(), 512, device=dev)
        mz_lib = mz.clone().requires_grad_(True)            # a source that wants gradients takes the library conv1d route
        pr = list(enc.parameters())
        ours = lambda: enc(t, motion_z=mz)['motion_v']
        lib = lambda: enc(t, motion_z=mz_lib)['motion_v']
        with torch.no_grad():
            f_ours = timeit(ours)
        f_lib = timeit(lambda: lib().detach())
        print(json.dumps(dict(op='motion_encoder', clips=B, frames=Fr, fwd_ms=f_ours, lib_fwd_ms=f_lib, fwd_bwd_ms=timeit(fwd_bwd(ours, pr)),
                              lib_fwd_bwd_ms=timeit(fwd_bwd(lib, pr)))), flush=True)


if __name__ == '__main__':
    main()
