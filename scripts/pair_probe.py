"""Parity + timing probe of the conv kernel variants at the BASELINE configs[1] layer shapes, for one setting of the tuning environment
(SGV_CONV_PAIR, SGV_CONV_CLUSTER, ... are read once per process): forward with the full epilogue vs an fp64 contraction of the same
TF32-rounded operands, the stride-2 data gradient, and the median launch time.  One JSON line per case.
    SGV_CONV_PAIR=1 timeout 300 python scripts/pair_probe.py"""
import json, os, sys
import numpy as np
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stylegan_v_b200 import conv as C


def tf32_round(t):
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def timeit(fn, iters=12, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def main():
    cl = lambda t: t.contiguous(memory_format=torch.channels_last)
    taps, offs = C.conv3x3_taps()
    env = {k: os.environ.get(k) for k in ('SGV_CONV_PAIR', 'SGV_CONV_CLUSTER')}
    for name, N, ci, co, r in (('b32.conv1', 32, 512, 512, 32), ('b64.conv1', 32, 256, 256, 64), ('b128.conv1', 32, 128, 128, 128), ('b256.conv1', 32, 64, 64, 256)):
        g = torch.Generator().manual_seed(r)
        x = torch.randn(N, ci, r, r, generator=g).cuda(); w = torch.randn(co, ci, 3, 3, generator=g).cuda()
        s = (torch.randn(N, ci, generator=g) + 1).cuda(); d = (torch.rand(N, co, generator=g) + 0.5).cuda() / np.sqrt(ci * 9); b = torch.randn(co, generator=g).cuda()
        wp = C.prep_weights(w, taps, x3=False)
        kw = dict(a_scale=s, o_scale=d, bias=b, act='lrelu', gain=float(np.sqrt(2)))
        var = C.igemm_conv(cl(x), wp, offs, query=True, **kw)
        y = C.igemm_conv(cl(x), wp, offs, **kw)
        torch.cuda.synchronize()
        n_chk = min(N, 4)                          # fp64 reference on a few samples (the kernel computed all of them)
        xs = tf32_round(x[:n_chk] * s[:n_chk, :, None, None])
        ref = F.conv2d(xs.double(), tf32_round(w).double(), padding=1) * d[:n_chk].double()[:, :, None, None] + b.double()[None, :, None, None]
        ref = F.leaky_relu(ref, 0.2) * np.sqrt(2)
        err = rel(y[:n_chk], ref)
        err_last = rel(y[N - 1:], F.leaky_relu(F.conv2d(tf32_round(x[N - 1:] * s[N - 1:, :, None, None]).double(), tf32_round(w).double(), padding=1)
                                               * d[N - 1:].double()[:, :, None, None] + b.double()[None, :, None, None], 0.2) * np.sqrt(2))
        xc = cl(x)
        ms = timeit(lambda: C.igemm_conv(xc, wp, offs, **kw))
        y3 = C.igemm_conv(xc, C.prep_weights(w, taps, x3=True), offs, **kw)
        ref32 = F.leaky_relu(F.conv2d((x[:n_chk] * s[:n_chk, :, None, None]).double(), w.double(), padding=1) * d[:n_chk].double()[:, :, None, None]
                             + b.double()[None, :, None, None], 0.2) * np.sqrt(2)
        print(json.dumps(dict(case=name, env=env, variant=var, err_vs_same_operands=err, err_last_sample=err_last, x3_err=rel(y3[:n_chk], ref32), ms=ms,
                              tflops=2.0 * N * r * r * ci * co * 9 / ms / 1e9)), flush=True)
        del x, y, y3, xc
    # stride-2 data gradient (b256.conv0 backward) with the styles epilogue + fused reduction
    g = torch.Generator().manual_seed(2)
    N, Cg, Cx, h = 8, 64, 128, 128
    du = tf32_round(torch.randn(N, Cg, 2 * h + 1, 2 * h + 1, generator=g).cuda()); w = torch.randn(Cg, Cx, 3, 3, generator=g).cuda()
    s = (torch.randn(N, Cx, generator=g) + 1).cuda(); x = torch.randn(N, Cx, h, h, generator=g).cuda()
    wp = C.prep_weights(w, C.TAPS_3x3, rows_dim=1, cols_dim=0, x3=False)
    ds = torch.zeros(N, Cx, device='cuda')
    kw = dict(out_hw=(h, h), in_stride=2, o_scale=s, a_ready=True)
    var = C.igemm_conv(cl(du), wp, C.TAPS_3x3, query=True, **kw)
    dx = C.igemm_conv(cl(du), wp, C.TAPS_3x3, red_x=cl(x), red_out=ds, **kw)
    raw = F.conv2d(du.double(), tf32_round(w).double().transpose(0, 1), stride=2)
    duc = cl(du)
    ms = timeit(lambda: C.igemm_conv(duc, wp, C.TAPS_3x3, **kw))
    print(json.dumps(dict(case='dgrad stride 2 (b256.conv0)', env=env, variant=var, err=rel(dx, raw * s.double()[:, :, None, None]),
                          err_dstyles=rel(ds, (raw * x.double()).sum(dim=[2, 3])), ms=ms)), flush=True)


if __name__ == '__main__':
    main()
