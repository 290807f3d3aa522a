"""CPU study (no GPU needed): what does rounding the contraction OPERANDS of every conv in the 256^2 synthesis network cost in image
accuracy?  The oracle network (fp32 accumulation, F.conv2d) is run with x and w rounded to TF32 (what the tcgen05 kernels do today), to fp16
(`kind::f16`: same 10-bit mantissa, 2 bytes per element => half the shared-memory bytes per FLOP, but a 5-bit exponent), to fp16 with a
per-sample power-of-two block scale, and to bf16.  Output: normwise relative error of the image (max |a-b| / max |b|) against fp32."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from oracle import synthesis_ref as sr

torch.set_num_threads(min(os.cpu_count() or 1, 16))
RES = int(sys.argv[1]) if len(sys.argv) > 1 else 256
oc, oct_ = F.conv2d, F.conv_transpose2d
STATS = {}


def rn_tf32(t):
    return ((t.contiguous().view(torch.int32) + 0x1000) & ~0x1fff).view(torch.float32)


def rn_fp16(t):
    STATS['max'] = max(STATS.get('max', 0.0), float(t.abs().max()))
    nz = t[t != 0].abs()
    if nz.numel():
        STATS['min'] = min(STATS.get('min', 1e30), float(nz.min()))
        STATS['sub'] = STATS.get('sub', 0) + int((nz < 6.1e-5).sum())      # below the smallest normal fp16
        STATS['n'] = STATS.get('n', 0) + nz.numel()
    return t.half().float()


def rn_fp16_scaled(t):
    # power-of-two scale per sample (activations) / per tensor (weights) so that the largest magnitude sits at 2^14
    dims = list(range(1, t.ndim))
    m = t.abs().amax(dim=dims, keepdim=True).clamp_min(1e-30)
    s = torch.exp2(14 - torch.ceil(torch.log2(m)))
    return (t * s).half().float() / s


def rn_bf16(t):
    return t.bfloat16().float()


def patched(rx, rw):
    def c2(input, weight, bias=None, **kw):
        return oc(rx(input), rw(weight), bias, **kw)

    def ct2(input, weight, bias=None, **kw):
        return oct_(rx(input), rw(weight), bias, **kw)
    return c2, ct2


cfg = sr.SynthesisConfig(img_resolution=RES)
P = sr.init_params(cfg, seed=0)
g = torch.Generator().manual_seed(1)
N = 2
ws = torch.randn(N, cfg.num_ws, cfg.w_dim, generator=g)
t = torch.zeros(N, 1)
mz = torch.randn(N, sr.max_traj_len(cfg, 0.0), cfg.motion_z_dim, generator=g)


def run():
    with torch.no_grad():
        return sr.synthesis_forward(P, cfg, ws, t, motion_z=mz, fused_modconv=False)


ref = run()
print(f'{RES}x{RES} synthesis network, random-init weights, {N} frames; image max |.| = {float(ref.abs().max()):.3f}')
for name, rx, rw in (('tf32 operands (today)', rn_tf32, rn_tf32), ('fp16 operands', rn_fp16, rn_fp16), ('fp16 operands, pow2 block scale', rn_fp16_scaled, rn_fp16_scaled),
                     ('bf16 operands', rn_bf16, rn_bf16), ('fp16 activations, tf32 weights', rn_fp16, rn_tf32)):
    STATS.clear()
    F.conv2d, F.conv_transpose2d = patched(rx, rw)
    torch.nn.functional.conv2d, torch.nn.functional.conv_transpose2d = F.conv2d, F.conv_transpose2d
    img = run()
    F.conv2d, F.conv_transpose2d = oc, oct_
    err = float((img - ref).abs().max() / ref.abs().max())
    extra = ''
    if STATS:
        extra = f"   operand |.| range [{STATS.get('min', 0):.2e}, {STATS.get('max', 0):.2e}], {100.0 * STATS.get('sub', 0) / max(STATS.get('n', 1), 1):.3f} % of non-zeros below the fp16 normal range"
    print(f'  {name:36s} rel err {err:.2e}{extra}')
