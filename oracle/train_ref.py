"""ORACLE (test infrastructure only — never imported by the product path).

CPU restatement of the pieces either side of the contraction stack:

* `time_encoder_tail_ref` — the elementwise part of MotionMappingNetwork / AlignedTimeEncoder
  (/root/reference/src/training/motion.py:111-115 for t_left / t_right / interp_weights, :198-212 for the embedding),
  taking the stacked predictor outputs the product kernel takes.  Pinned through `synthesis_ref.motion_encoder`, which is
  checked against tests/golden/synthesis_tiny.npz (`motion_v` minted from the unmodified reference).
* `optimizer_step_ref` — the parameter update of one training phase (training_loop.py:381-386: nan_to_num = clamp(nansum),
  torch_utils/misc.py:49-56; then torch.optim.Adam.step(), the very optimiser class the reference constructs,
  train.py:192-193) and the G_ema update (training_loop.py:392-400).  torch.optim.Adam is third-party arithmetic for the
  reference too (torch; environment.yaml:8-10 pins pytorch 1.7.1, environment-ampere.yaml:15-17 pytorch 1.9); it is CALLED here
  (torch 2.11 in this image), not restated.  Its published algorithm for the options the reference passes (no weight decay, no
  amsgrad): m = b1*m + (1-b1)*g; v = b2*v + (1-b2)*g*g; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps) — unchanged between
  those versions.
"""
import torch


def time_encoder_tail_ref(heads_left, aligners_right, t, freqs, phase_scales, d):
    """heads_left [M,4F] = [P u_L | Phi u_L | A u_L], aligners_right [M,2F], t [M] -> [M,2F]."""
    nf = freqs.numel()
    freqs, phase_scales = freqs.reshape(1, nf), phase_scales.reshape(1, nf)
    t = t.reshape(-1)
    t_left = t - t % d                                                    # motion.py:111
    t_right = t_left + d                                                  # :112
    a = ((t % d) / d).reshape(-1, 1).to(torch.float32)                    # :114
    periods = heads_left[:, :nf].tanh() + 1                               # :198
    phases = heads_left[:, nf:2 * nf]                                     # :199
    al_left, al_right = heads_left[:, 2 * nf:], aligners_right            # :200-201

    def emb(tt):
        raw = freqs * periods * tt.float().unsqueeze(1) + phases * phase_scales      # :203-205
        return torch.cat([raw.sin(), raw.cos()], dim=1)                               # :207-209
    remove = emb(t_left) * (1 - a) + emb(t_right) * a                     # :212
    add = al_left * (1 - a) + al_right * a                                # :213
    return emb(t) - remove + add                                          # :214


def time_encoder_tail_c(heads_left, aligners_right, t, freqs, phase_scales, d):
    """The same tail through the plain-C port (oracle/c/sgv_oracle.c::oracle_time_encoder_tail_f32): scalar loops, libm sin/cos/tanh."""
    import ctypes
    from . import ops_ref
    L = ops_ref.lib()
    hl, ar = heads_left.contiguous().float(), aligners_right.contiguous().float()
    tt, fr, ps = t.reshape(-1).contiguous().float(), freqs.reshape(-1).contiguous().float(), phase_scales.reshape(-1).contiguous().float()
    m, nf = tt.numel(), fr.numel()
    out = torch.empty(m, 2 * nf)
    vp = lambda x: ctypes.c_void_p(x.data_ptr())
    L.oracle_time_encoder_tail_f32(vp(hl), vp(ar), vp(tt), vp(fr), vp(ps), vp(out), ctypes.c_int(m), ctypes.c_int(nf), ctypes.c_float(float(d)))
    return out


def nan_to_num_ref(g, nan=0.0, posinf=1e5, neginf=-1e5):
    """torch_utils/misc.py:49-56 — note that it clamps finite values as well."""
    assert nan == 0
    return torch.clamp(g.unsqueeze(0).nansum(0), min=neginf, max=posinf)


class OptimizerRef:
    """One phase's optimiser + EMA, on plain tensors (CPU)."""

    def __init__(self, params, ema_params=None, lr=0.002, betas=(0.0, 0.99), eps=1e-8):
        self.params = [p.detach().clone().requires_grad_(True) for p in params]
        self.ema = [p.detach().clone() for p in ema_params] if ema_params is not None else None
        self.opt = torch.optim.Adam(self.params, lr=lr, betas=betas, eps=eps)          # train.py:192-193, training_loop.py:241-250

    def step(self, grads, ema_beta=None, grad_scale=1.0):
        for p, g in zip(self.params, grads):
            p.grad = nan_to_num_ref(g.detach().clone() * grad_scale)                   # training_loop.py:383-385
        self.opt.step()                                                                # :386
        if ema_beta is not None:
            with torch.no_grad():
                for pe, p in zip(self.ema, self.params):
                    pe.copy_(p.lerp(pe, ema_beta))                                     # :397-398
