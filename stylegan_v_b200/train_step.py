"""The training phases around the hot path (BASELINE configs 3 and 4): loss terms, gradient exchange and parameter update.

Follows the reference's StyleGAN2Loss.accumulate_gradients (src/training/loss.py:73-173) and the phase loop of
training_loop.py:236-262,340-400, restricted to what StyleGAN-V's own config uses (c_dim = 0, no ADA pipe, style mixing off):

    Gmain   softplus(-D(G(z, t)))                           gradients into G                 loss.py:84-99
    Greg    path-length penalty (second order through G)    every G_reg_interval steps        loss.py:101-119
    Dmain   softplus(D(G(z, t).detach())) + softplus(-D(real))                                loss.py:121-147
    Dreg    R1 penalty gamma/2 * |d logits / d real|^2      every D_reg_interval steps        loss.py:149-160

What is B200-specific is the plumbing: parameters, gradients, Adam moments and the EMA copy of each network live in flat fp32
buffers (optim.FlatModuleState); a phase ends with ONE NCCL all-reduce of the flat gradient buffer and ONE fused
nan_to_num + Adam (+ EMA) launch (csrc/optim_step.cu) that also re-zeroes the gradients.  Gmain / Dmain run the fused NHWC
synthesis layers; Greg needs gradients of gradients and runs the synthesis network's `unfused` formulation on the drop-in ops.
"""
import copy

import numpy as np
import torch
import torch.nn.functional as F

from .ops import conv2d_gradfix
from .optim import FlatModuleState, FusedAdamEMA


# fused=True routes the discriminator's conv layers through the first-order-only fused nodes (stylegan_v_b200/dconv.py); the R1 term below
# differentiates D twice and therefore never does.
FUSED_DISCRIMINATOR = True


def _d_kwargs(x):
    return dict(fused=True) if (FUSED_DISCRIMINATOR and x.is_cuda) else dict(fused=False)


def run_discriminator(D, img, c, t, augment_pipe=None, video_consistent_aug=True, **d_kwargs):
    """D on (optionally augmented) frames — loss.py:58-72.  With video_consistent_aug the F frames of a clip enter the pipe as one
    [B, F*3, H, W] image so that they receive the same geometric and colour transform."""
    if augment_pipe is not None:
        if video_consistent_aug:
            nf, ch, h, w = img.shape
            f = t.shape[1]
            img = augment_pipe(img.view(nf // f, f * ch, h, w)).view(nf, ch, h, w)
        else:
            img = augment_pipe(img)
    return D(img, c, t, **d_kwargs)


def generator_main_loss(G, D, z, c, t, augment_pipe=None, video_consistent_aug=True, **synthesis_kwargs):
    img = G(z, c, t, **synthesis_kwargs)
    return F.softplus(-run_discriminator(D, img, c, t, augment_pipe, video_consistent_aug, **_d_kwargs(img))['image_logits']).mean()


def discriminator_main_loss(G, D, real_img, real_c, real_t, z, c, t, augment_pipe=None, video_consistent_aug=True, **synthesis_kwargs):
    """Returns (loss on generated frames, loss on real frames); the reference backpropagates them separately (loss.py:139,173)."""
    with torch.no_grad():
        fake = G(z, c, t, **synthesis_kwargs)
    loss_gen = F.softplus(run_discriminator(D, fake, c, t, augment_pipe, video_consistent_aug, **_d_kwargs(fake))['image_logits']).mean()
    loss_real = F.softplus(-run_discriminator(D, real_img, real_c, real_t, augment_pipe, video_consistent_aug, **_d_kwargs(fake))['image_logits']).mean()
    return loss_gen, loss_real


def discriminator_r1_loss(D, real_img, real_c, real_t, r1_gamma, augment_pipe=None, video_consistent_aug=True):
    img = real_img.detach().requires_grad_(True)
    logits = run_discriminator(D, img, real_c, real_t, augment_pipe, video_consistent_aug, fused=False)['image_logits']
    with conv2d_gradfix.no_weight_gradients():
        grads, = torch.autograd.grad([logits.sum()], [img], create_graph=True, only_inputs=True)
    penalty = grads.square().sum([1, 2, 3]) * (r1_gamma / 2)                      # per frame
    per_clip = penalty.view(-1, len(img) // len(logits)).mean(dim=1)              # loss.py:158
    return (logits * 0 + per_clip).mean()


def generator_path_length_loss(G, z, c, t, pl_mean, pl_weight=2.0, pl_decay=0.01, pl_batch_shrink=2, **synthesis_kwargs):
    """Path-length regularisation (loss.py:101-119).  pl_mean: 0-dim tensor, updated in place.  The reference's last line adds a [B]
    penalty to a [B * F] dummy, which only broadcasts for F = 1 (its stylegan-v config disables the term); here the per-latent penalty
    is averaged over latents, which is the same value for every F."""
    n = max(z.shape[0] // pl_batch_shrink, 1)
    ws = G.mapping(z[:n], c[:n])
    img = G.synthesis(ws, t=t[:n], c=c[:n], unfused=True, **synthesis_kwargs)
    noise = torch.randn_like(img) / np.sqrt(img.shape[2] * img.shape[3])
    with conv2d_gradfix.no_weight_gradients():
        grads, = torch.autograd.grad([(img * noise).sum()], [ws], create_graph=True, only_inputs=True)
    lengths = grads.square().sum(2).mean(1).sqrt()
    mean = pl_mean.lerp(lengths.mean(), pl_decay)
    pl_mean.copy_(mean.detach())
    return (img[:, 0, 0, 0].sum() * 0 + (lengths - mean).square() * pl_weight).mean()


class TrainingPhases:
    """G and D with flat state, lazily-regularised Adam (training_loop.py:243-250) and the EMA generator.

    step(...) runs the phases due at this iteration in the reference's order (Gmain, [Greg], Dmain, [Dreg]) and returns their loss
    values (tensors; no host sync).

    Lazy regularisation follows the reference exactly: whenever a reg interval is configured the optimiser of that network runs with
    lr * r and betas ** r, r = interval / (interval + 1) (training_loop.py:243-250) — independently of the loss WEIGHTS.  StyleGAN-V's own
    config has pl_weight = 0 with G_reg_interval = 4: G then trains at 0.8 * lr with beta2 = 0.99 ** 0.8 and the Greg phase produces no
    gradients (loss.py:101: do_Gpl needs pl_weight != 0), so torch's Adam skips every parameter; here the phase is skipped outright."""

    def __init__(self, G, D, lr=0.0025, betas=(0.0, 0.99), eps=1e-8, r1_gamma=0.2048, pl_weight=0.0, G_reg_interval=4, D_reg_interval=16,
                 ema_kimg=20.0, ema_rampup=None, batch_size=64, num_frames_per_video=None, process_group=None, device_step=False,
                 augment_pipe=None, video_consistent_aug=True):
        assert next(G.parameters()).is_cuda, 'TrainingPhases drives the CUDA path only'
        conv2d_gradfix.enabled = True                                               # training_loop.py:143
        self.G, self.D = G, D
        self.aug = dict(augment_pipe=augment_pipe, video_consistent_aug=video_consistent_aug)      # ADA pipe in front of D (loss.py:58-70); None = off
        if getattr(G.synthesis, '_pstream', None) is not None:
            G.synthesis._pstream = None                                             # a CUDA stream handle is not deep-copyable; it is re-created lazily
        self.G_ema = copy.deepcopy(G).eval().requires_grad_(False)
        self.r1_gamma, self.pl_weight = r1_gamma, pl_weight
        self.G_reg_interval, self.D_reg_interval = G_reg_interval, D_reg_interval       # optimiser scaling: keyed on the interval alone
        self.run_greg = G_reg_interval is not None and pl_weight != 0                   # whether the reg phase has anything to do
        self.run_dreg = D_reg_interval is not None and r1_gamma != 0
        self.ema_kimg, self.ema_rampup, self.batch_size = ema_kimg, ema_rampup, batch_size
        self.num_frames_per_video = num_frames_per_video
        self.pl_mean = torch.zeros([], device=next(G.parameters()).device)
        self.G_state = FlatModuleState(list(G.parameters()), list(self.G_ema.parameters()), process_group)
        self.D_state = FlatModuleState(list(D.parameters()), None, process_group)
        # replicas start from rank 0's parameters and buffers (training_loop.py:215-232 "Distribute across GPUs"); only gradients are
        # exchanged afterwards
        self.G_state.broadcast(0)
        self.D_state.broadcast(0)
        if self.G_state.world_size() > 1:
            import torch.distributed as dist
            for module in (G, self.G_ema, D):
                for buf in module.buffers():
                    dist.broadcast(buf, src=0 if process_group is None else dist.get_global_rank(process_group, 0), group=process_group)

        def make_opt(state, interval):
            if interval is None:
                return FusedAdamEMA(state, lr=lr, betas=betas, eps=eps, device_step=device_step)
            r = interval / (interval + 1)                                            # lazy regularisation: training_loop.py:245-248
            return FusedAdamEMA(state, lr=lr * r, betas=[b ** r for b in betas], eps=eps, device_step=device_step)
        # main and regularisation phases of a network share one optimiser, like the reference's (name + 'main', name + 'reg') pairs
        self.G_opt = make_opt(self.G_state, self.G_reg_interval)
        self.D_opt = make_opt(self.D_state, self.D_reg_interval)
        self.cur_nimg = 0
        self.it = 0

    def ema_beta(self):
        """training_loop.py:393-396, evaluated with cur_nimg BEFORE this iteration's increment (first iteration with a ramp-up: beta = 0,
        i.e. G_ema = G)."""
        nimg = self.ema_kimg * 1000
        if self.ema_rampup is not None:
            nimg = min(nimg, self.cur_nimg * self.ema_rampup)
        return 0.5 ** (self.batch_size / max(nimg, 1e-8))

    def _finish(self, state, opt, ema_beta=None):
        state.all_reduce()                                                           # SUM over ranks; 1/world is applied by the update kernel
        opt.step(ema_beta=ema_beta, zero_grad=True)
        if ema_beta is not None:
            # buffers are copied, not averaged (training_loop.py:399-400); w_avg is the only G buffer that changes during training
            self.G_ema.mapping.w_avg.copy_(self.G.mapping.w_avg)

    def _grad_mode(self, train_G):
        self.G.requires_grad_(train_G)
        self.D.requires_grad_(not train_G)

    PHASES = ('Gmain', 'Greg', 'Dmain', 'Dreg')

    @staticmethod
    def _per_phase(v, name):
        """Per-phase generator inputs: a dict {phase: tensor} (missing phases fall back to 'Gmain') or one tensor shared by all phases."""
        if isinstance(v, dict):
            return v.get(name, v['Gmain'])
        return v

    # The two main phases up to (not including) the gradient exchange: loss + backward into the flat gradient buffer.  Separate entry points so that
    # a caller can replay each as a CUDA graph and keep the NCCL all-reduce + update (`finish_g` / `finish_d`) between the replays (bench.py, N > 1).
    def backward_gmain(self, z, c, t, **synthesis_kwargs):
        self._grad_mode(True)
        loss = generator_main_loss(self.G, self.D, z, c, t, **self.aug, **synthesis_kwargs)
        loss.backward()
        return loss.detach()

    def backward_dmain(self, real_img, real_c, real_t, z, c, t, **synthesis_kwargs):
        self._grad_mode(False)
        loss_gen, loss_real = discriminator_main_loss(self.G, self.D, real_img, real_c, real_t, z, c, t, **self.aug, **synthesis_kwargs)
        (loss_gen + loss_real).backward()
        return (loss_gen + loss_real).detach()

    def finish_g(self, ema=True):
        self._finish(self.G_state, self.G_opt, ema_beta=self.ema_beta() if ema else None)

    def finish_d(self):
        self._finish(self.D_state, self.D_opt)

    def step(self, real_img, real_t, z, t, c=None, real_c=None, **synthesis_kwargs):
        """One iteration: real_img [B*F, 3, R, R], real_t [B, F]; generator latents z [B, z_dim] and times t [B, F].

        The reference draws INDEPENDENT gen_z / gen_c / gen_t for every phase (training_loop.py:333-348: all_gen_z split per phase):
        pass z / t / c as dicts keyed by phase name ('Gmain', 'Greg', 'Dmain', 'Dreg') to reproduce that — `sample_phase_latents` builds
        them; a plain tensor is shared by all phases (benchmarks, CUDA-graph capture with static inputs).
        EMA is applied with the last G update of the iteration (the reference applies it after all phases, training_loop.py:392-400 —
        same values, since only G phases change G)."""
        z0 = self._per_phase(z, 'Gmain')
        B = z0.shape[0]
        zeros_c = torch.zeros(B, 0, device=z0.device)
        c = zeros_c if c is None else c
        real_c = zeros_c if real_c is None else real_c
        pick = lambda name: (self._per_phase(z, name), self._per_phase(c, name), self._per_phase(t, name))
        out = {}
        do_greg = self.run_greg and self.it % self.G_reg_interval == 0
        do_dreg = self.run_dreg and self.it % self.D_reg_interval == 0
        ema_beta = self.ema_beta()                                                   # from the pre-increment image count
        # ---- G phases
        pz, pc, pt = pick('Gmain')
        out['Gmain'] = self.backward_gmain(pz, pc, pt, **synthesis_kwargs)
        self._finish(self.G_state, self.G_opt, ema_beta=None if do_greg else ema_beta)
        if do_greg:
            pz, pc, pt = pick('Greg')
            loss = generator_path_length_loss(self.G, pz, pc, pt, self.pl_mean, self.pl_weight, **synthesis_kwargs)
            loss.mul(self.G_reg_interval).backward()
            out['Greg'] = loss.detach()
            self._finish(self.G_state, self.G_opt, ema_beta=ema_beta)
        # ---- D phases
        pz, pc, pt = pick('Dmain')
        out['Dmain'] = self.backward_dmain(real_img, real_c, real_t, pz, pc, pt, **synthesis_kwargs)
        self._finish(self.D_state, self.D_opt)
        if do_dreg:
            loss = discriminator_r1_loss(self.D, real_img, real_c, real_t, self.r1_gamma, **self.aug)
            loss.mul(self.D_reg_interval).backward()
            out['Dreg'] = loss.detach()
            self._finish(self.D_state, self.D_opt)
        frames = self.num_frames_per_video if self.num_frames_per_video is not None else int(real_t.shape[1])
        self.cur_nimg += self.batch_size * frames                                    # training_loop.py:403: batch_size * num_frames_per_video
        self.it += 1
        return out

    def sample_phase_latents(self, batch, z_dim, t_sampler, device, generator=None):
        """Independent generator inputs per phase, like the reference's data fetch (training_loop.py:333-348): returns (z, t) dicts keyed by
        phase name.  t_sampler(batch) -> [batch, F] frame positions (the reference's sample_frames)."""
        z = {name: torch.randn(batch, z_dim, device=device, generator=generator) for name in self.PHASES}
        t = {name: t_sampler(batch).to(device) for name in self.PHASES}
        return z, t
