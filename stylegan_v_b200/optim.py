"""Parameter update of the training step as ONE kernel launch over flat buffers.

Reference sequence per phase (src/training/training_loop.py:381-386,392-400):
    for p in module.parameters(): nan_to_num(p.grad, nan=0, posinf=1e5, neginf=-1e5, out=p.grad)     # clamp(nansum), misc.py:49-56
    opt.step()                                                                                        # torch.optim.Adam, betas (0, .99), eps 1e-8
    for p_ema, p in zip(G_ema.parameters(), G.parameters()): p_ema.copy_(p.lerp(p_ema, ema_beta))    # after the G phase

`FlatModuleState` re-homes every parameter of a module (and, optionally, of its EMA twin) as a view into one flat fp32
buffer, with gradients in a matching flat buffer — the buffer the data-parallel all-reduce runs on (the reference: torch DDP over NCCL,
training_loop.py:215-232; here SUM over ranks, the 1/world_size is folded into the update kernel).  `FusedAdamEMA.step()` is then a single
`sgv_adam_ema_step` launch (csrc/optim_step.cu).  There is no CPU implementation: CPU tensors raise.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib

_ALIGN = 64      # elements: every parameter view starts 256-byte aligned (TMA / 128-bit loads of the conv kernels need >= 16 B)


def _round_up(n, a):
    return (n + a - 1) // a * a


class FlatModuleState:
    """Flat parameter / gradient (/ EMA) storage of a module.

    params       parameters in `module.parameters()` order (all fp32, one device)
    ema_params   optional parameters of the EMA copy, same order and shapes
    After construction `p.data` and `p.grad` of every parameter are views into `self.param` / `self.grad`
    (values preserved), `p_ema.data` into `self.ema`."""

    def __init__(self, params, ema_params=None, process_group=None, early=None):
        """early: optional predicate on a parameter.  Parameters it selects are placed FIRST in the flat buffers (`early_numel` elements) and form
        a gradient bucket that `begin_backward()` / `finish_backward()` all-reduce as soon as the last of them has received its gradient,
        i.e. while the rest of the backward pass still runs (the reference's DDP buckets overlap the same way, training_loop.py:215-232)."""
        params = [p for p in params]
        self._early_params = [p for p in params if early is not None and early(p)]
        if self._early_params:
            assert ema_params is None, 'the early bucket re-orders the parameters; give EMA parameters in the same (re-ordered) order via state.params'
            ids = {id(p) for p in self._early_params}
            params = self._early_params + [p for p in params if id(p) not in ids]
        self.params = params
        assert self.params, 'no parameters'
        dev = self.params[0].device
        assert all(p.dtype == torch.float32 and p.device == dev for p in self.params), 'all parameters must be float32 on one device'
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += _round_up(p.numel(), _ALIGN)
        self.numel = off
        self.param = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(off, dtype=torch.float32, device=dev)
        self.ema = None
        self.group = process_group
        for p, o in zip(self.params, self.offsets):
            view = self.param[o:o + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            p.grad = self.grad[o:o + p.numel()].view_as(p)
        if ema_params is not None:
            self.ema_params = [p for p in ema_params]
            assert len(self.ema_params) == len(self.params) and all(a.shape == b.shape for a, b in zip(self.ema_params, self.params))
            self.ema = torch.zeros(off, dtype=torch.float32, device=dev)
            for p, o in zip(self.ema_params, self.offsets):
                view = self.ema[o:o + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view

        self.early_numel = self.offsets[len(self._early_params)] if 0 < len(self._early_params) < len(self.params) else (off if self._early_params else 0)
        self._bucket = dict(armed=False, left=0, work=None)
        if self._early_params:
            for p in self._early_params:
                p.register_post_accumulate_grad_hook(self._early_ready)

    # ---- overlapped gradient exchange: early bucket during the backward pass, the rest after it ----
    def _early_ready(self, _p):
        b = self._bucket
        if not b['armed']:
            return
        b['left'] -= 1
        if b['left'] == 0:
            b['work'] = dist.all_reduce(self.grad[:self.early_numel], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def begin_backward(self):
        """Arm the early bucket for ONE backward pass (no-op without an early bucket or with one rank).  Returns whether it is armed."""
        b = self._bucket
        b['armed'] = bool(self._early_params) and self.world_size() > 1
        b['left'], b['work'] = len(self._early_params), None
        return b['armed']

    def finish_backward(self):
        """After the backward pass: wait for the early bucket (a stream-level wait on CUDA: capturable in a CUDA graph together with the
        collectives) and reduce the rest of the buffer.  Without an armed bucket: the plain single all-reduce."""
        b = self._bucket
        if not b['armed']:
            return self.all_reduce()
        b['armed'] = False
        assert b['work'] is not None, 'the early gradient bucket never became ready (a selected parameter received no gradient)'
        b['work'].wait()
        if self.early_numel < self.numel:
            dist.all_reduce(self.grad[self.early_numel:], op=dist.ReduceOp.SUM, group=self.group)
        return None

    def zero_grad(self):
        self.grad.zero_()

    def world_size(self):
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.group)
        return 1

    def all_reduce(self, async_op=False):
        """SUM of the flat gradient buffer over ranks (one collective).  The average's 1/world_size is applied by the update
        kernel (`grad_scale`), not by an extra pass."""
        if self.world_size() == 1:
            return None
        return dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)

    def broadcast(self, src=0):
        """Replicas start identical: parameters (and the EMA copy) are broadcast from rank `src` before the first step, as the reference
        does for every parameter and buffer ("Distribute across GPUs", training_loop.py:215-232).  Two collectives over the flat buffers."""
        if self.world_size() == 1:
            return
        src_global = dist.get_global_rank(self.group, src) if self.group is not None else src
        dist.broadcast(self.param, src=src_global, group=self.group)
        if self.ema is not None:
            dist.broadcast(self.ema, src=src_global, group=self.group)

    def nbytes(self):
        return self.numel * 4


class FusedAdamEMA:
    """Adam (torch.optim.Adam arithmetic) + nan_to_num on the gradient + optional EMA lerp + optional gradient zeroing in one launch.

    device_step=True keeps the step counter on the device and advances it inside the launch sequence, so a captured CUDA
    graph containing `step()` is a valid optimiser step on every replay."""

    def __init__(self, state: FlatModuleState, lr=0.002, betas=(0.0, 0.99), eps=1e-8, grad_clamp=1e5, device_step=False):
        assert state.param.is_cuda, 'FusedAdamEMA is CUDA-only (there is no CPU path in libsgv_b200)'
        self.state = state
        self.lr, self.betas, self.eps, self.grad_clamp = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(grad_clamp)
        self.exp_avg = torch.zeros_like(state.param)
        self.exp_avg_sq = torch.zeros_like(state.param)
        self.t = 0
        self.step_count = torch.zeros(1, dtype=torch.int32, device=state.param.device) if device_step else None

    def step(self, ema_beta=None, zero_grad=False, grad_scale=None):
        """One update.  ema_beta=None skips the EMA (D phase).  grad_scale defaults to 1/world_size of the state's group
        (gradients were SUM-reduced by FlatModuleState.all_reduce)."""
        st = self.state
        if grad_scale is None:
            grad_scale = 1.0 / st.world_size()
        use_ema = ema_beta is not None
        assert not use_ema or st.ema is not None, 'no EMA parameters were given to FlatModuleState'
        self.t += 1
        q = _lib.AdamParams()
        q.param, q.grad, q.exp_avg, q.exp_avg_sq = st.param.data_ptr(), st.grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
        q.param_ema = st.ema.data_ptr() if use_ema else None
        q.numel = st.numel
        q.lr, q.beta1, q.beta2, q.eps = self.lr, self.betas[0], self.betas[1], self.eps
        q.ema_beta = float(ema_beta) if use_ema else 0.0
        q.grad_scale, q.grad_clamp = float(grad_scale), self.grad_clamp
        q.step = self.t
        q.step_count = self.step_count.data_ptr() if self.step_count is not None else None
        q.advance_step = 1 if self.step_count is not None else 0
        q.zero_grad = 1 if zero_grad else 0
        dev = st.param.device
        with torch.cuda.device(dev):
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(_lib.lib().sgv_adam_ema_step(ctypes.byref(q), stream), 'sgv_adam_ema_step')

    def algorithmic_bytes(self, ema, zero_grad):
        """HBM bytes one launch has to move (roofline numerator): loads p, g, m, v (+ p_ema), stores p, m, v (+ p_ema) (+ g)."""
        return self.state.numel * 4 * (4 + 3 + (2 if ema else 0) + (1 if zero_grad else 0))
