"""Data-parallel gradient exchange for the batch-sharded hot path.

The path shards by clips/frames with replicated weights; the only exchange is one gradient all-reduce per step
(reference: torch DDP over NCCL, training_loop.py:215-232).  `FlatGradReducer` keeps all gradients in ONE flat
fp32 buffer (parameters' .grad are views into it, so backward kernels write straight into the communication
buffer — no gather copy) and averages it across ranks with a single NCCL all-reduce (gloo on CPU for tests).
"""
import torch
import torch.distributed as dist


class FlatGradReducer:
    def __init__(self, params, process_group=None):
        self.params = [p for p in params if p.requires_grad]
        assert self.params, 'no parameters to reduce'
        dev, dt = self.params[0].device, self.params[0].dtype
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=dt, device=dev)
        self.group = process_group
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    def zero(self):
        self.flat.zero_()

    def all_reduce(self):
        """Average gradients over ranks (no-op without an initialised process group / world size 1)."""
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        ws = dist.get_world_size(self.group)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.div_(ws)

    def nbytes(self):
        return self.flat.numel() * self.flat.element_size()
