"""Flat gradient bucket + fused Adam.

``GradBucket`` makes every ``p.grad`` a view into ONE fp32 buffer and registers those views as *gradient sinks*:
the backward passes of the B200 autograd functions then accumulate parameter gradients straight into the
bucket (no per-parameter accumulate kernels) and the data-parallel exchange is a single NCCL all-reduce (sum,
matching ``totLoss = allLosses.sum()`` over replicas, cpc/train.py:85).  ``FlatAdam`` keeps the parameters and
both moments flat as well, so ``optimizer.step()`` is one kernel (torch.optim.Adam semantics, cpc/train.py:335-337).
"""
from __future__ import annotations

import weakref

import torch

from . import _lib as L

_SINKS = {}  # parameter data_ptr -> (weakref to the parameter, gradient view inside a GradBucket)


def sinks_for(params):
    """Gradient views for `params` if every one of them belongs to a live bucket (and still uses it), else None."""
    out = []
    for p in params:
        ent = _SINKS.get(p.data_ptr())
        if ent is None:
            return None
        owner, v = ent[0](), ent[1]
        if owner is None or owner.data_ptr() != p.data_ptr() or owner.grad is not v or v.shape != p.shape:
            return None
        out.append(v)
    return out


class GradBucket:
    def __init__(self, params):
        self.params = [p for p in params]
        dev = self.params[0].device
        sizes = [p.numel() for p in self.params]
        self.flat = torch.zeros(sum(sizes), device=dev, dtype=torch.float32)
        self.views = [v.view_as(p) for v, p in zip(self.flat.split(sizes), self.params)]
        self.attach()

    def attach(self):
        for p, v in zip(self.params, self.views):
            p.grad = v
            _SINKS[p.data_ptr()] = (weakref.ref(p), v)

    def detach(self):
        for p in self.params:
            _SINKS.pop(p.data_ptr(), None)

    def zero(self):
        self.flat.zero_()
        for p, v in zip(self.params, self.views):
            if p.grad is not v:
                p.grad = v

    def allreduce(self):
        import torch.distributed as dist
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)


class FlatAdam:
    """torch.optim.Adam(lr, betas, eps, weight_decay) over one flat parameter buffer, one kernel per step.

    ``capturable=True`` keeps the step count on the device (like ``torch.optim.Adam(capturable=True)``) so that
    ``step()`` can be captured in a CUDA graph; ``fuse_zero_grad=True`` additionally clears the gradient bucket inside
    the same kernel, which makes the following ``zero_grad()`` free."""

    def __init__(self, params, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, capturable=False,
                 fuse_zero_grad=False):
        self.params = [p for p in params]
        dev = self.params[0].device
        sizes = [p.numel() for p in self.params]
        self.flat_p = torch.empty(sum(sizes), device=dev, dtype=torch.float32)
        for v, p in zip(self.flat_p.split(sizes), self.params):
            v.view_as(p).copy_(p.data)
            p.data = v.view_as(p)
        self.bucket = GradBucket(self.params)
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.capturable = bool(capturable or fuse_zero_grad)
        self.fuse_zero_grad = bool(fuse_zero_grad)
        self._steps = 0
        self._state = torch.zeros(8, device=dev, dtype=torch.int32) if self.capturable else None  # [steps, ticket, b1^t, b2^t]
        self._grads_clean = False

    @property
    def steps(self):
        return int(self._state[0].item()) if self.capturable else self._steps

    def step(self):
        dev = self.flat_p.device
        with torch.cuda.device(dev):
            if self.capturable:
                L.check(L.lib().cpcb200_adam_step_dev(L.ptr(self.flat_p), L.ptr(self.bucket.flat), L.ptr(self.exp_avg),
                                                      L.ptr(self.exp_avg_sq), self.flat_p.numel(), self.lr, self.betas[0],
                                                      self.betas[1], self.eps, self.weight_decay, L.ptr(self._state),
                                                      1 if self.fuse_zero_grad else 0, L.stream_ptr(dev)), "adam_step_dev")
                self._grads_clean = self.fuse_zero_grad
                return
            self._steps += 1
            L.check(L.lib().cpcb200_adam_step(L.ptr(self.flat_p), L.ptr(self.bucket.flat), L.ptr(self.exp_avg),
                                              L.ptr(self.exp_avg_sq), self.flat_p.numel(), self.lr, self.betas[0],
                                              self.betas[1], self.eps, self.weight_decay, self._steps, L.stream_ptr(dev)),
                    "adam_step")

    def zero_grad(self, set_to_none=False):
        if self._grads_clean:  # the fused step already cleared the bucket; only re-attach the views if needed
            self._grads_clean = False
            for p, v in zip(self.bucket.params, self.bucket.views):
                if p.grad is not v:
                    p.grad = v
            return
        self.bucket.zero()
