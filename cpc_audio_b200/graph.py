"""Whole-step CUDA graph: the body of cpc/train.py:83-91 captured once and replayed.

A CPC step is ~45 kernel launches of 3-300 us each; replaying them as one graph removes the per-launch CPU work from
the step (it matters when the host must wait for the loss of step i before it may enqueue step i+1, as
cpc/train.py:98 does) and lets the driver pre-resolve the programmatic dependencies between consecutive kernels.

    step = GraphedTrainStep(model, criterion, optimizer, example_batch, label)
    losses, acc = step(batch)         # batch: device tensor of example_batch's shape; outputs are static tensors

Requirements: a capturable optimizer (``FlatAdam(capturable=True)`` or ``torch.optim.Adam(capturable=True)``), fixed
shapes, and no host-side control flow that depends on device values inside the step.  ``torch.randint`` (negative
sampling, criterion.py:181-189) is graph-safe: the default CUDA generator registers its Philox offset with the graph,
so every replay draws fresh indices exactly as the eager loop would.
"""
from __future__ import annotations

import torch


class GraphedTrainStep:
    def __init__(self, model, criterion, optimizer, example_batch, label, allreduce=None, warmup=3, before_backward=None):
        self.model, self.criterion, self.optimizer = model, criterion, optimizer
        self.allreduce = allreduce
        self.before_backward = before_backward  # e.g. GradBucket.arm_overlap
        self.static_x = example_batch.clone()
        self.label = label
        dev = example_batch.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):  # lazy initialisation (function attributes, allocator pools) outside the capture
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.losses, self.acc = self._body()
        self.replays = 0

    def _body(self):
        c, z, _ = self.model(self.static_x, self.label)
        losses, acc = self.criterion(c, z, self.label)
        if self.before_backward is not None:
            self.before_backward()
        losses.sum().backward()
        if self.allreduce is not None:
            self.allreduce()
        self.optimizer.step()
        self.optimizer.zero_grad()  # set_to_none=True (torch default) is fine: GradBucket re-attaches its views
        return losses.detach(), acc.detach()

    def __call__(self, batch=None):
        if batch is not None and batch.data_ptr() != self.static_x.data_ptr():
            self.static_x.copy_(batch, non_blocking=True)
        sync_lr = getattr(self.optimizer, "sync_lr", None)
        if sync_lr is not None:
            sync_lr()  # an lr_scheduler changed group['lr'] (cpc/train.py:351-370): one 4-byte upload, no re-capture
        self.graph.replay()
        self.replays += 1
        return self.losses, self.acc
