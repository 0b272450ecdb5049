"""Flat gradient bucket + fused Adam.

``GradBucket`` makes every ``p.grad`` a view into ONE fp32 buffer and registers those views as *gradient sinks*:
the backward passes of the B200 autograd functions then accumulate parameter gradients straight into the
bucket (no per-parameter accumulate kernels) and the data-parallel exchange is a single NCCL all-reduce (sum,
matching ``totLoss = allLosses.sum()`` over replicas, cpc/train.py:85).  ``FlatAdam`` keeps the parameters and
both moments flat as well, so ``optimizer.step()`` is one kernel (torch.optim.Adam semantics, cpc/train.py:335-337).
"""
from __future__ import annotations

import ctypes
import os
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


def split_ranges(sizes, late):
    """Flat-buffer ranges [lo, hi) of the early and of the late tensors, adjacent tensors of the same kind merged."""
    out = ([], [])
    off = 0
    for n, is_late in zip(sizes, late):
        dst = out[1 if is_late else 0]
        if dst and dst[-1][1] == off:
            dst[-1] = (dst[-1][0], off + n)
        else:
            dst.append((off, off + n))
        off += n
    return out


class GradBucket:
    def __init__(self, params, flat=None):
        self.params = [p for p in params]
        dev = self.params[0].device
        sizes = [p.numel() for p in self.params]
        if flat is None:
            flat = torch.zeros(sum(sizes), device=dev, dtype=torch.float32)
        assert flat.numel() == sum(sizes) and flat.dtype == torch.float32 and flat.is_contiguous()
        self.flat = flat  # may live in symmetric (peer-mapped) memory: see PeerAdam
        self.views = [v.view_as(p) for v, p in zip(self.flat.split(sizes), self.params)]
        self._overlap = None
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
        """Sum the bucket over the data-parallel ranks (cpc/train.py:85 semantics).  If ``arm_overlap`` was called before
        the backward pass, everything but the late (conv0 / batchNorm0) gradients is reduced on a side stream as soon as
        the encoder backward signals that it is final - concurrently with the last data-gradient GEMM and the conv0
        backward - and only the late range (a few KB) is reduced after the backward pass."""
        import torch.distributed as dist
        ov = self._overlap
        if ov is None or not ov["armed"]:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            return
        ov["armed"] = False
        cur = torch.cuda.current_stream(self.flat.device)
        side = ov["side"]
        side.wait_event(ov["ready"])
        with torch.cuda.stream(side):
            for lo, hi in ov["early"]:
                dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM)
            ov["done"].record(side)
        for lo, hi in ov["late"]:
            dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM)
        cur.wait_event(ov["done"])

    def setup_overlap(self, late_params):
        """`late_params`: the parameters whose gradients only exist at the very end of the backward pass (the encoder's
        conv0.weight / conv0.bias / batchNorm0.weight / batchNorm0.bias)."""
        dev = self.flat.device
        late_ids = {id(p) for p in late_params}
        early, late = split_ranges([p.numel() for p in self.params], [id(p) in late_ids for p in self.params])
        ready, done = torch.cuda.Event(), torch.cuda.Event()
        ready.record(torch.cuda.current_stream(dev))  # materialise the cudaEvent_t handles
        done.record(torch.cuda.current_stream(dev))
        self._overlap = {"early": early, "late": late, "side": torch.cuda.Stream(device=dev), "ready": ready, "done": done,
                         "armed": False}

    def arm_overlap(self):
        """Call right before ``backward()``: the next encoder backward records the 'early gradients are final' event."""
        ov = self._overlap
        if ov is None:
            return
        L.check(L.lib().cpcb200_encoder_bwd_set_event(ctypes.c_void_p(ov["ready"].cuda_event)), "encoder_bwd_set_event")
        ov["armed"] = True


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
        self.bucket = GradBucket(self.params, flat=self._alloc_grad_bucket(sum(sizes), dev))
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.capturable = bool(capturable or fuse_zero_grad)
        self.fuse_zero_grad = bool(fuse_zero_grad)
        self._steps = 0
        self._state = torch.zeros(8, device=dev, dtype=torch.int32) if self.capturable else None  # [steps, ticket, b1^t, b2^t]
        self._grads_clean = False

    def _alloc_grad_bucket(self, n, dev):
        return None  # GradBucket allocates an ordinary device buffer

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


class PeerAdam(FlatAdam):
    """FlatAdam whose ``step()`` also performs the data-parallel gradient exchange: ONE kernel per step does the
    all-reduce(sum) of the bucket over the GPUs of the node through peer memory (NVLink / NVSwitch loads and stores, two
    node-wide barriers on signal words), the Adam update and ``zero_grad`` (``cpcb200_allreduce_adam_step``) - instead of
    an NCCL all-reduce followed by an optimizer kernel.  The gradient bucket and the signal words live in
    ``torch.distributed._symmetric_memory`` (that module is only used to allocate and peer-map them).

    Use it exactly like FlatAdam, WITHOUT calling ``bucket.allreduce()``; every rank must step the same number of times."""

    def __init__(self, params, group=None, **kw):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self._symm, self._dist = symm, dist
        self._group = group if group is not None else dist.group.WORLD
        kw["capturable"] = True
        try:  # older torch releases want the group registered first; newer ones do it on rendezvous
            symm.enable_symm_mem_for_group(self._group.group_name)
        except Exception:  # noqa: BLE001
            pass
        super().__init__(params, **kw)
        dev = self.flat_p.device
        self._hdl = symm.rendezvous(self.bucket.flat, self._group)
        self._sig = symm.empty(64, dtype=torch.int32, device=dev)
        self._sig.zero_()
        self._sig_hdl = symm.rendezvous(self._sig, self._group)
        world, rank = self._hdl.world_size, self._hdl.rank
        if world > 8:
            raise RuntimeError("PeerAdam: at most 8 GPUs (one NVSwitch node)")
        self._peers = L.Peers()
        for i in range(world):
            self._peers.grads[i] = int(self._hdl.buffer_ptrs[i])
            self._peers.signals[i] = int(self._sig_hdl.buffer_ptrs[i])
        self._peers.rank, self._peers.world = rank, world
        # NVSwitch multicast mapping of the bucket, when the fabric offers one: in-switch reduction (multimem.ld_reduce)
        mc = 0
        try:
            # measured on B200 (tools/peer_adam_check.py): with 2 GPUs plain peer loads/stores reduce a slice faster
            # (23 us vs 39 us); the in-switch reduction pays off once a slice has several remote copies to sum
            mode = os.environ.get("CPC_B200_MULTIMEM", "auto")
            if mode == "1" or (mode == "auto" and world >= 4):
                mc = int(self._hdl.multicast_ptr or 0)  # 0 when the allocation has no multicast mapping
        except Exception:  # noqa: BLE001
            mc = 0
        self._peers.grads_mc = mc if mc else None
        self.multicast = bool(mc)
        torch.cuda.synchronize(dev)
        dist.barrier(group=self._group)  # every rank's buffers are zeroed and mapped before the first step

    def _alloc_grad_bucket(self, n, dev):
        flat = self._symm.empty(n, dtype=torch.float32, device=dev)
        flat.zero_()
        return flat

    def step(self):
        dev = self.flat_p.device
        with torch.cuda.device(dev):
            L.check(L.lib().cpcb200_allreduce_adam_step(ctypes.byref(self._peers), L.ptr(self.flat_p), L.ptr(self.exp_avg),
                                                        L.ptr(self.exp_avg_sq), self.flat_p.numel(), self.lr, self.betas[0],
                                                        self.betas[1], self.eps, self.weight_decay, L.ptr(self._state),
                                                        1 if self.fuse_zero_grad else 0, L.stream_ptr(dev)),
                    "allreduce_adam_step")
        self._grads_clean = self.fuse_zero_grad
