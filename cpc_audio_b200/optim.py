"""Flat gradient bucket + fused Adam.

``GradBucket`` makes every ``p.grad`` a view into ONE fp32 buffer and registers those views as *gradient sinks*:
the backward passes of the B200 autograd functions then accumulate parameter gradients straight into the
bucket (no per-parameter accumulate kernels) and the data-parallel exchange is a single all-reduce (sum,
matching ``totLoss = allLosses.sum()`` over replicas, cpc/train.py:85).

``FlatAdam`` is a ``torch.optim.Optimizer`` (cpc/train.py:335-337 builds ``torch.optim.Adam``; :339-343 loads its
``state_dict``; :351-370 wraps it in ``StepLR`` / ``LambdaLR``): same ``param_groups`` keys, same ``state_dict()``
format - a checkpoint written by either optimizer loads into the other - but parameters, gradients and both moments
of a group live in flat buffers and ``step()`` is one kernel per group.  ``PeerAdam`` additionally performs the
data-parallel gradient exchange inside that kernel, over peer memory.
"""
from __future__ import annotations

import ctypes
import os
import struct
import weakref

import torch

from . import _lib as L

_SINKS = {}  # parameter data_ptr -> (weakref to the parameter, gradient view inside a GradBucket, weakref to the bucket)


def sinks_for(params):
    """Gradient views for `params` if every one of them belongs to a live bucket (and still uses it), else None.

    ``Optimizer.zero_grad()`` defaults to ``set_to_none=True``: a parameter of a live bucket whose ``.grad`` is None is
    re-attached to its (freshly cleared) view here, so that a stock ``torch.optim.Adam`` loop keeps accumulating into
    the bucket instead of silently leaving it stale."""
    out = []
    for p in params:
        ent = _SINKS.get(p.data_ptr())
        if ent is None:
            return None
        owner, v, bucket = ent[0](), ent[1], ent[2]()
        if owner is None or bucket is None or owner.data_ptr() != p.data_ptr() or v.shape != p.shape:
            return None
        if owner.grad is None:
            bucket.reattach()
        if owner.grad is not v:
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
        me = weakref.ref(self)
        for p, v in zip(self.params, self.views):
            p.grad = v
            _SINKS[p.data_ptr()] = (weakref.ref(p), v, me)

    def detach(self):
        for p in self.params:
            _SINKS.pop(p.data_ptr(), None)

    def zero(self):
        self.flat.zero_()
        for p, v in zip(self.params, self.views):
            if p.grad is not v:
                p.grad = v

    def reattach(self):
        """Make ``p.grad`` the bucket view again for every parameter that lost it: views of parameters whose gradient was
        set to None (``zero_grad(set_to_none=True)``) are cleared, a stray gradient tensor (autograd allocated one while
        the view was detached) is copied into its view.  One memset when every gradient is None (the usual case)."""
        none = [p.grad is None for p in self.params]
        if all(none):
            self.flat.zero_()
        for p, v, n in zip(self.params, self.views, none):
            if p.grad is v:
                continue
            if n:
                if not all(none):
                    v.zero_()
            else:
                v.copy_(p.grad)
            p.grad = v

    def allreduce(self):
        """Sum the bucket over the data-parallel ranks (cpc/train.py:85 semantics).  If ``arm_overlap`` was called before
        the backward pass, everything but the late (conv0 / batchNorm0) gradients is reduced on a side stream as soon as
        the encoder backward signals that it is final - concurrently with the last data-gradient GEMM and the conv0
        backward - and only the late range (a few KB) is reduced after the backward pass."""
        import torch.distributed as dist
        if any(p.grad is not v for p, v in zip(self.params, self.views)):
            self.reattach()  # never reduce a stale buffer while the optimizer steps on detached gradients
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
        """Call right before ``backward()``: the next encoder backward ON THIS STREAM records the 'early gradients are
        final' event."""
        ov = self._overlap
        if ov is None:
            return
        dev = self.flat.device
        with torch.cuda.device(dev):
            L.check(L.lib().cpcb200_encoder_bwd_set_event(L.stream_ptr(dev), ctypes.c_void_p(ov["ready"].cuda_event)),
                    "encoder_bwd_set_event")
        ov["armed"] = True


_STATE_WORDS = 16  # int32 words of the device-side optimizer state (include/cpc_b200.h: cpcb200_adam_step_dev)
_ST_STEPS, _ST_ERR, _ST_LR = 0, 7, 8


def _f32_bits(x):
    return struct.unpack("<i", struct.pack("<f", float(x)))[0]


class FlatAdam(torch.optim.Optimizer):
    """torch.optim.Adam(lr, betas, eps, weight_decay) over flat parameter buffers, one kernel per param group and step.

    ``capturable=True`` keeps the step count AND the learning rate on the device (like ``torch.optim.Adam(capturable=
    True)``), so that ``step()`` can be captured in a CUDA graph and an ``lr_scheduler`` still takes effect on replays
    (``sync_lr()`` uploads ``group['lr']`` when it changed); ``fuse_zero_grad=True`` additionally clears the gradient
    bucket inside the same kernel, which makes the following ``zero_grad()`` free."""

    def __init__(self, params, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, capturable=False,
                 fuse_zero_grad=False, amsgrad=False, maximize=False):
        if amsgrad or maximize:
            raise NotImplementedError("cpc_audio_b200.FlatAdam: amsgrad / maximize are not implemented")
        capturable = bool(capturable or fuse_zero_grad)
        # same keys as torch.optim.Adam's groups, so that state_dict()s interchange (cpc/train.py:339-343)
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False, foreach=None,
                        capturable=capturable, differentiable=False, fused=None, decoupled_weight_decay=False)
        super().__init__(params, defaults)
        self.capturable = capturable
        self.fuse_zero_grad = bool(fuse_zero_grad)
        self.params = [p for g in self.param_groups for p in g["params"]]
        dev = self.params[0].device
        if any(p.device != dev or p.dtype != torch.float32 for p in self.params):
            raise ValueError("FlatAdam: all parameters must be fp32 tensors on one device")
        sizes = [p.numel() for p in self.params]
        n = sum(sizes)
        self.flat_p = torch.empty(n, device=dev, dtype=torch.float32)
        for v, p in zip(self.flat_p.split(sizes), self.params):
            v.view_as(p).copy_(p.data)
            p.data = v.view_as(p)
        self.bucket = GradBucket(self.params, flat=self._alloc_grad_bucket(n, dev))
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        # flat ranges of the groups (16-byte aligned starts are required by the vectorised kernels: every group but the
        # first starts where the previous one ends, so group sizes must be multiples of 4 elements when there are several)
        self._ranges, off = [], 0
        for g in self.param_groups:
            m = sum(p.numel() for p in g["params"])
            self._ranges.append((off, off + m))
            off += m
        if len(self._ranges) > 1 and any(lo % 4 for lo, _ in self._ranges):
            raise ValueError("FlatAdam: with several param groups every group must hold a multiple of 4 elements")
        self._steps = [0] * len(self.param_groups)
        self._states = [torch.zeros(_STATE_WORDS, device=dev, dtype=torch.int32) for _ in self.param_groups] if capturable else None
        self._dev_lr = [None] * len(self.param_groups)
        self._grads_clean = False
        if capturable:
            self.sync_lr()

    def _alloc_grad_bucket(self, n, dev):
        return None  # GradBucket allocates an ordinary device buffer

    # ---- learning rate on the device (graph replays see scheduler updates) ----
    def sync_lr(self):
        """Upload ``group['lr']`` to the device-side state when it changed (capturable mode).  Call it outside a graph
        capture; ``GraphedTrainStep`` does so before every replay."""
        if not self.capturable:
            return
        for gi, g in enumerate(self.param_groups):
            lr = float(g["lr"])
            if self._dev_lr[gi] != lr:
                self._states[gi][_ST_LR:_ST_LR + 1].copy_(torch.tensor([_f32_bits(lr)], dtype=torch.int32), non_blocking=False)
                self._dev_lr[gi] = lr

    @property
    def steps(self):
        return int(self._states[0][_ST_STEPS].item()) if self.capturable else self._steps[0]

    def _group_steps(self, gi):
        return int(self._states[gi][_ST_STEPS].item()) if self.capturable else self._steps[gi]

    def _launch(self, gi, g, lo, hi):
        dev = self.flat_p.device
        b1, b2 = g["betas"]
        args = (L.ptr(self.flat_p[lo:hi]), L.ptr(self.bucket.flat[lo:hi]), L.ptr(self.exp_avg[lo:hi]),
                L.ptr(self.exp_avg_sq[lo:hi]), hi - lo)
        if self.capturable:
            L.check(L.lib().cpcb200_adam_step_dev(*args, -1.0, b1, b2, g["eps"], g["weight_decay"], L.ptr(self._states[gi]),
                                                  1 if self.fuse_zero_grad else 0, L.stream_ptr(dev)), "adam_step_dev")
        else:
            self._steps[gi] += 1
            L.check(L.lib().cpcb200_adam_step(*args, float(g["lr"]), b1, b2, g["eps"], g["weight_decay"], self._steps[gi],
                                              L.stream_ptr(dev)), "adam_step")

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if any(p.grad is not v for p, v in zip(self.bucket.params, self.bucket.views)):
            self.bucket.reattach()
        if self.capturable and not torch.cuda.is_current_stream_capturing():
            self.sync_lr()
        with torch.cuda.device(self.flat_p.device):
            for gi, (g, (lo, hi)) in enumerate(zip(self.param_groups, self._ranges)):
                self._launch(gi, g, lo, hi)
        self._grads_clean = self.capturable and self.fuse_zero_grad
        return loss

    def zero_grad(self, set_to_none=False):
        """Clears the bucket in place (``set_to_none`` is ignored: the gradients ARE the all-reduce payload)."""
        if self._grads_clean:  # the fused step already cleared the bucket; only re-attach the views if needed
            self._grads_clean = False
            for p, v in zip(self.bucket.params, self.bucket.views):
                if p.grad is not v:
                    p.grad = v
            return
        self.bucket.zero()

    # ---- checkpoint interchange with torch.optim.Adam (cpc/train.py:220 saves, :339-343 loads) ----
    def _export_state(self):
        sizes = [p.numel() for p in self.params]
        gi_of = [gi for gi, g in enumerate(self.param_groups) for _ in g["params"]]
        steps = [self._group_steps(gi) for gi in range(len(self.param_groups))]
        for p, m, v, gi in zip(self.params, self.exp_avg.split(sizes), self.exp_avg_sq.split(sizes), gi_of):
            if steps[gi] == 0:
                self.state.pop(p, None)  # torch.optim.Adam has no state before its first step either
                continue
            step_t = torch.tensor(float(steps[gi]), dtype=torch.float32, device=p.device if self.capturable else "cpu")
            self.state[p] = {"step": step_t, "exp_avg": m.view_as(p), "exp_avg_sq": v.view_as(p)}

    def state_dict(self):
        self._export_state()
        return super().state_dict()

    def load_state_dict(self, state_dict):
        own = [{k: g[k] for k in ("capturable", "foreach", "fused", "differentiable")} for g in self.param_groups]
        super().load_state_dict(state_dict)
        for g, keep in zip(self.param_groups, own):
            if g.get("amsgrad") or g.get("maximize"):
                raise NotImplementedError("cpc_audio_b200.FlatAdam: cannot load an amsgrad / maximize Adam state")
            g.update(keep)
        sizes = [p.numel() for p in self.params]
        gi_of = [gi for gi, g in enumerate(self.param_groups) for _ in g["params"]]
        steps = [None] * len(self.param_groups)
        for p, m, v, gi in zip(self.params, self.exp_avg.split(sizes), self.exp_avg_sq.split(sizes), gi_of):
            st = self.state.get(p)
            if not st:
                m.zero_(); v.zero_()
                s = 0
            else:
                m.view_as(p).copy_(st["exp_avg"])
                v.view_as(p).copy_(st["exp_avg_sq"])
                s = int(float(st["step"]))
            if steps[gi] is None:
                steps[gi] = s
            elif steps[gi] != s:
                raise ValueError("FlatAdam.load_state_dict: parameters of one group carry different step counts")
        for gi, s in enumerate(steps):
            self._set_steps(gi, s or 0)
        self._dev_lr = [None] * len(self.param_groups)
        self.sync_lr()
        self._export_state()  # state entries are views of the flat buffers again

    def _set_steps(self, gi, s):
        if not self.capturable:
            self._steps[gi] = s
            return
        b1, b2 = self.param_groups[gi]["betas"]
        st = self._states[gi]
        host = torch.zeros(_STATE_WORDS, dtype=torch.int32)
        host[:] = st.cpu()
        host[_ST_STEPS] = s
        # running products beta^(steps+1) as float64 (include/cpc_b200.h): words 2..5
        pw = torch.tensor([float(b1) ** (s + 1), float(b2) ** (s + 1)], dtype=torch.float64).view(torch.int32)
        host[2:6] = pw
        st.copy_(host)


_TORCH_ADAM = torch.optim.Adam


def Adam(params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, **kw):
    """``torch.optim.Adam``'s call signature over ``FlatAdam`` (``cpc_audio_b200.patch.install(adam=True)`` puts it in
    torch.optim's place, so that the unchanged ``torch.optim.Adam(g_params, lr=..., betas=..., eps=...)`` of cpc/train.py:
    335-337 builds the flat optimizer: ONE kernel per step, zero_grad folded in, gradients written by the backward kernels
    straight into its bucket - same state_dict format, same param_groups).  Anything FlatAdam does not cover (CPU or
    non-fp32 parameters, amsgrad, maximize) gets the stock optimizer."""
    params = list(params)
    flat_ok = not amsgrad and not kw.get("maximize", False) and not kw.get("differentiable", False)
    tensors = [p for g in params for p in g["params"]] if params and isinstance(params[0], dict) else params
    flat_ok = flat_ok and len(tensors) > 0 and all(isinstance(p, torch.Tensor) and p.is_cuda and p.dtype == torch.float32
                                                   and p.device == tensors[0].device for p in tensors)
    if params and isinstance(params[0], dict) and len(params) > 1:
        flat_ok = flat_ok and all(sum(p.numel() for p in g["params"]) % 4 == 0 for g in params)
    if not flat_ok:
        return _TORCH_ADAM(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad, **kw)
    return FlatAdam(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, capturable=True, fuse_zero_grad=True)


class PeerAdam(FlatAdam):
    """FlatAdam whose ``step()`` also performs the data-parallel gradient exchange: the all-reduce(sum) of the bucket over
    the GPUs of the node runs through peer memory (NVLink / NVSwitch loads and stores, node-wide barriers on signal words)
    inside the optimizer kernel (``cpcb200_allreduce_adam_step``) - instead of an NCCL all-reduce followed by an optimizer
    kernel.  With ``overlap=True`` every gradient except the late ones (conv0 / batchNorm0, a few KB) is reduced by a small
    side-stream kernel (``cpcb200_peer_reduce_range``) that starts as soon as the encoder backward has finalised them and
    runs concurrently with the last data-gradient GEMM and the conv0 backward; the step kernel then only exchanges the late
    range.  The gradient bucket and the signal words live in ``torch.distributed._symmetric_memory`` (that module is only
    used to allocate and peer-map them).

    Use it exactly like FlatAdam, WITHOUT calling ``bucket.allreduce()``; every rank must step the same number of times.
    A rank that waits longer than ``timeout_s`` (default 600 s, env ``CPC_B200_PEER_TIMEOUT_S``) for its peers leaves the
    kernel WITHOUT applying the update and raises at the next ``check()`` / ``state_dict()`` - it does not fault the
    context.  Phases that can skew by minutes (rank-0 checkpointing, validation) should end with ``host_barrier()``."""

    def __init__(self, params, group=None, overlap=False, late_params=None, timeout_s=None, **kw):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self._symm, self._dist = symm, dist
        self._group = group if group is not None else dist.group.WORLD
        kw["capturable"] = True
        try:  # older torch releases want the group registered first; newer ones do it on rendezvous (and warn)
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", FutureWarning)
                symm.enable_symm_mem_for_group(self._group.group_name)
        except (AttributeError, RuntimeError, TypeError):
            pass
        super().__init__(params, **kw)
        if len(self.param_groups) != 1:
            raise ValueError("PeerAdam: one param group (the whole bucket is exchanged as one payload)")
        dev = self.flat_p.device
        self._hdl = symm.rendezvous(self.bucket.flat, self._group)
        self._sig = symm.empty(128, dtype=torch.int32, device=dev)
        self._sig.zero_()
        self._sig_hdl = symm.rendezvous(self._sig, self._group)
        world, rank = self._hdl.world_size, self._hdl.rank
        if world > 8:
            raise RuntimeError("PeerAdam: at most 8 GPUs (one NVSwitch node)")
        # the kernels use grads[rank] / signals[rank] as the LOCAL buffers: they must be the tensors this object holds
        if int(self._hdl.buffer_ptrs[rank]) != self.bucket.flat.data_ptr():
            raise RuntimeError("PeerAdam: symmetric-memory handle does not map the gradient bucket at its own address "
                               f"({int(self._hdl.buffer_ptrs[rank]):#x} != {self.bucket.flat.data_ptr():#x})")
        if int(self._sig_hdl.buffer_ptrs[rank]) != self._sig.data_ptr():
            raise RuntimeError("PeerAdam: symmetric-memory handle does not map the signal words at their own address")
        self._peers = L.Peers()
        for i in range(world):
            self._peers.grads[i] = int(self._hdl.buffer_ptrs[i])
            self._peers.signals[i] = int(self._sig_hdl.buffer_ptrs[i])
        self._peers.rank, self._peers.world = rank, world
        # NVSwitch multicast mapping of the bucket, when the fabric offers one: in-switch reduction (multimem.ld_reduce)
        # measured on B200 (tools/peer_adam_check.py): with 2 GPUs plain peer loads/stores reduce a slice faster
        # (23 us vs 39 us); the in-switch reduction pays off once a slice has several remote copies to sum
        mode = os.environ.get("CPC_B200_MULTIMEM", "auto")
        mc = 0
        if mode == "1" or (mode == "auto" and world >= 4):
            mc = int(getattr(self._hdl, "multicast_ptr", 0) or 0)  # 0 when the allocation has no multicast mapping
            if not mc and rank == 0:
                print("cpc_audio_b200.PeerAdam: no NVSwitch multicast mapping for the gradient bucket; using peer loads/stores",
                      flush=True)
        self._peers.grads_mc = mc if mc else None
        self.multicast = bool(mc)
        if timeout_s is None:
            timeout_s = float(os.environ.get("CPC_B200_PEER_TIMEOUT_S", "600"))
        self._peers.timeout_ns = int(timeout_s * 1e9)
        # overlap: early / late ranges of the bucket
        self.overlap = bool(overlap)
        self._early = self._late = None
        if self.overlap:
            late_ids = {id(p) for p in (late_params or [])}
            if not late_ids:
                raise ValueError("PeerAdam(overlap=True) needs late_params (the encoder's conv0 / batchNorm0 parameters)")
            early, late = split_ranges([p.numel() for p in self.params], [id(p) in late_ids for p in self.params])
            # ranges are handed to the kernels in units of 4 floats: widen the late ranges to 16-byte boundaries and cut
            # them out of the early ones
            n = self.flat_p.numel()
            late = [(lo // 4 * 4, min(n, (hi + 3) // 4 * 4)) for lo, hi in late]
            self._late = _merge(late)
            self._early = _subtract([(0, n)], self._late)
            self._side = torch.cuda.Stream(device=dev)
            self._ev_ready, self._ev_done = torch.cuda.Event(), torch.cuda.Event()
            self._ev_ready.record(torch.cuda.current_stream(dev))
            self._ev_done.record(torch.cuda.current_stream(dev))
            self._armed = False
            if len(self._late) > 4 or len(self._early) > 4:
                raise ValueError("PeerAdam: at most 4 early and 4 late ranges")
        torch.cuda.synchronize(dev)
        dist.barrier(group=self._group)  # every rank's buffers are zeroed and mapped before the first step

    def _alloc_grad_bucket(self, n, dev):
        flat = self._symm.empty(n, dtype=torch.float32, device=dev)
        flat.zero_()
        return flat

    def host_barrier(self):
        """Host-side rendezvous of the ranks (NCCL barrier): call it after a phase that only some ranks execute."""
        torch.cuda.synchronize(self.flat_p.device)
        self._dist.barrier(group=self._group)

    def check(self):
        """Raise if a step gave up waiting for its peers (synchronises the device)."""
        if int(self._states[0][_ST_ERR].item()) != 0:
            raise RuntimeError("cpc_audio_b200.PeerAdam: a rank did not reach the gradient exchange within the timeout; the "
                               "optimizer step was NOT applied on this rank (replicas have diverged) - restart from a checkpoint")

    def state_dict(self):
        self.check()
        return super().state_dict()

    def arm_overlap(self):
        """Call right before ``backward()`` (``GraphedTrainStep(before_backward=opt.arm_overlap)``): the next encoder
        backward on this stream records the 'early gradients are final' event."""
        if not self.overlap:
            return
        dev = self.flat_p.device
        with torch.cuda.device(dev):
            L.check(L.lib().cpcb200_encoder_bwd_set_event(L.stream_ptr(dev), ctypes.c_void_p(self._ev_ready.cuda_event)),
                    "encoder_bwd_set_event")
        self._armed = True

    def _ranges_arg(self, ranges):
        arr = (ctypes.c_int64 * 8)()
        for i, (lo, hi) in enumerate(ranges):
            arr[2 * i], arr[2 * i + 1] = lo, hi
        return arr, len(ranges)

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise NotImplementedError("PeerAdam.step(closure)")
        dev = self.flat_p.device
        g = self.param_groups[0]
        b1, b2 = g["betas"]
        if not torch.cuda.is_current_stream_capturing():
            self.sync_lr()
        lib = L.lib()
        with torch.cuda.device(dev):
            late, n_late = None, 0
            if self.overlap and self._armed:
                self._armed = False
                cur = torch.cuda.current_stream(dev)
                self._side.wait_event(self._ev_ready)
                arr, cnt = self._ranges_arg(self._early)
                with torch.cuda.stream(self._side):
                    L.check(lib.cpcb200_peer_reduce_range(ctypes.byref(self._peers), arr, cnt, L.ptr(self._states[0]),
                                                          ctypes.c_void_p(self._side.cuda_stream)), "peer_reduce_range")
                    self._ev_done.record(self._side)
                cur.wait_event(self._ev_done)
                late, n_late = self._ranges_arg(self._late)
            L.check(lib.cpcb200_allreduce_adam_step(ctypes.byref(self._peers), L.ptr(self.flat_p), L.ptr(self.exp_avg),
                                                    L.ptr(self.exp_avg_sq), self.flat_p.numel(), -1.0, b1, b2, g["eps"],
                                                    g["weight_decay"], L.ptr(self._states[0]), 1 if self.fuse_zero_grad else 0,
                                                    late, n_late, L.stream_ptr(dev)),
                    "allreduce_adam_step")
        self._grads_clean = self.fuse_zero_grad


def _merge(ranges):
    out = []
    for lo, hi in sorted(ranges):
        if out and lo <= out[-1][1]:
            out[-1] = (out[-1][0], max(out[-1][1], hi))
        else:
            out.append((lo, hi))
    return out


def _subtract(ranges, holes):
    out = []
    for lo, hi in ranges:
        cur = lo
        for hlo, hhi in holes:
            if hhi <= cur or hlo >= hi:
                continue
            if hlo > cur:
                out.append((cur, hlo))
            cur = max(cur, hhi)
        if cur < hi:
            out.append((cur, hi))
    return out
