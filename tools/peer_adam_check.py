"""torchrun --nproc-per-node N tools/peer_adam_check.py : PeerAdam (all-reduce + Adam + zero_grad in one kernel over peer
memory) against NCCL all-reduce + torch.optim.Adam on the same per-rank gradients; also times both.  Development tool."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpc_audio_b200.optim import FlatAdam, PeerAdam  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
shapes = [(12, 256, 256), (256, 1, 10), (256,), (256, 256, 8), (768, 256), (768,), (3,)]
gen = torch.Generator(device=dev).manual_seed(0)
p_ref = [torch.nn.Parameter(torch.randn(s, device=dev, generator=gen)) for s in shapes]
p_new = [torch.nn.Parameter(p.detach().clone()) for p in p_ref]
o_ref = torch.optim.Adam(p_ref, lr=2e-4)
o_new = PeerAdam(p_new, lr=2e-4, fuse_zero_grad=True)
ggen = torch.Generator(device=dev).manual_seed(100 + rank)
for it in range(5):
    for a, b in zip(p_ref, p_new):
        g = torch.randn(a.shape, device=dev, generator=ggen) * (10.0 ** (it - 3))
        tot = g.clone()
        dist.all_reduce(tot)
        a.grad = tot
        b.grad.copy_(g)
    o_ref.step()
    o_new.step()
    assert all(b.grad.eq(0).all() for b in p_new), "gradients not cleared"
    o_new.zero_grad()
err = max(((a - b).abs().max() / a.abs().max().clamp_min(1.0)).item() for a, b in zip(p_ref, p_new))
print(f"rank {rank}: multicast={o_new.multicast} steps {o_new.steps}, max rel param error vs NCCL + torch Adam: {err:.2e}", flush=True)
assert err < 5e-6, err

# the same with the early exchange on a side stream (everything but two 'late' tensors is all-reduced by
# cpcb200_peer_reduce_range, the step kernel only exchanges the late ranges); several steps so that the epochs advance
p_ref2 = [torch.nn.Parameter(p.detach().clone()) for p in p_ref]
p_ovl = [torch.nn.Parameter(p.detach().clone()) for p in p_ref]
o_ref2 = torch.optim.Adam(p_ref2, lr=2e-4)
o_ovl = PeerAdam(p_ovl, lr=2e-4, fuse_zero_grad=True, overlap=True, late_params=[p_ovl[1], p_ovl[2], p_ovl[6]])
for it in range(6):
    for a, b in zip(p_ref2, p_ovl):
        g = torch.randn(a.shape, device=dev, generator=ggen) * (10.0 ** (it - 3))
        tot = g.clone()
        dist.all_reduce(tot)
        a.grad = tot
        b.grad.copy_(g)
    o_ref2.step()
    o_ovl._ev_ready.record(torch.cuda.current_stream(dev))  # what the encoder backward does once the early gradients are final
    o_ovl._armed = True
    o_ovl.step()
    assert all(b.grad.eq(0).all() for b in p_ovl), "gradients not cleared"
    o_ovl.zero_grad()
o_ovl.check()
err2 = max(((a - b).abs().max() / a.abs().max().clamp_min(1.0)).item() for a, b in zip(p_ref2, p_ovl))
print(f"rank {rank}: overlap early={o_ovl._early} late={o_ovl._late}: max rel param error vs NCCL + torch Adam: {err2:.2e}", flush=True)
assert err2 < 5e-6, err2

if "--timeout-test" in sys.argv:
    # a rank that never arrives: its peers give up after timeout_s, flag the error, do NOT apply the update and keep
    # their CUDA context (no __trap)
    p_t = [torch.nn.Parameter(torch.randn(4096, device=dev))]
    o_t = PeerAdam(p_t, lr=2e-4, fuse_zero_grad=True, timeout_s=1.0)
    p_t[0].grad.fill_(1.0)
    o_t.step()
    o_t.check()
    before = p_t[0].detach().clone()
    if rank != world - 1:
        p_t[0].grad.fill_(1.0)
        o_t.step()
        try:
            o_t.check()
            raise SystemExit("timeout was not detected")
        except RuntimeError as e:
            assert "did not reach" in str(e)
        assert torch.equal(before, p_t[0].detach()), "the update must not be applied after a timeout"
        x = torch.ones(8, device=dev).sum().item()   # the context is alive
        print(f"rank {rank}: timeout flagged, parameters untouched, context alive ({x})", flush=True)
    torch.cuda.synchronize(); dist.barrier()

# timing at the size of the CPC bucket (2.5 M floats)
big = [torch.nn.Parameter(torch.randn(2498304, device=dev))]
o_big = PeerAdam(big, lr=2e-4, fuse_zero_grad=True)
big2 = [torch.nn.Parameter(torch.randn(2498304, device=dev))]
o_nccl = FlatAdam(big2, lr=2e-4, capturable=True, fuse_zero_grad=True)


def timeit(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def nccl_step():
    dist.all_reduce(o_nccl.bucket.flat)
    o_nccl.step()


t_new, t_old = timeit(o_big.step), timeit(nccl_step)
stamps = o_big._sig[40:45].tolist()
print(f"rank {rank}: last fused step, ns since kernel start: barrier1 {stamps[0]}, slice reduced {stamps[1]}, grid sync {stamps[2]}, "
      f"barrier2 {stamps[3]}, adam done {stamps[4]}", flush=True)
if rank == 0:
    print(f"world {world}: fused peer-memory all-reduce+Adam {t_new:.1f} us/step; NCCL all-reduce + Adam kernel {t_old:.1f} us/step", flush=True)
torch.cuda.synchronize(); dist.barrier()
sys.stdout.flush()
os._exit(0)
