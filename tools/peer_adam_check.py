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
