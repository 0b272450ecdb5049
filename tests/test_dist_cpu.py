"""Host-side logic of the data-parallel path on CPU with gloo, world_size 2:
one flat fp32 gradient bucket, one all-reduce(sum) per step, replicas stay bit-identical."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cpc_audio_b200.optim import GradBucket
    torch.manual_seed(0)  # replicas: same init on every rank
    params = [torch.nn.Parameter(torch.randn(7, 5)), torch.nn.Parameter(torch.randn(11)), torch.nn.Parameter(torch.randn(3, 2, 4))]
    bucket = GradBucket(params)
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(params, bucket.views))
    opt = torch.optim.Adam(params, lr=1e-2)
    for step in range(3):
        bucket.zero()
        g = torch.Generator().manual_seed(100 * step + rank)  # per-rank data
        loss = sum((p * torch.randn(p.shape, generator=g)).sum() for p in params)
        loss.backward()                      # autograd accumulates IN PLACE into the bucket views
        assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(params, bucket.views))
        local = bucket.flat.clone()
        bucket.allreduce()                   # one collective over the whole bucket, SUM (cpc/train.py:85 semantics)
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        assert torch.allclose(bucket.flat, sum(gathered))
        opt.step()
    flat_p = torch.cat([p.detach().reshape(-1) for p in params])
    gathered = [torch.empty_like(flat_p) for _ in range(world)]
    dist.all_gather(gathered, flat_p)
    assert all(torch.equal(gathered[0], t) for t in gathered), "replicas diverged"
    bucket.detach()
    out[rank] = True
    dist.destroy_process_group()


def _worker_default_zero_grad(rank, world, port, out):
    """a stock torch.optim.Adam loop - optimizer.zero_grad() with its default set_to_none=True - must keep the
    replicas in sync: allreduce() re-attaches the bucket (copying the gradients autograd allocated while it was detached)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cpc_audio_b200.optim import GradBucket
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(7, 5)), torch.nn.Parameter(torch.randn(11))]
    bucket = GradBucket(params)
    opt = torch.optim.Adam(params, lr=1e-2)
    for step in range(4):
        g = torch.Generator().manual_seed(100 * step + rank)
        loss = sum((p * torch.randn(p.shape, generator=g)).sum() for p in params)
        loss.backward()
        local = torch.cat([p.grad.reshape(-1) for p in params]).clone()
        bucket.allreduce()
        assert all(p.grad is v for p, v in zip(params, bucket.views)), "bucket not re-attached"
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        assert torch.allclose(bucket.flat, sum(gathered)), "a stale bucket was reduced"
        opt.step()
        opt.zero_grad()  # set_to_none=True
        assert all(p.grad is None for p in params)
    flat_p = torch.cat([p.detach().reshape(-1) for p in params])
    gathered = [torch.empty_like(flat_p) for _ in range(world)]
    dist.all_gather(gathered, flat_p)
    assert all(torch.equal(gathered[0], t) for t in gathered), "replicas diverged"
    bucket.detach()
    out[rank] = True
    dist.destroy_process_group()


def test_bucket_with_default_zero_grad_gloo_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_default_zero_grad, args=(world, port, out), nprocs=world, join=True)
    assert all(out.get(r) for r in range(world))


def test_flat_adam_state_dict_format_cpu():
    """FlatAdam is a torch.optim.Optimizer: param_groups with torch.optim.Adam's keys, state_dict() in torch.optim.Adam's
    format, load_state_dict() of a torch.optim.Adam checkpoint (cpc/train.py:339-343) - checked here without a GPU (no step)."""
    from cpc_audio_b200.optim import FlatAdam
    torch.manual_seed(0)
    shapes = [(8, 4), (12,), (2, 3, 4)]
    pt = [torch.nn.Parameter(torch.randn(s)) for s in shapes]
    pf = [torch.nn.Parameter(p.detach().clone()) for p in pt]
    ot = torch.optim.Adam(pt, lr=3e-4, betas=(0.8, 0.95), eps=1e-7)
    for _ in range(3):
        for p in pt:
            p.grad = torch.randn_like(p)
        ot.step()
    of = FlatAdam(pf, lr=1e-3)
    assert isinstance(of, torch.optim.Optimizer) and set(ot.param_groups[0]) <= set(of.param_groups[0]) | {"initial_lr"}
    assert of.state_dict()["state"] == {}              # like torch.optim.Adam before its first step
    sched = torch.optim.lr_scheduler.StepLR(of, 1, gamma=0.5)   # cpc/train.py:351-355 needs param_groups / 'lr'
    import io
    buf = io.BytesIO()
    torch.save(ot.state_dict(), buf)
    buf.seek(0)
    of.load_state_dict(torch.load(buf))
    g = of.param_groups[0]
    assert g["lr"] == 3e-4 and tuple(g["betas"]) == (0.8, 0.95) and g["eps"] == 1e-7 and of.steps == 3
    sizes = [p.numel() for p in pf]
    for p, m, v in zip(pt, of.exp_avg.split(sizes), of.exp_avg_sq.split(sizes)):
        assert torch.equal(m.view_as(p), ot.state[p]["exp_avg"]) and torch.equal(v.view_as(p), ot.state[p]["exp_avg_sq"])
    sd = of.state_dict()
    assert sorted(sd["state"]) == [0, 1, 2] and float(sd["state"][1]["step"]) == 3.0
    ot2 = torch.optim.Adam([torch.nn.Parameter(p.detach().clone()) for p in pf], lr=1.0)
    ot2.load_state_dict(sd)                              # and back into torch.optim.Adam
    assert ot2.param_groups[0]["lr"] == 3e-4 and float(ot2.state[ot2.param_groups[0]["params"][2]]["step"]) == 3.0
    of.bucket.detach()
    del sched


def test_flat_bucket_allreduce_gloo_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out.get(r) for r in range(world))


def test_split_ranges_for_overlapped_allreduce():
    """GradBucket.setup_overlap: the late (conv0 / batchNorm0) tensors sit in the middle of the flat bucket; the early ranges
    on both sides and the late range must tile it exactly, adjacent tensors merged."""
    from cpc_audio_b200.optim import split_ranges
    sizes = [786432, 2560, 256, 256, 256, 524288, 256, 196608, 768]
    late = [False, True, True, True, True, False, False, False, False]
    early, lt = split_ranges(sizes, late)
    assert lt == [(786432, 786432 + 3328)]
    assert early == [(0, 786432), (786432 + 3328, sum(sizes))]
    covered = sorted(early + lt)
    assert covered[0][0] == 0 and covered[-1][1] == sum(sizes) and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    assert split_ranges([4, 4], [True, False]) == ([(4, 8)], [(0, 4)])
