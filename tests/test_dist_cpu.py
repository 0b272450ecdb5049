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
