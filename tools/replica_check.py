"""torchrun --nproc-per-node N tools/replica_check.py : 50 data-parallel training steps of the whole CPC step (whole-step CUDA
graph, PeerAdam: gradient exchange + Adam in one kernel) with per-rank data and per-rank negatives; afterwards every rank's
parameters must be BIT-identical to rank 0's (the exchange hands every rank the same sums, cpc/train.py:85 semantics), and
one step's all-reduced gradient must equal an NCCL all-reduce of the per-rank gradients.  Development / evidence tool."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpc_audio_b200 as M  # noqa: E402
from cpc_audio_b200.graph import GraphedTrainStep  # noqa: E402
from cpc_audio_b200.optim import PeerAdam  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
model = M.CPCModel(M.CPCEncoder(256, "layerNorm"), M.CPCAR(256, 256, False, 1, mode="GRU")).to(dev)
crit = M.CPCUnsupersivedCriterion(12, 256, 256, 128, rnnMode="linear", sizeInputSeq=128).to(dev)
params = list(crit.parameters()) + list(model.parameters())
enc = model.gEncoder
overlap = os.environ.get("CPC_B200_PEER_OVERLAP", "0") == "1"
opt = PeerAdam(params, lr=2e-4, fuse_zero_grad=True, overlap=overlap,
               late_params=[enc.conv0.weight, enc.conv0.bias, enc.batchNorm0.weight, enc.batchNorm0.bias])
gen = torch.Generator(device=dev).manual_seed(1234 + rank)
torch.cuda.manual_seed(4321 + rank)
B = 16
x = torch.randn(B, 1, 20480, device=dev, generator=gen) * 0.1
label = torch.zeros(B, dtype=torch.long, device=dev)
step = GraphedTrainStep(model, crit, opt, x, label, warmup=3, before_backward=opt.arm_overlap if overlap else None)
losses = []
for i in range(50):
    xb = torch.randn(B, 1, 20480, device=dev, generator=gen) * 0.1
    losses.append(step(xb)[0].mean().item())
opt.check()
flat = opt.flat_p.clone()
ref = flat.clone()
dist.broadcast(ref, 0)
same = bool(torch.equal(flat, ref))
allsame = torch.tensor([int(same)], device=dev)
dist.all_reduce(allsame, op=dist.ReduceOp.MIN)
print(f"rank {rank}: loss {losses[0]:.4f} -> {losses[-1]:.4f}, parameters bit-identical to rank 0 after {opt.steps} steps: {same}", flush=True)
if rank == 0:
    print(f"world {world} overlap={overlap}: replicas bit-identical on every rank: {bool(allsame.item())}", flush=True)
assert allsame.item() == 1
step.graph.reset()
torch.cuda.synchronize(); dist.barrier()
sys.stdout.flush()
os._exit(0)
