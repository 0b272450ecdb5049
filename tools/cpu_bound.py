import sys, time, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
sys.argv = ['bench.py']


# minimal re-implementation of the step (bench.py builds everything inside main())
import cpc_audio_b200 as M
from cpc_audio_b200.optim import FlatAdam
dev = torch.device('cuda:0')
for B in (64, 128):
    enc = M.CPCEncoder(256, 'layerNorm', compute_dtype='bf16'); ar = M.CPCAR(256, 256, False, 1, mode='GRU', reverse=False, compute_dtype='bf16')
    model = M.CPCModel(enc, ar).to(dev)
    crit = M.CPCUnsupersivedCriterion(12, 256, 256, 128, mode=None, rnnMode='linear', dropout=False, speakerEmbedding=0, nSpeakers=0, sizeInputSeq=128, compute_dtype='bf16').to(dev)
    params = list(crit.parameters()) + list(model.parameters())
    opt = FlatAdam(params, lr=2e-4)
    x = torch.randn(B, 1, 20480, device=dev) * 0.1; label = torch.zeros(B, dtype=torch.long, device=dev)
    def step():
        c, z, _ = model(x, label); losses, acc = crit(c, z, label); losses.sum().backward(); opt.step(); opt.zero_grad()
    for _ in range(5): step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"B={B}: cpu enqueue {(t1-t0)*100:.3f} ms/step, total {(t2-t0)*100:.3f} ms/step")
