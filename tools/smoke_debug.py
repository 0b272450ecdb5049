"""Development probe: f32 path, encoder + GRU + criterion gradients vs the CPU oracle for several batch sizes / seeds."""
import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from oracle import cpc_oracle as O
from tests import helpers as Hh
for B, seed, scale in ((1, 3, 30.0), (2, 3, 30.0), (2, 3, 1.0), (2, 11, 30.0), (3, 3, 30.0)):
    d = O.Dims(B=B, L=20480, H=256, Har=256, K=12, N=128, nLayers=1)
    mp, cp = O.make_params(d, seed=seed, pred_scale=scale)
    x, label = O.make_batch(d, seed=5)
    bi, si = O.make_raw_indices(d, seed=6)
    ref = Hh.oracle_run(d, mp, cp, x, bi, si)
    model, crit = Hh.build_modules(d, mp, cp, "f32")
    out = Hh.run_modules(model, crit, x, label, bi, si)
    keys = ["model.gEncoder.batchNorm4.bias", "model.gEncoder.batchNorm3.bias", "model.gEncoder.batchNorm2.bias", "model.gEncoder.batchNorm2.weight",
            "model.gEncoder.conv2.weight", "model.gEncoder.batchNorm1.bias", "model.gEncoder.conv0.weight"]
    print(f"B={B} seed={seed} scale={scale}:", [f"{Hh.rel_err(out['grads'][k], ref['grads'][k]):.1e}" for k in keys], flush=True)
    g, r = out["grads"]["model.gEncoder.batchNorm2.bias"].flatten().cpu(), ref["grads"]["model.gEncoder.batchNorm2.bias"].flatten()
    dlt = (g - r)
    print("   bn2.bias: max|d| %.2e at ch %d, |ref| max %.2e; #ch with |d|>1e-3*max: %d" % (dlt.abs().max(), dlt.abs().argmax(), r.abs().max(), int((dlt.abs() > 1e-3 * r.abs().max()).sum())))
