"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules (authoring container only).

    python oracle/make_golden.py            # needs /root/reference; writes tests/golden/{small,cfg1,cfg1_scaled}.npz

The reference is Python and cannot travel to the GPU box, so its outputs are committed as small fixtures.
Parameters / inputs / negative-sample draws come from the seeded generators in ``oracle/cpc_oracle.py`` (so the
tests can rebuild them bit-for-bit without the reference); the reference modules get them through
``load_state_dict`` and through a recording wrapper around ``torch.randint`` (criterion.py:181-189 calls it
twice, batchIdx then seqIdx).  Import recipe: SURVEY.md 8(c) - stub ``progressbar`` / ``soundfile``.
TEST INFRASTRUCTURE ONLY.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("CPC_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)

from oracle import cpc_oracle as O  # noqa: E402


def import_reference():
    for name in ("progressbar", "soundfile"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.path.insert(0, REF)
    import cpc.model as ref_model
    import cpc.transformers as ref_tr
    sys.modules["transformers"] = ref_tr  # criterion.py:83 does a bare `from transformers import ...`
    import cpc.criterion as ref_crit
    return ref_model, ref_crit


def run_reference(d: O.Dims, pred_scale: float, seed: int, heads="linear"):
    ref_model, ref_crit = import_reference()
    torch.manual_seed(0)
    enc = ref_model.CPCEncoder(d.H, "layerNorm")
    ar = ref_model.CPCAR(d.H, d.Har, False, d.nLayers, mode="GRU", reverse=False)
    model = ref_model.CPCModel(enc, ar)
    crit = ref_crit.CPCUnsupersivedCriterion(d.K, d.Har, d.H, d.N, mode=None, rnnMode=heads, dropout=False,
                                             speakerEmbedding=0, nSpeakers=0, sizeInputSeq=d.S)
    mp, cp = O.make_params(d, seed=seed, pred_scale=pred_scale)
    if heads == "transformer":
        cp = O.make_params_transformer(d, seed=seed, out_scale=pred_scale)
    missing = model.load_state_dict(mp, strict=True)
    crit.load_state_dict(cp, strict=False)  # the transformer heads also hold the constant buffers Att.z / Att.mask
    assert all(k.endswith(("Att.z", "Att.mask")) for k in set(crit.state_dict()) - set(cp)), set(crit.state_dict()) - set(cp)
    x, label = O.make_batch(d, seed=1234 + seed)
    bi, si = O.make_raw_indices(d, seed=4321 + seed)

    draws = [bi.clone(), si.clone()]
    real_randint = torch.randint

    def fake_randint(*a, **kw):
        out = draws.pop(0)
        chk = real_randint(*a, **kw)
        assert chk.shape == out.shape and chk.dtype == out.dtype
        lo, hi = kw["low"], kw["high"]
        assert int(out.min()) >= lo and int(out.max()) < hi
        return out

    torch.randint = fake_randint
    try:
        model.train()
        crit.train()
        if heads == "transformer":
            crit.eval()  # dropout 0.1 inside the heads (transformers.py:18,92): parity is defined in eval mode
        c, z, _ = model(x, label)
        losses, acc = crit(c, z, label)
    finally:
        torch.randint = real_randint
    assert not draws
    losses.sum().backward()
    grads = {f"model.{k}": v.grad.detach().clone() for k, v in model.named_parameters()}
    grads.update({f"crit.{k}": v.grad.detach().clone() for k, v in crit.named_parameters()})
    return dict(x=x, bi=bi, si=si, c=c.detach(), z=z.detach().contiguous(), losses=losses.detach(),
                acc=acc.detach(), grads=grads, mp=mp, cp=cp)


def check_oracle(d, r, heads="linear"):
    """The restatement must reproduce the reference (this is what pins the oracle)."""
    mp = {k: v.clone().requires_grad_(True) for k, v in r["mp"].items()}
    cp = {k: v.clone().requires_grad_(True) for k, v in r["cp"].items()}
    c, z = O.model_forward(r["x"], mp, d.nLayers)
    losses, acc, _ = O.criterion_forward(c, z, cp, r["bi"], r["si"], d.K, d.N, heads=heads)
    losses.sum().backward()
    errs = dict(z=(z - r["z"]).abs().max().item(), c=(c - r["c"]).abs().max().item(),
                loss=(losses - r["losses"]).abs().max().item(), acc=(acc - r["acc"]).abs().max().item())
    for k, v in mp.items():
        g = r["grads"][f"model.{k}"]
        errs[f"g.{k}"] = ((v.grad - g).abs().max() / (g.abs().max() + 1e-12)).item()
    for k, v in cp.items():
        g = r["grads"][f"crit.{k}"]
        errs[f"g.{k}"] = ((v.grad - g).abs().max() / (g.abs().max() + 1e-12)).item()
    worst = max(errs.values())
    print(f"  oracle-vs-reference worst err {worst:.3e}  (z {errs['z']:.2e} c {errs['c']:.2e} loss {errs['loss']:.2e})")
    assert errs["acc"] == 0.0 and errs["loss"] < 1e-5 and errs["z"] < 1e-4 and worst < 5e-4, errs


def subsample(t, n=4096):
    f = t.detach().reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step][:n].numpy().copy()


def save(name, d, r, pred_scale, seed, heads="linear"):
    out = dict(heads=np.array(heads), dims=np.array([d.B, d.L, d.H, d.Har, d.K, d.N, d.nLayers], dtype=np.int64),
               pred_scale=np.float32(pred_scale), seed=np.int64(seed),
               losses=r["losses"].numpy(), acc=r["acc"].numpy(),
               z_sub=subsample(r["z"]), c_sub=subsample(r["c"]),
               z_norm=np.float64(r["z"].double().norm().item()), c_norm=np.float64(r["c"].double().norm().item()),
               ext_idx=O.ext_indices_np(r["bi"].numpy(), r["si"].numpy(), d.B, d.N, d.W, d.S).astype(np.int32),
               bi_sum=np.int64(r["bi"].sum().item()), si_sum=np.int64(r["si"].sum().item()))
    for k, g in r["grads"].items():
        out[f"gsub.{k}"] = subsample(g, 512)
        out[f"gnorm.{k}"] = np.float64(g.double().norm().item())
    path = os.path.join(REPO, "tests", "golden", f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"  wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


CASES = {
    # name: (Dims, pred_scale, seed)
    "small": (O.Dims(B=3, L=2560, H=64, Har=64, K=4, N=8, nLayers=1), 30.0, 1),
    "small2l": (O.Dims(B=2, L=3200, H=128, Har=64, K=5, N=16, nLayers=2), 30.0, 2),
    "cfg1": (O.Dims(B=2, L=20480, H=256, Har=256, K=12, N=128, nLayers=1), 1.0, 3),      # BASELINE config 1
    "cfg1_scaled": (O.Dims(B=2, L=20480, H=256, Har=256, K=12, N=128, nLayers=1), 30.0, 4),
    # rnnMode='transformer' prediction heads (BASELINE config 4 dims at B=2; small variant with 64-wide model)
    "cfg4": (O.Dims(B=2, L=20480, H=256, Har=256, K=12, N=128, nLayers=1), 30.0, 5, "transformer"),
    "cfg4_small": (O.Dims(B=3, L=3200, H=64, Har=64, K=3, N=8, nLayers=1), 30.0, 6, "transformer"),
}

if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    only = sys.argv[1:]
    for name, case in CASES.items():
        if only and name not in only:
            continue
        d, ps, seed = case[:3]
        heads = case[3] if len(case) > 3 else "linear"
        print(f"[{name}] {d} pred_scale={ps} heads={heads}")
        r = run_reference(d, ps, seed, heads)
        print("  losses", np.round(r["losses"].numpy().ravel(), 4), "\n  acc", np.round(r["acc"].numpy().ravel(), 4))
        check_oracle(d, r, heads)
        save(name, d, r, ps, seed, heads)
