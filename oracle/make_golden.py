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
sys.path.insert(0, REPO)

from oracle import cpc_oracle as O  # noqa: E402
from oracle.ref_import import import_reference as _import_reference  # noqa: E402


def import_reference():
    ref = _import_reference(os.environ.get("CPC_REFERENCE"))
    if ref is None:
        raise SystemExit("the reference package is not importable here (/root/reference or baseline/_ref)")
    return ref.model, ref.criterion


def build_reference_ar(ref_model, d, ar):
    """feature_loader.py:137-152 (getAR) without the argparse namespace."""
    if ar == "transformer":
        import cpc.transformers as ref_tr
        return ref_tr.buildTransformerAR(d.H, 1, d.S, False)
    return ref_model.CPCAR(d.H, d.Har, False, d.nLayers, mode=ar, reverse=False)


def run_reference(d: O.Dims, pred_scale: float, seed: int, heads="linear", ar="GRU", train_dropout=False):
    ref_model, ref_crit = import_reference()
    torch.manual_seed(0)
    enc = ref_model.CPCEncoder(d.H, "layerNorm")
    model = ref_model.CPCModel(enc, build_reference_ar(ref_model, d, ar))
    crit = ref_crit.CPCUnsupersivedCriterion(d.K, d.Har, d.H, d.N, mode=None, rnnMode=heads, dropout=False,
                                             speakerEmbedding=0, nSpeakers=0, sizeInputSeq=d.S)
    mp, cp = O.make_params(d, seed=seed, pred_scale=pred_scale, ar=ar)
    if heads == "transformer":
        cp = O.make_params_transformer(d, seed=seed, out_scale=pred_scale)
    missing = model.load_state_dict(mp, strict=False)  # a transformer context net holds the constant buffers Att.z / Att.mask
    assert not missing.unexpected_keys and all(k.endswith(("Att.z", "Att.mask")) for k in missing.missing_keys), missing
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

    # train-mode dropout (transformers.py:18,49,92): nn.Dropout.forward is replaced by one that applies the seeded keep-masks
    # of O.make_dropout_masks in call order (context network first, then head by head: attention, FFN) - this pins WHERE
    # the reference applies dropout, the 1/(1-p) scale and the order of the draws
    ar_masks, head_masks = O.make_dropout_masks(d, seed, ar=ar, heads=heads) if train_dropout else (None, None)
    queue = []
    if ar_masks is not None:
        queue += [ar_masks[0], ar_masks[1]]
    if head_masks is not None:
        for k in range(d.K):
            queue += [head_masks[0][k], head_masks[1][k]]
    real_dropout_forward = torch.nn.Dropout.forward

    def fake_dropout_forward(self, inp):
        if not self.training:
            return inp
        keep = queue.pop(0)
        assert keep.numel() == inp.numel() and abs(self.p - 0.1) < 1e-12, (keep.shape, inp.shape, self.p)
        return inp * keep.to(inp.dtype).view(inp.shape) / (1.0 - self.p)

    torch.randint = fake_randint
    torch.nn.Dropout.forward = fake_dropout_forward
    try:
        model.train()
        crit.train()
        if not train_dropout:
            crit.eval()   # dropout 0.1 inside the transformer layers (transformers.py:18,92): eval-mode fixture
            if ar == "transformer":
                model.eval()
        c, z, _ = model(x, label)
        losses, acc = crit(c, z, label)
    finally:
        torch.randint = real_randint
        torch.nn.Dropout.forward = real_dropout_forward
    assert not draws and not queue
    losses.sum().backward()
    grads = {f"model.{k}": v.grad.detach().clone() for k, v in model.named_parameters()}
    grads.update({f"crit.{k}": v.grad.detach().clone() for k, v in crit.named_parameters()})
    return dict(x=x, bi=bi, si=si, c=c.detach(), z=z.detach().contiguous(), losses=losses.detach(),
                acc=acc.detach(), grads=grads, mp=mp, cp=cp)


def check_oracle(d, r, heads="linear", ar="GRU", train_dropout=False, seed=0):
    """The restatement must reproduce the reference (this is what pins the oracle)."""
    mp = {k: v.clone().requires_grad_(True) for k, v in r["mp"].items()}
    cp = {k: v.clone().requires_grad_(True) for k, v in r["cp"].items()}
    ar_masks, head_masks = O.make_dropout_masks(d, seed, ar=ar, heads=heads) if train_dropout else (None, None)
    c, z = O.model_forward(r["x"], mp, d.nLayers, ar_masks=ar_masks)
    losses, acc, _ = O.criterion_forward(c, z, cp, r["bi"], r["si"], d.K, d.N, heads=heads, head_masks=head_masks)
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


def save(name, d, r, pred_scale, seed, heads="linear", ar="GRU", train_dropout=False):
    out = dict(heads=np.array(heads), ar=np.array(ar), train_dropout=np.int64(int(train_dropout)), dims=np.array([d.B, d.L, d.H, d.Har, d.K, d.N, d.nLayers], dtype=np.int64),
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

# context networks other than the GRU and train-mode dropout (row T): name -> (Dims, scale, seed, heads, ar, train)
CASES2 = {
    "lstm_small": (O.Dims(B=3, L=3200, H=128, Har=64, K=5, N=16, nLayers=2), 30.0, 7, "linear", "LSTM", False),
    "cfg1_lstm": (O.Dims(B=2, L=20480, H=256, Har=256, K=12, N=128, nLayers=1), 30.0, 8, "linear", "LSTM", False),   # reference default arMode
    "tar_small": (O.Dims(B=3, L=3200, H=64, Har=64, K=3, N=8, nLayers=1), 30.0, 9, "linear", "transformer", False),
    "cfg1_tar": (O.Dims(B=2, L=20480, H=256, Har=256, K=12, N=128, nLayers=1), 30.0, 10, "linear", "transformer", False),
    "cfg4_small_train": (O.Dims(B=3, L=3200, H=64, Har=64, K=3, N=8, nLayers=1), 30.0, 11, "transformer", "GRU", True),
    "cfg4_train": (O.Dims(B=2, L=20480, H=256, Har=256, K=12, N=128, nLayers=1), 30.0, 12, "transformer", "GRU", True),
    "tar_small_train": (O.Dims(B=2, L=3200, H=64, Har=64, K=3, N=8, nLayers=1), 30.0, 13, "transformer", "transformer", True),
    # BASELINE config 5 widths (hidden 512, 2 levels, K = 16, 256 negatives) on a 20480-sample window, and a 2-level LSTM at 512:
    # they pin the ORACLE at the wide dims (tests/test_oracle_golden.py); the CUDA path is compared with the oracle at these dims
    "cfg5_s128": (O.Dims(B=2, L=20480, H=512, Har=512, K=16, N=256, nLayers=2), 30.0, 14, "linear", "GRU", False),
    "lstm512": (O.Dims(B=2, L=10240, H=512, Har=512, K=8, N=32, nLayers=2), 30.0, 15, "linear", "LSTM", False),
}


def make_feature_fixture():
    """cpc/feature_loader.py: FeatureModule + buildFeature, UNMODIFIED, on CPU: torchaudio.load is replaced by a function
    returning a seeded waveform and Tensor.cuda by the identity (the only device-specific calls of that code path)."""
    ref_model, _ = import_reference()
    import cpc.feature_loader as fl
    out = {}
    for name, ar, nl, n, chunk in (("feat_gru", "GRU", 1, 50000, 20480), ("feat_lstm", "LSTM", 2, 33333, 12345)):
        d = O.Dims(B=1, L=chunk, H=64, Har=64, nLayers=nl)
        mp, _ = O.make_params(d, seed=21, ar=ar)
        seq = torch.randn(n, generator=torch.Generator().manual_seed(22)) * 0.1
        model = ref_model.CPCModel(ref_model.CPCEncoder(d.H, "layerNorm"), ref_model.CPCAR(d.H, d.Har, True, nl, mode=ar, reverse=False))
        model.load_state_dict(mp, strict=True)
        model.eval()
        fm = fl.FeatureModule(model, False)
        real_load, real_cuda = fl.torchaudio.load, torch.Tensor.cuda
        fl.torchaudio.load = lambda path: (seq.view(1, -1), 16000)
        torch.Tensor.cuda = lambda self, *a, **k: self
        try:
            feat = fl.buildFeature(fm, "seeded.wav", strict=False, maxSizeSeq=chunk)
        finally:
            fl.torchaudio.load, torch.Tensor.cuda = real_load, real_cuda
        mine = O.feature_forward(seq, mp, nl, max_size_seq=chunk, keep_hidden=True)
        err = (feat - mine).abs().max().item()
        print(f"[{name}] {tuple(feat.shape)} oracle-vs-reference buildFeature max err {err:.2e}")
        assert feat.shape == mine.shape and err < 1e-5
        out[name] = dict(ar=np.array(ar), nLayers=np.int64(nl), n=np.int64(n), chunk=np.int64(chunk), H=np.int64(d.H),
                         feat_sub=subsample(feat), feat_norm=np.float64(feat.double().norm().item()),
                         shape=np.array(feat.shape, dtype=np.int64))
    for name, v in out.items():
        path = os.path.join(REPO, "tests", "golden", f"{name}.npz")
        np.savez_compressed(path, **v)
        print(f"  wrote {path}")


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
    for name, (d, ps, seed, heads, ar, train) in CASES2.items():
        if only and name not in only:
            continue
        print(f"[{name}] {d} pred_scale={ps} heads={heads} ar={ar} train_dropout={train}")
        r = run_reference(d, ps, seed, heads, ar=ar, train_dropout=train)
        print("  losses", np.round(r["losses"].numpy().ravel(), 4), "\n  acc", np.round(r["acc"].numpy().ravel(), 4))
        check_oracle(d, r, heads, ar=ar, train_dropout=train, seed=seed)
        save(name, d, r, ps, seed, heads, ar=ar, train_dropout=train)
    if not only or "features" in only:
        make_feature_fixture()
