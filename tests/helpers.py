"""Shared test helpers: golden fixtures, module construction from oracle parameters."""
import os

import numpy as np
import torch

from oracle import cpc_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["small", "small2l", "cfg1", "cfg1_scaled"]
T_CASES = ["cfg4_small", "cfg4"]  # rnnMode='transformer' prediction heads


def case_heads(g):
    return str(g["heads"]) if "heads" in g.files else "linear"


def load_case(name):
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    B, L, H, Har, K, N, nL = [int(v) for v in g["dims"]]
    d = O.Dims(B=B, L=L, H=H, Har=Har, K=K, N=N, nLayers=nL)
    seed = int(g["seed"])
    mp, cp = O.make_params(d, seed=seed, pred_scale=float(g["pred_scale"]))
    if case_heads(g) == "transformer":
        cp = O.make_params_transformer(d, seed=seed, out_scale=float(g["pred_scale"]))
    x, label = O.make_batch(d, seed=1234 + seed)
    bi, si = O.make_raw_indices(d, seed=4321 + seed)
    return g, d, mp, cp, x, label, bi, si


def subsample(t, n=4096):
    f = t.detach().reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step][:n].cpu().numpy().copy()


def oracle_run(d, mp, cp, x, bi, si, materialize=True, heads="linear"):
    """fwd + bwd of the CPU oracle; returns dict(c, z, losses, acc, grads{model.*, crit.*})."""
    mp = {k: v.clone().requires_grad_(True) for k, v in mp.items()}
    cp = {k: v.clone().requires_grad_(True) for k, v in cp.items()}
    c, z = O.model_forward(x, mp, d.nLayers)
    losses, acc, logits = O.criterion_forward(c, z, cp, bi, si, d.K, d.N, materialize=materialize, heads=heads)
    losses.sum().backward()
    grads = {f"model.{k}": v.grad for k, v in mp.items()}
    grads.update({f"crit.{k}": v.grad for k, v in cp.items()})
    return dict(c=c.detach(), z=z.detach(), losses=losses.detach(), acc=acc.detach(), grads=grads, logits=logits)


def build_modules(d, mp, cp, dtype, device="cuda", heads="linear"):
    import cpc_audio_b200 as M
    enc = M.CPCEncoder(d.H, "layerNorm", compute_dtype=dtype)
    ar = M.CPCAR(d.H, d.Har, False, d.nLayers, mode="GRU", reverse=False, compute_dtype=dtype)
    model = M.CPCModel(enc, ar)
    crit = M.CPCUnsupersivedCriterion(d.K, d.Har, d.H, d.N, mode=None, rnnMode=heads, dropout=False,
                                      speakerEmbedding=0, nSpeakers=0, sizeInputSeq=d.S, compute_dtype=dtype)
    model.load_state_dict(mp, strict=True)
    missing = crit.load_state_dict(cp, strict=False)
    assert not missing.unexpected_keys and all(k.endswith(("Att.z", "Att.mask")) for k in missing.missing_keys), missing
    return model.to(device), crit.to(device)


def run_modules(model, crit, x, label, bi, si):
    """fwd + bwd through the B200 modules with the negative draws forced to (bi, si)."""
    dev = next(model.parameters()).device
    crit.sampleIndices = lambda B, W, S, device: (bi.to(device), si.to(device))
    if getattr(crit.wPrediction, "transformer", False):
        crit.eval()  # the heads' dropout makes train() mode non-deterministic; parity is defined in eval mode
    model.zero_grad(set_to_none=True)
    crit.zero_grad(set_to_none=True)
    c, z, _ = model(x.to(dev), label.to(dev))
    losses, acc = crit(c, z, label.to(dev))
    losses.sum().backward()
    grads = {f"model.{k}": v.grad for k, v in model.named_parameters()}
    grads.update({f"crit.{k}": v.grad for k, v in crit.named_parameters()})
    return dict(c=c.detach(), z=z.detach(), losses=losses.detach(), acc=acc.detach(), grads=grads)


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def max_rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def acc_tolerance(logits_list, d, gap=1e-4):
    """Per-step bound on |acc - acc_ref|: the fraction of rows whose best negative is within `gap` of the positive
    (rel. to the logit scale).  Exact ties happen by construction: the positive z[b, w+k] can be drawn as a negative."""
    tol = []
    for lg in logits_list:
        lg = lg.detach().float().cpu()
        scale = lg.abs().max().clamp_min(1e-6)
        near = ((lg[:, 1:].max(1)[0] - lg[:, 0]).abs() <= gap * scale).float().sum()
        tol.append(near / lg.shape[0] + 1e-7)
    return torch.stack(tol).view(1, -1)
