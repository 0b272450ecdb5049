"""CPU oracle for the CPC training-step hot path.  *** TEST INFRASTRUCTURE ONLY ***

This file is a CPU restatement of the reference algorithm (facebookresearch/CPC_audio @ b98a1bd).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it - and only as the checker / the timed CPU baseline, never as part of the product path
(``cpc_audio_b200`` never imports ``oracle``; it raises if the CUDA library is missing).

Where the arithmetic lives: the reference is pure Python on top of PyTorch (unpinned in
``environment.yml:8``); every op on the path is a stock torch op.  The restatement below therefore spells the
published algorithms out in fp32 torch / int64 numpy and cites the reference call site it follows:

  channel_norm          cpc/model.py:50-58      (unbiased variance, eps inside rsqrt, affine (1,C,1))
  encoder_forward       cpc/model.py:83-105     (5 strided Conv1d, each + ChannelNorm + ReLU)
  gru_forward           cpc/model.py:175-176,185-204 -> torch.nn.GRU equations, gate order (r, z, n)
  model_forward         cpc/model.py:286-289
  ext_indices_np        cpc/criterion/criterion.py:181-199  (integer arithmetic, numpy int64)
  criterion_forward     cpc/criterion/criterion.py:97-118, 174-219, 225-257

Pinning: the reference's own tests hold no golden vector on this path, so the oracle is
pinned against outputs of the *unmodified reference modules* imported in the authoring container
(``oracle/make_golden.py`` -> ``tests/golden/*.npz``, checked by ``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

# (kernel, stride, padding) of conv0..conv4 - cpc/model.py:83-92
CONV_GEOMETRY = ((10, 5, 3), (8, 4, 2), (4, 2, 1), (4, 2, 1), (4, 2, 1))
DOWNSAMPLING = 160  # cpc/model.py:94


@dataclass
class Dims:
    B: int = 2
    L: int = 20480
    H: int = 256          # hiddenEncoder
    Har: int = 256        # hiddenGar
    K: int = 12           # nPredicts
    N: int = 128          # negativeSamplingExt
    nLayers: int = 1      # nLevelsGRU

    @property
    def S(self):
        return self.L // DOWNSAMPLING

    @property
    def W(self):
        return self.S - self.K


# --------------------------------------------------------------------------------------------------------------
# deterministic parameters / inputs (shared by make_golden.py and the tests; independent of torch's module inits)
# --------------------------------------------------------------------------------------------------------------

def make_params(d: Dims, seed: int = 0, pred_scale: float = 1.0):
    """Deterministic fp32 parameters keyed exactly like the reference state_dicts.

    ``pred_scale`` multiplies the prediction-head weights: at the reference init every logit is ~0 and every
    loss is ln(N+1) whatever the gather does, so parity tests use a scaled set as well.
    """
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std

    model, crit = {}, {}
    cin = 1
    for i, (k, _, _) in enumerate(CONV_GEOMETRY):
        model[f"gEncoder.conv{i}.weight"] = rn(d.H, cin, k, std=1.0 / math.sqrt(cin * k))
        model[f"gEncoder.conv{i}.bias"] = rn(d.H, std=0.1)
        model[f"gEncoder.batchNorm{i}.weight"] = 1.0 + rn(1, d.H, 1, std=0.1)
        model[f"gEncoder.batchNorm{i}.bias"] = rn(1, d.H, 1, std=0.1)
        cin = d.H
    hin = d.H
    for l in range(d.nLayers):
        s = 1.0 / math.sqrt(d.Har)
        model[f"gAR.baseNet.weight_ih_l{l}"] = rn(3 * d.Har, hin, std=s)
        model[f"gAR.baseNet.weight_hh_l{l}"] = rn(3 * d.Har, d.Har, std=s)
        model[f"gAR.baseNet.bias_ih_l{l}"] = rn(3 * d.Har, std=s)
        model[f"gAR.baseNet.bias_hh_l{l}"] = rn(3 * d.Har, std=s)
        hin = d.Har
    for k in range(d.K):
        crit[f"wPrediction.predictors.{k}.weight"] = rn(d.H, d.Har) * pred_scale
    return model, crit


T_HEADS, T_DFF = 8, 2048  # cpc/transformers.py:98-99 defaults (nheads=8, dff=2048)


def make_params_transformer(d: Dims, seed: int = 0, out_scale: float = 1.0):
    """Deterministic parameters of the K one-layer transformer prediction heads (rnnMode='transformer',
    criterion.py:82-88 -> transformers.py:129-139), keyed like the reference state_dict.  Requires Har == H."""
    assert d.H == d.Har
    g = torch.Generator().manual_seed(seed + 7919)

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std

    D, dk, W = d.H, d.H // T_HEADS, d.W
    crit = {}
    for k in range(d.K):
        pre = f"wPrediction.predictors.{k}.0."
        for nm in ("Wo", "Wk", "Wq", "Wv"):
            crit[pre + f"multihead.{nm}.weight"] = rn(D, D, std=1.0 / math.sqrt(D))
        crit[pre + "multihead.Att.Krelpos"] = rn(dk, W, std=1.0 / math.sqrt(dk))
        crit[pre + "ln_multihead.weight"] = 1.0 + rn(D, std=0.1)
        crit[pre + "ln_multihead.bias"] = rn(D, std=0.1)
        crit[pre + "ffnetwork.lin1.weight"] = rn(T_DFF, D, std=1.0 / math.sqrt(D))
        crit[pre + "ffnetwork.lin1.bias"] = rn(T_DFF, std=0.1)
        crit[pre + "ffnetwork.lin2.weight"] = rn(D, T_DFF, std=1.0 / math.sqrt(T_DFF))
        crit[pre + "ffnetwork.lin2.bias"] = rn(D, std=0.1)
        crit[pre + "ln_ffnetwork.weight"] = (1.0 + rn(D, std=0.1)) * out_scale
        crit[pre + "ln_ffnetwork.bias"] = rn(D, std=0.1) * out_scale
    return crit


def transformer_head_forward(x, p, pre):
    """One TransformerLayer in eval mode (dropout = identity): transformers.py:38-49 (attention with the relative-
    position skew), 76-83 (multi-head), 86-95 (FFN), 109-111 (post-LN residuals).  x (B, W, D) -> (B, W, D)."""
    B, W, D = x.shape
    nh, dk = T_HEADS, D // T_HEADS

    def heads(t):  # transformers.py:67-69
        return t.view(B, W, nh, dk).transpose(1, 2).contiguous().view(B * nh, W, dk)

    q = heads(x @ p[pre + "multihead.Wq.weight"].t())
    k = heads(x @ p[pre + "multihead.Wk.weight"].t())
    v = heads(x @ p[pre + "multihead.Wv.weight"].t())
    qk = torch.bmm(q, k.transpose(-2, -1))
    qp = q.matmul(p[pre + "multihead.Att.Krelpos"])                         # (B*nh, W, W)
    i = torch.arange(W).view(W, 1)
    c = torch.arange(W).view(1, W)
    idx = (W - 1 - i + c).clamp(0, W - 1)                                   # skew: key c <= query i reads column W-1-(i-c)
    qk = qk + torch.gather(qp, 2, idx.unsqueeze(0).expand(B * nh, W, W))
    mask = torch.zeros(W, W).masked_fill(c > i, float("-inf"))            # transformers.py:29-32 (causal)
    a = torch.softmax(qk / math.sqrt(dk) + mask, dim=2)
    y = torch.bmm(a, v).view(B, nh, W, dk).transpose(1, 2).contiguous().view(B, W, D)
    y = y @ p[pre + "multihead.Wo.weight"].t()
    y1 = F.layer_norm(x + y, (D,), p[pre + "ln_multihead.weight"], p[pre + "ln_multihead.bias"], 1e-5)
    f = torch.relu(y1 @ p[pre + "ffnetwork.lin1.weight"].t() + p[pre + "ffnetwork.lin1.bias"])
    f = f @ p[pre + "ffnetwork.lin2.weight"].t() + p[pre + "ffnetwork.lin2.bias"]
    return F.layer_norm(y1 + f, (D,), p[pre + "ln_ffnetwork.weight"], p[pre + "ln_ffnetwork.bias"], 1e-5)


def make_batch(d: Dims, seed: int = 1234):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(d.B, 1, d.L, generator=g, dtype=torch.float32) * 0.1
    label = torch.zeros(d.B, dtype=torch.int64)
    return x, label


def make_raw_indices(d: Dims, seed: int = 4321):
    """The two ``torch.randint`` draws of criterion.py:181-189, CPU generator, same order/shape/dtype."""
    g = torch.Generator().manual_seed(seed)
    n = d.N * d.W * d.B
    batch_idx = torch.randint(low=0, high=d.B, size=(n,), generator=g)
    seq_idx = torch.randint(low=1, high=d.S, size=(n,), generator=g)
    return batch_idx, seq_idx


# --------------------------------------------------------------------------------------------------------------
# model
# --------------------------------------------------------------------------------------------------------------

def channel_norm(x, weight, bias, eps=1e-5):
    """cpc/model.py:50-58 - normalise across channels (dim 1) with the *unbiased* variance."""
    mean = x.mean(dim=1, keepdim=True)
    var = x.var(dim=1, keepdim=True)            # N-1 divisor (model.py:53)
    x = (x - mean) * torch.rsqrt(var + eps)     # eps inside rsqrt (model.py:54)
    return x * weight + bias


def encoder_forward(x, p, prefix="gEncoder."):
    """cpc/model.py:99-105.  x (B,1,L) -> (B,H,S)."""
    for i, (_, s, pad) in enumerate(CONV_GEOMETRY):
        x = F.conv1d(x, p[f"{prefix}conv{i}.weight"], p[f"{prefix}conv{i}.bias"], stride=s, padding=pad)
        x = F.relu(channel_norm(x, p[f"{prefix}batchNorm{i}.weight"], p[f"{prefix}batchNorm{i}.bias"]))
    return x


def gru_forward(z, p, n_layers=1, h0=None, prefix="gAR.baseNet."):
    """torch.nn.GRU(batch_first=True) restated (reference call site cpc/model.py:193).

    r = sigma(W_ir x + b_ir + W_hr h + b_hr); z = sigma(W_iz x + b_iz + W_hz h + b_hz)
    n = tanh(W_in x + b_in + r * (W_hn h + b_hn)); h' = (1 - z) * n + z * h      (gate packing r,z,n)
    Returns (c (B,S,Har), hT (nLayers,B,Har)).
    """
    B, S, _ = z.shape
    inp = z
    h_last = []
    for l in range(n_layers):
        w_ih, w_hh = p[f"{prefix}weight_ih_l{l}"], p[f"{prefix}weight_hh_l{l}"]
        b_ih, b_hh = p[f"{prefix}bias_ih_l{l}"], p[f"{prefix}bias_hh_l{l}"]
        Har = w_hh.shape[1]
        h = torch.zeros(B, Har, dtype=z.dtype) if h0 is None else h0[l]
        gi_all = inp @ w_ih.t() + b_ih
        outs = []
        for t in range(S):
            gi = gi_all[:, t]
            gh = h @ w_hh.t() + b_hh
            r = torch.sigmoid(gi[:, :Har] + gh[:, :Har])
            u = torch.sigmoid(gi[:, Har:2 * Har] + gh[:, Har:2 * Har])
            n = torch.tanh(gi[:, 2 * Har:] + r * gh[:, 2 * Har:])
            h = (1.0 - u) * n + u * h
            outs.append(h)
        inp = torch.stack(outs, dim=1)
        h_last.append(h)
    return inp, torch.stack(h_last, dim=0)


def model_forward(x, p, n_layers=1, h0=None):
    """cpc/model.py:286-289 -> (cFeature (B,S,Har), encodedData (B,S,H))."""
    z = encoder_forward(x, p).permute(0, 2, 1)
    c, _ = gru_forward(z, p, n_layers, h0)
    return c, z


# --------------------------------------------------------------------------------------------------------------
# criterion
# --------------------------------------------------------------------------------------------------------------

def ext_indices_np(batch_idx, seq_idx, B, N, W, S):
    """criterion.py:191-199 in int64 numpy.  Flat layout (B, N, W): i <-> (b=i//(N*W), n=(i//W)%N, w=i%W).

    seq = (seq_raw + w) mod S   (never equals w because seq_raw in [1, S));   ext = seq + batch * S.
    """
    batch_idx = np.asarray(batch_idx, dtype=np.int64)
    seq_idx = np.asarray(seq_idx, dtype=np.int64)
    assert batch_idx.shape == seq_idx.shape == (B * N * W,)
    w = np.arange(B * N * W, dtype=np.int64) % W
    seq = np.remainder(seq_idx + w, S)
    return (seq + batch_idx * S).reshape(B, N, W)


def criterion_forward(c, z, crit_p, batch_idx, seq_idx, K, N, materialize=True, heads="linear"):
    """criterion.py:225-257 with the linear heads of criterion.py:89-95,106-117.

    c (B,S,Har), z (B,S,H) fp32.  Returns (losses (1,K), acc (1,K), logits list of (B*W, N+1)).
    ``materialize=True`` follows the reference op for op (cat of K candidate tensors, broadcast mul, mean);
    ``False`` computes the same numbers with einsum (used for the big-shape checks).
    """
    B, S, H = z.shape
    W = S - K
    ext = torch.from_numpy(ext_indices_np(batch_idx.numpy(), seq_idx.numpy(), B, N, W, S)).reshape(-1)
    neg = z.contiguous().view(-1, H)[ext].view(B, N, W, H)            # criterion.py:200-201
    cw = c[:, :W]                                                       # criterion.py:234
    label = torch.zeros(B * W, dtype=torch.long)                        # criterion.py:203-205
    losses, accs, all_logits = [], [], []
    for k in range(1, K + 1):
        pos = z[:, k:k + W].reshape(B, 1, W, H)                        # criterion.py:207-215
        if heads == "linear":
            pred = cw @ crit_p[f"wPrediction.predictors.{k - 1}.weight"].t()   # criterion.py:108
        else:                                                                   # criterion.py:82-88 (eval mode)
            pred = transformer_head_forward(cw, crit_p, f"wPrediction.predictors.{k - 1}.0.")
        if materialize:
            full = torch.cat((pos, neg), dim=1)                        # criterion.py:216
            out = (pred.view(B, 1, W, H) * full).mean(dim=3)           # criterion.py:115-116
        else:
            out = torch.cat((torch.einsum("bwd,bwd->bw", pred, pos[:, 0]).unsqueeze(1),
                             torch.einsum("bwd,bnwd->bnw", pred, neg)), dim=1) / H
        logits = out.permute(0, 2, 1).contiguous().view(-1, N + 1)     # criterion.py:249-250
        losses.append(F.cross_entropy(logits, label).view(1, 1))       # criterion.py:251
        accs.append((logits.max(1)[1] == label).sum().float().view(1, 1))  # criterion.py:253-254
        all_logits.append(logits)
    return torch.cat(losses, dim=1), torch.cat(accs, dim=1) / (W * B), all_logits


def train_step_reference_style(x, model_p, crit_p, batch_idx, seq_idx, d: Dims):
    """fwd(model) + fwd(criterion) + backward, mirroring cpc/train.py:83-87 (no optimizer).

    Uses torch's own fused ops where the reference does (F.conv1d, torch GRU kernel via nn.GRU weights) so the
    timing of this function is a faithful stand-in for the reference's CPU path.
    """
    import torch.nn as nn
    z = encoder_forward(x, model_p).permute(0, 2, 1)
    gru = nn.GRU(d.H, d.Har, num_layers=d.nLayers, batch_first=True)
    # share storage with model_p so that gradients land on the dict's tensors
    flat = []
    for l in range(d.nLayers):
        for nm in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
            flat.append(model_p[f"gAR.baseNet.{nm}_l{l}"])
    c = torch._VF.gru(z, torch.zeros(d.nLayers, d.B, d.Har), flat, True, d.nLayers, 0.0, True, False, True)[0]
    losses, acc, _ = criterion_forward(c, z, crit_p, batch_idx, seq_idx, d.K, d.N, materialize=True)
    losses.sum().backward()
    return losses.detach(), acc.detach()
