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
  lstm_forward          cpc/model.py:171-173,185-204 -> torch.nn.LSTM equations, gate order (i, f, g, o)
  transformer_head_forward  cpc/transformers.py:38-49,76-95,109-111 (eval, or train with explicit dropout keep-masks)
  feature_forward       cpc/feature_loader.py:228-269 (buildFeature: chunks of maxSizeSeq samples, carried hidden state)
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

def make_params(d: Dims, seed: int = 0, pred_scale: float = 1.0, ar: str = "GRU"):
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
    ng = {"GRU": 3, "LSTM": 4, "transformer": 0}[ar]  # gate rows per hidden unit (model.py:171-179)
    for l in range(d.nLayers if ng else 0):
        s = 1.0 / math.sqrt(d.Har)
        model[f"gAR.baseNet.weight_ih_l{l}"] = rn(ng * d.Har, hin, std=s)
        model[f"gAR.baseNet.weight_hh_l{l}"] = rn(ng * d.Har, d.Har, std=s)
        model[f"gAR.baseNet.bias_ih_l{l}"] = rn(ng * d.Har, std=s)
        model[f"gAR.baseNet.bias_hh_l{l}"] = rn(ng * d.Har, std=s)
        hin = d.Har
    if ar == "transformer":  # feature_loader.py:138-142: buildTransformerAR(hiddenEncoder, 1, sizeWindow // 160, abspos)
        assert d.H == d.Har
        model.update(transformer_layer_params(d.H, d.S, "gAR.0.", seed + 104729, 1.0))
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

    crit = {}
    for k in range(d.K):
        crit.update(_transformer_layer_params(rn, d.H, d.W, f"wPrediction.predictors.{k}.0.", out_scale))
    return crit


def _transformer_layer_params(rn, D, W, pre, out_scale):
    dk = D // T_HEADS
    p = {}
    for nm in ("Wo", "Wk", "Wq", "Wv"):
        p[pre + f"multihead.{nm}.weight"] = rn(D, D, std=1.0 / math.sqrt(D))
    p[pre + "multihead.Att.Krelpos"] = rn(dk, W, std=1.0 / math.sqrt(dk))
    p[pre + "ln_multihead.weight"] = 1.0 + rn(D, std=0.1)
    p[pre + "ln_multihead.bias"] = rn(D, std=0.1)
    p[pre + "ffnetwork.lin1.weight"] = rn(T_DFF, D, std=1.0 / math.sqrt(D))
    p[pre + "ffnetwork.lin1.bias"] = rn(T_DFF, std=0.1)
    p[pre + "ffnetwork.lin2.weight"] = rn(D, T_DFF, std=1.0 / math.sqrt(T_DFF))
    p[pre + "ffnetwork.lin2.bias"] = rn(D, std=0.1)
    p[pre + "ln_ffnetwork.weight"] = (1.0 + rn(D, std=0.1)) * out_scale
    p[pre + "ln_ffnetwork.bias"] = rn(D, std=0.1) * out_scale
    return p


def transformer_layer_params(D, W, pre, seed, out_scale=1.0):
    """Deterministic parameters of ONE TransformerLayer(sizeSeq=W, dmodel=D) under the key prefix `pre`."""
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std

    return _transformer_layer_params(rn, D, W, pre, out_scale)


def transformer_head_forward(x, p, pre, att_keep=None, ffn_keep=None, p_drop=0.1):
    """One TransformerLayer: transformers.py:38-49 (attention with the relative-position skew), 76-83 (multi-head),
    86-95 (FFN), 109-111 (post-LN residuals).  x (B, W, D) -> (B, W, D).  eval mode (dropout = identity) unless the
    keep-masks of train mode are given: att_keep (B*nheads, W, W) multiplies the attention probabilities
    (transformers.py:49 self.drop(A)), ffn_keep (B, W, dff) the FFN hidden (transformers.py:95), both scaled by
    1/(1-p_drop) exactly as nn.Dropout does."""
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
    if att_keep is not None:
        a = a * att_keep.to(a.dtype).view(B * nh, W, W) / (1.0 - p_drop)
    y = torch.bmm(a, v).view(B, nh, W, dk).transpose(1, 2).contiguous().view(B, W, D)
    y = y @ p[pre + "multihead.Wo.weight"].t()
    y1 = F.layer_norm(x + y, (D,), p[pre + "ln_multihead.weight"], p[pre + "ln_multihead.bias"], 1e-5)
    f = torch.relu(y1 @ p[pre + "ffnetwork.lin1.weight"].t() + p[pre + "ffnetwork.lin1.bias"])
    if ffn_keep is not None:
        f = f * ffn_keep.to(f.dtype).view(B, W, -1) / (1.0 - p_drop)
    f = f @ p[pre + "ffnetwork.lin2.weight"].t() + p[pre + "ffnetwork.lin2.bias"]
    return F.layer_norm(y1 + f, (D,), p[pre + "ln_ffnetwork.weight"], p[pre + "ln_ffnetwork.bias"], 1e-5)


def make_dropout_masks(d: Dims, seed: int, ar: str = "GRU", heads: str = "transformer", p_drop: float = 0.1):
    """Seeded keep-masks (uint8, 1 = kept) for the train-mode dropouts of the transformer layers on the path, in the order
    the reference draws them: the context network (one layer over S frames: attention (B*nh, S, S), FFN (B, S, dff)) and
    then the K prediction heads (attention (B*nh, W, W), FFN (B, W, dff) each).  Returns (ar_masks | None, head_masks | None)
    with head_masks = (att (K, B*nh, W, W), ffn (K, B, W, dff))."""
    g = torch.Generator().manual_seed(seed + 15485863)

    def keep(*shape):
        return (torch.rand(*shape, generator=g) >= p_drop).to(torch.uint8)

    ar_masks = (keep(d.B * T_HEADS, d.S, d.S), keep(d.B, d.S, T_DFF)) if ar == "transformer" else None
    head_masks = None
    if heads == "transformer":
        head_masks = (torch.stack([keep(d.B * T_HEADS, d.W, d.W) for _ in range(d.K)]),
                      torch.stack([keep(d.B, d.W, T_DFF) for _ in range(d.K)]))
    return ar_masks, head_masks


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


def lstm_forward(z, p, n_layers=1, h0=None, c0=None, prefix="gAR.baseNet."):
    """torch.nn.LSTM(batch_first=True) restated (reference call site cpc/model.py:171-173, 193).

    i = sigma(W_ii x + b_ii + W_hi h + b_hi); f = sigma(W_if x + b_if + W_hf h + b_hf)
    g = tanh(W_ig x + b_ig + W_hg h + b_hg);  o = sigma(W_io x + b_io + W_ho h + b_ho)
    c' = f * c + i * g ;  h' = o * tanh(c')                                    (gate packing i, f, g, o)
    Returns (out (B,S,Har), (hT, cT) each (nLayers,B,Har)).
    """
    B, S, _ = z.shape
    inp = z
    h_last, c_last = [], []
    for l in range(n_layers):
        w_ih, w_hh = p[f"{prefix}weight_ih_l{l}"], p[f"{prefix}weight_hh_l{l}"]
        b_ih, b_hh = p[f"{prefix}bias_ih_l{l}"], p[f"{prefix}bias_hh_l{l}"]
        Har = w_hh.shape[1]
        h = torch.zeros(B, Har, dtype=z.dtype) if h0 is None else h0[l]
        c = torch.zeros(B, Har, dtype=z.dtype) if c0 is None else c0[l]
        gi_all = inp @ w_ih.t() + b_ih
        outs = []
        for t in range(S):
            a = gi_all[:, t] + h @ w_hh.t() + b_hh
            i = torch.sigmoid(a[:, :Har])
            f = torch.sigmoid(a[:, Har:2 * Har])
            g = torch.tanh(a[:, 2 * Har:3 * Har])
            o = torch.sigmoid(a[:, 3 * Har:])
            c = f * c + i * g
            h = o * torch.tanh(c)
            outs.append(h)
        inp = torch.stack(outs, dim=1)
        h_last.append(h)
        c_last.append(c)
    return inp, (torch.stack(h_last, dim=0), torch.stack(c_last, dim=0))


def ar_kind(p):
    """Which context network a model parameter dict describes (model.py:171-179 / feature_loader.py:138-152)."""
    if "gAR.0.multihead.Wq.weight" in p:
        return "transformer"
    return "LSTM" if p["gAR.baseNet.weight_hh_l0"].shape[0] == 4 * p["gAR.baseNet.weight_hh_l0"].shape[1] else "GRU"


def ar_forward(z, p, n_layers=1, hidden=None, ar_masks=None):
    """The context network over z (B,S,H): returns (c, new hidden).  `hidden` as the reference keeps it (model.py:193-198):
    a tensor for the GRU, a tuple (h, c) for the LSTM; the transformer is stateless."""
    kind = ar_kind(p)
    if kind == "transformer":
        att, ffn = ar_masks if ar_masks is not None else (None, None)
        return transformer_head_forward(z, p, "gAR.0.", att, ffn), None
    if kind == "LSTM":
        h0, c0 = hidden if hidden is not None else (None, None)
        return lstm_forward(z, p, n_layers, h0, c0)
    return gru_forward(z, p, n_layers, hidden)


def model_forward(x, p, n_layers=1, h0=None, ar_masks=None):
    """cpc/model.py:286-289 -> (cFeature (B,S,Har), encodedData (B,S,H))."""
    z = encoder_forward(x, p).permute(0, 2, 1)
    c, _ = ar_forward(z, p, n_layers, h0, ar_masks)
    return c, z


def feature_forward(seq, p, n_layers=1, max_size_seq=64000, keep_hidden=True, get_encoded=False):
    """cpc/feature_loader.py:228-269 (buildFeature, strict=False, seqNorm=False) over FeatureModule (:15-38): the sequence
    (n_samples,) is cut into chunks of `max_size_seq` samples - the last one has whatever length is left - and each chunk
    goes through CPCModel.forward; with keepHidden (model.py:194-198) the recurrent state carries over from chunk to
    chunk.  Returns (1, n_frames, dim)."""
    out, hidden, start = [], None, 0
    n = seq.numel()
    while start < n:
        sub = seq[start:min(n, start + max_size_seq)].view(1, 1, -1)
        z = encoder_forward(sub, p).permute(0, 2, 1)
        c, h = ar_forward(z, p, n_layers, hidden if keep_hidden else None)
        if keep_hidden and h is not None:
            hidden = tuple(t.detach() for t in h) if isinstance(h, tuple) else h.detach()
        out.append(z if get_encoded else c)
        start += max_size_seq
    return torch.cat(out, dim=1)


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


def criterion_forward(c, z, crit_p, batch_idx, seq_idx, K, N, materialize=True, heads="linear", head_masks=None):
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
        else:                                                                   # criterion.py:82-88
            att, ffn = (head_masks[0][k - 1], head_masks[1][k - 1]) if head_masks is not None else (None, None)
            pred = transformer_head_forward(cw, crit_p, f"wPrediction.predictors.{k - 1}.0.", att, ffn)
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
