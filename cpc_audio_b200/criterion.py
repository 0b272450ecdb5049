"""Drop-in mirror of cpc/criterion/criterion.py:{BaseCriterion, PredictionNetwork, CPCUnsupersivedCriterion}.

Same constructor arguments, state_dict keys (``wPrediction.predictors.{k}.weight``) and outputs
(losses (1,K), acc (1,K)) as the reference.  The two ``torch.randint`` draws of criterion.py:181-189 stay in
PyTorch, issued in the same order / shape / dtype / device, so that negative-sample indices are bit-exact with
the reference under the same generator state; everything after that runs in libcpc_b200.so.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib as L
from .model import _bytes, _dtype_code, _plan, _require_cuda, default_dtype  # noqa: I001


class BaseCriterion(nn.Module):
    """cpc/criterion/criterion.py:121-127."""

    def warmUp(self):
        return False

    def update(self):
        return


class _PredictorWeight(nn.Module):
    """Stands for nn.Linear(dimOutputAR, dimOutputEncoder, bias=False): holds ``weight`` (H, Har) only."""

    def __init__(self, weight):
        super().__init__()
        self.weight = nn.Parameter(weight)
        self.in_features = weight.shape[1]
        self.out_features = weight.shape[0]


T_HEADS, T_DFF, T_DROPOUT = 8, 2048, 0.1  # cpc/transformers.py:98-100 defaults used by criterion.py:84-88


def draw_dropout_masks(n_layers, batch, seq, nheads, dff, p, device):
    """Keep-masks of the train-mode dropouts of `n_layers` TransformerLayers applied one after the other
    (transformers.py:49 on the attention probabilities (B*nheads, W, W), then :92 on the FFN hidden (B, W, dff)), drawn
    with the SAME torch call the reference's nn.Dropout makes, on same-shaped fp32 tensors, in the same order: the fused
    dropout kernel's mask depends on the generator state and the tensor's size only, so a reference run with the same
    ``torch.cuda.manual_seed`` sees identical masks.  Returns uint8 tensors (n_layers, B*nheads, W, W), (n_layers, B*W, dff)."""
    att = torch.empty(n_layers, batch * nheads, seq, seq, dtype=torch.uint8, device=device)
    ffn = torch.empty(n_layers, batch * seq, dff, dtype=torch.uint8, device=device)
    ones_a = torch.ones(batch * nheads, seq, seq, device=device)
    ones_f = torch.ones(batch, seq, dff, device=device)
    ffn_v = ffn.view(n_layers, batch, seq, dff)
    for k in range(n_layers):  # (the comparison writes straight into the stacked uint8 buffers: one kernel per draw, no copy)
        torch.ne(torch.nn.functional.dropout(ones_a, p, True), 0, out=att[k])
        torch.ne(torch.nn.functional.dropout(ones_f, p, True), 0, out=ffn_v[k])
    return att, ffn


def _thead_struct(tensors, masks):
    tp = L.THeadParams(*[t.data_ptr() for t in tensors], T_DFF, T_HEADS, None, None, 1.0)
    if masks is not None:
        tp.att_keep, tp.ffn_keep, tp.keep_scale = masks[0].data_ptr(), masks[1].data_ptr(), float(masks[2])
    return tp


class _Attention(nn.Module):
    """Parameter / buffer holder of transformers.py:10-32 (relpos=True): Krelpos (dk, S), buffers z and mask."""

    def __init__(self, sizeSeq, dk):
        super().__init__()
        import math
        self.sizeSeq = sizeSeq
        self.Krelpos = nn.Parameter(torch.empty(dk, sizeSeq).uniform_(-1.0 / math.sqrt(dk), 1.0 / math.sqrt(dk)))
        self.register_buffer("z", torch.zeros(1, sizeSeq, 1))
        mask = 1 - torch.tril(torch.ones(sizeSeq, sizeSeq), diagonal=0)
        mask[mask == 1] = -float("inf")
        self.register_buffer("mask", mask.unsqueeze(0))


class _MultiHead(nn.Module):
    def __init__(self, sizeSeq, dmodel, nheads):
        super().__init__()
        self.Wo = nn.Linear(dmodel, dmodel, bias=False)
        self.Wk = nn.Linear(dmodel, dmodel, bias=False)
        self.Wq = nn.Linear(dmodel, dmodel, bias=False)
        self.Wv = nn.Linear(dmodel, dmodel, bias=False)
        self.nheads, self.dk = nheads, dmodel // nheads
        self.Att = _Attention(sizeSeq, self.dk)


class _FFN(nn.Module):
    def __init__(self, dmodel, dff):
        super().__init__()
        self.lin1 = nn.Linear(dmodel, dff, bias=True)
        self.lin2 = nn.Linear(dff, dmodel, bias=True)


class _TransformerLayer(nn.Module):
    """Parameter holder with the module tree (hence state_dict keys) of transformers.py:98-106."""

    def __init__(self, sizeSeq, dmodel, dff=T_DFF, nheads=T_HEADS, dropout=T_DROPOUT, compute_dtype=None):
        super().__init__()
        if dff != T_DFF or nheads != T_HEADS:
            raise NotImplementedError("cpc_audio_b200: TransformerLayer supports the reference defaults dff=2048, nheads=8")
        self.dropout = float(dropout)
        self.compute_dtype = compute_dtype or default_dtype()
        self.multihead = _MultiHead(sizeSeq, dmodel, nheads)
        self.ln_multihead = nn.LayerNorm(dmodel)
        self.ffnetwork = _FFN(dmodel, dff)
        self.ln_ffnetwork = nn.LayerNorm(dmodel)

    def thead_tensors(self):
        m = self.multihead
        return [m.Wq.weight, m.Wk.weight, m.Wv.weight, m.Wo.weight, m.Att.Krelpos, self.ln_multihead.weight,
                self.ln_multihead.bias, self.ffnetwork.lin1.weight, self.ffnetwork.lin1.bias, self.ffnetwork.lin2.weight,
                self.ffnetwork.lin2.bias, self.ln_ffnetwork.weight, self.ln_ffnetwork.bias]

    def forward(self, x):
        """transformers.py:109-111 over a whole window: x (B, S, D) -> (B, S, D).  This is how the layer runs as the
        CONTEXT network (--arMode transformer, feature_loader.py:138-142); as a prediction head the criterion batches the K
        layers itself.  train() applies the reference's dropout (masks drawn with torch's generator)."""
        _require_cuda(x, "TransformerLayer")
        B, S, D = x.shape
        if S != self.multihead.Att.sizeSeq:
            raise ValueError(f"TransformerLayer was built for sequences of {self.multihead.Att.sizeSeq} frames, got {S}")
        masks = None
        if self.training and self.dropout > 0:
            att, ffn = self.dropoutMasks(B, S, x.device)
            masks = (att, ffn, 1.0 / (1.0 - self.dropout))
        return _TLayerFn.apply(x, _dtype_code(self.compute_dtype), masks, *self.thead_tensors())

    def dropoutMasks(self, batchSize, seqSize, device):
        """Train-mode keep-masks of this layer: attention (1, B*nheads, S, S), FFN (1, B*S, dff), uint8."""
        return draw_dropout_masks(1, batchSize, seqSize, T_HEADS, T_DFF, self.dropout, device)


class _TLayerFn(torch.autograd.Function):
    """One TransformerLayer over (B, S, D): cpcb200_tlayer_fwd / _bwd."""

    @staticmethod
    def forward(ctx, x, dtype_code, masks, *params):
        lib = L.lib()
        B, S, D = x.shape
        dev = x.device
        x = L.f32c(x)
        params = [p.detach().contiguous() for p in params]
        tp = _thead_struct(params, masks)
        d = L.make_dims(B, S * 160, D, D, 1, 1, 1, dtype_code)
        y = torch.empty(B, S, D, device=dev, dtype=torch.float32)
        save = _bytes(lib.cpcb200_tlayer_save_bytes(d, T_DFF, T_HEADS), dev)
        wsn = lib.cpcb200_tlayer_ws_bytes(d, T_DFF, T_HEADS, 0)
        ws = _bytes(wsn, dev)
        with L.device_guard(dev):
            L.check(lib.cpcb200_tlayer_fwd(d, L.ptr(x), tp, L.ptr(y), L.ptr(save), L.ptr(ws), wsn, L.stream_ptr(dev)), "tlayer_fwd")
        ctx.save_for_backward(x, save, *params)
        ctx.masks = masks
        ctx.dims = (B, S, D, dtype_code)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = L.lib()
        x, save, *params = ctx.saved_tensors
        B, S, D, dtype_code = ctx.dims
        dev = x.device
        d = L.make_dims(B, S * 160, D, D, 1, 1, 1, dtype_code)
        tp = _thead_struct(params, ctx.masks)
        grads = [torch.zeros_like(t) for t in params]
        tg = _thead_struct(grads, None)
        dx = torch.empty_like(x)
        wsn = lib.cpcb200_tlayer_ws_bytes(d, T_DFF, T_HEADS, 1)
        ws = _bytes(wsn, dev)
        dy = dy.contiguous().float()
        with L.device_guard(dev):
            L.check(lib.cpcb200_tlayer_bwd(d, L.ptr(x), tp, L.ptr(dy), L.ptr(save), L.ptr(dx), tg, L.ptr(ws), wsn,
                                           L.stream_ptr(dev)), "tlayer_bwd")
        return (dx, None, None, *grads)


class PredictionNetwork(nn.Module):
    """cpc/criterion/criterion.py:44-95: K prediction heads, ``rnnMode`` 'linear' (nn.Linear, criterion.py:89-95) or
    'transformer' (one-layer transformers, criterion.py:82-88).  The linear weights are K views into one (K, H, Har)
    buffer so that the K projections run as a single GEMM; ``stacked()`` re-packs them if .to()/.cuda() split them."""

    def __init__(self, nPredicts, dimOutputAR, dimOutputEncoder, rnnMode=None, dropout=False, sizeInputSeq=116):
        super().__init__()
        if rnnMode in ("RNN", "LSTM", "ffd", "conv4", "conv8", "conv12"):
            raise NotImplementedError(f"cpc_audio_b200: rnnMode={rnnMode!r} prediction heads are outside the accelerated "
                                      f"hot path of this build (use --rnnMode linear or --rnnMode transformer; both train "
                                      f"and evaluate through the unmodified cpc/train.py)")
        if dropout:
            raise NotImplementedError("cpc_audio_b200: criterion dropout is outside the accelerated hot path")
        self.RESIDUAL_STD = 0.01
        self.dimOutputAR = dimOutputAR
        self.dropout = None
        self.transformer = rnnMode == "transformer"
        if self.transformer:
            if dimOutputAR != dimOutputEncoder or dimOutputEncoder % T_HEADS != 0:
                raise NotImplementedError("cpc_audio_b200: transformer heads need hiddenGar == hiddenEncoder (criterion.py:85)")
            if sizeInputSeq > 128:
                raise NotImplementedError("cpc_audio_b200: transformer heads support at most 128 anchor positions")
            self.sizeInputSeq = sizeInputSeq
            # criterion.py:84-88: buildTransformerAR(dimOutputEncoder, 1, sizeInputSeq, False) -> nn.Sequential(TransformerLayer)
            self.predictors = nn.ModuleList([nn.Sequential(_TransformerLayer(sizeInputSeq, dimOutputEncoder))
                                             for _ in range(nPredicts)])
            return
        flat = torch.empty(nPredicts, dimOutputEncoder, dimOutputAR)
        for i in range(nPredicts):
            # nn.Linear default init first (consumes the generator like the reference constructor does) ...
            flat[i].copy_(nn.Linear(dimOutputAR, dimOutputEncoder, bias=False).weight.data)
            if dimOutputEncoder > dimOutputAR:  # ... then criterion.py:92-95
                residual = dimOutputEncoder - dimOutputAR
                flat[i].copy_(torch.cat([torch.randn(dimOutputAR, dimOutputAR),
                                         self.RESIDUAL_STD * torch.randn(residual, dimOutputAR)], dim=0))
        self.predictors = nn.ModuleList([_PredictorWeight(flat[i]) for i in range(nPredicts)])

    def stacked(self):
        """Return the (K, H, Har) tensor the K weights live in, re-packing them if they are not contiguous."""
        f = getattr(self, "_flat", None)
        if f is not None:  # fast path (every step): first and last weight still sit where the packed buffer says
            w0, wl = self.predictors[0].weight, self.predictors[-1].weight
            if w0.data_ptr() == f.data_ptr() and wl.data_ptr() == f.data_ptr() + (f.shape[0] - 1) * f.stride(0) * 4 \
                    and w0.device == f.device:
                return f
        ws = [p.weight for p in self.predictors]
        K, (H, Har) = len(ws), ws[0].shape
        step = H * Har * ws[0].element_size()
        base = ws[0].data_ptr()
        sto = ws[0].untyped_storage()
        ok = all(w.is_contiguous() and w.data_ptr() == base + i * step and w.device == ws[0].device
                 and w.untyped_storage().data_ptr() == sto.data_ptr() for i, w in enumerate(ws))
        ok = ok and (base - sto.data_ptr()) + K * step <= sto.nbytes()
        if ok and getattr(self, "_flat", None) is not None and self._flat.data_ptr() == base:
            return self._flat
        if ok and all(isinstance(w, nn.Parameter) and w.is_leaf for w in ws):
            # already packed inside a larger buffer (e.g. FlatAdam's flat parameters): view it, do not copy
            self._flat = ws[0].detach().as_strided((K, H, Har), (H * Har, Har, 1))
            return self._flat
        flat = torch.stack([w.detach() for w in ws]).contiguous()
        if all(isinstance(w, nn.Parameter) and w.is_leaf for w in ws):
            for i, w in enumerate(ws):  # re-point the parameters at the packed buffer (like GRU.flatten_parameters)
                w.data = flat[i]
            self._flat = flat
        return flat  # DataParallel replicas hold non-leaf copies: use a packed copy, leave them alone


class _CriterionFn(torch.autograd.Function):
    """(c, z, ext, K weights) -> (losses (K), acc (K)).  criterion.py:97-118, 207-217, 245-257."""

    @staticmethod
    def forward(ctx, c, z, ext, dims, w_flat, *weights):
        lib = L.lib()
        B, S, H, Har, K, N, dtype_code = dims
        dev = c.device
        c = L.f32c(c)
        z = L.f32c(z)
        d, save_n, wsns = _plan("crit", B, S * 160, H, Har, K, N, 1, dtype_code)
        losses = torch.empty(K, device=dev, dtype=torch.float32)
        acc = torch.empty(K, device=dev, dtype=torch.float32)
        save = _bytes(save_n, dev)
        wsn = wsns[0]
        ws = _bytes(wsn, dev)
        with L.device_guard(dev):
            L.check(lib.cpcb200_criterion_fwd(d, L.ptr(c), L.ptr(z), L.ptr(w_flat), L.ptr(ext), L.ptr(losses), L.ptr(acc),
                                              L.ptr(save), L.ptr(ws), wsn, L.stream_ptr(dev)), "criterion_fwd")
        ctx.save_for_backward(c, z, ext, save, w_flat)
        ctx.weights = weights
        ctx.dims = dims
        ctx.mark_non_differentiable(acc)
        return losses, acc

    @staticmethod
    def backward(ctx, dlosses, _dacc):
        lib = L.lib()
        c, z, ext, save, w_flat = ctx.saved_tensors
        B, S, H, Har, K, N, dtype_code = ctx.dims
        dev = c.device
        d, _, wsns = _plan("crit", B, S * 160, H, Har, K, N, 1, dtype_code)
        dc = torch.empty_like(c)
        dz = torch.empty_like(z)
        # the weight gradient is accumulated (+=): straight into the gradient bucket when the K weights have their sinks
        # there as one contiguous (K, H, Har) block, else into a fresh zero buffer handed back to autograd
        from .optim import sinks_for
        sinks = sinks_for(ctx.weights)
        sunk = sinks is not None and all(s.is_contiguous() and s.data_ptr() == sinks[0].data_ptr() + i * H * Har * 4
                                         for i, s in enumerate(sinks))
        dw = sinks[0].as_strided((K, H, Har), (H * Har, Har, 1)) if sunk else torch.zeros(K, H, Har, device=dev, dtype=torch.float32)
        wsn = wsns[1]
        ws = _bytes(wsn, dev)
        dlosses = L.f32c(dlosses)
        with L.device_guard(dev):
            L.check(lib.cpcb200_criterion_bwd(d, L.ptr(c), L.ptr(z), L.ptr(w_flat), L.ptr(ext), L.ptr(dlosses), L.ptr(save),
                                              L.ptr(dc), L.ptr(dz), L.ptr(dw), L.ptr(ws), wsn, L.stream_ptr(dev)),
                    "criterion_bwd")
        if sunk:
            return (dc, dz, None, None, None, *([None] * K))
        return (dc, dz, None, None, None, *dw.unbind(0))


class _CriterionTFn(torch.autograd.Function):
    """Transformer heads + scoring + InfoNCE.  params = 13 tensors per head, head-major (see _TransformerLayer)."""

    @staticmethod
    def forward(ctx, c, z, ext, dims, masks, *params):
        lib = L.lib()
        B, S, H, Har, K, N, dtype_code = dims
        dev = c.device
        c = L.f32c(c)
        z = L.f32c(z)
        nf = len(L.THEAD_FIELDS)
        stacked = [torch.stack([params[k * nf + j].detach() for k in range(K)]).contiguous() for j in range(nf)]
        tp = _thead_struct(stacked, masks)
        d = L.make_dims(B, S * 160, H, Har, K, N, 1, dtype_code)
        losses = torch.empty(K, device=dev, dtype=torch.float32)
        acc = torch.empty(K, device=dev, dtype=torch.float32)
        save = _bytes(lib.cpcb200_criterion_t_save_bytes(d, T_DFF, T_HEADS), dev)
        wsn = lib.cpcb200_criterion_t_ws_bytes(d, T_DFF, T_HEADS, 0)
        ws = _bytes(wsn, dev)
        with L.device_guard(dev):
            L.check(lib.cpcb200_criterion_t_fwd(d, L.ptr(c), L.ptr(z), tp, L.ptr(ext), L.ptr(losses), L.ptr(acc), L.ptr(save),
                                                L.ptr(ws), wsn, L.stream_ptr(dev)), "criterion_t_fwd")
        ctx.save_for_backward(c, z, ext, save, *stacked)
        ctx.dims = dims
        ctx.masks = masks
        ctx.mark_non_differentiable(acc)
        return losses, acc

    @staticmethod
    def backward(ctx, dlosses, _dacc):
        lib = L.lib()
        c, z, ext, save, *stacked = ctx.saved_tensors
        B, S, H, Har, K, N, dtype_code = ctx.dims
        dev = c.device
        d = L.make_dims(B, S * 160, H, Har, K, N, 1, dtype_code)
        tp = _thead_struct(stacked, ctx.masks)
        grads = [torch.zeros_like(t) for t in stacked]
        tg = _thead_struct(grads, None)
        dc = torch.empty_like(c)
        dz = torch.empty_like(z)
        wsn = lib.cpcb200_criterion_t_ws_bytes(d, T_DFF, T_HEADS, 1)
        ws = _bytes(wsn, dev)
        dlosses = dlosses.contiguous().float()
        with L.device_guard(dev):
            L.check(lib.cpcb200_criterion_t_bwd(d, L.ptr(c), L.ptr(z), tp, L.ptr(ext), L.ptr(dlosses), L.ptr(save), L.ptr(dc),
                                                L.ptr(dz), tg, L.ptr(ws), wsn, L.stream_ptr(dev)), "criterion_t_bwd")
        nf = len(L.THEAD_FIELDS)
        flat = [grads[j][k] for k in range(K) for j in range(nf)]
        return (dc, dz, None, None, None, *flat)


class CPCUnsupersivedCriterion(BaseCriterion):
    """cpc/criterion/criterion.py:139-257 (class name spelled as in the reference)."""

    def __init__(self, nPredicts, dimOutputAR, dimOutputEncoder, negativeSamplingExt, mode=None, rnnMode=False,
                 dropout=False, speakerEmbedding=0, nSpeakers=0, sizeInputSeq=128, compute_dtype=None):
        super().__init__()
        if speakerEmbedding > 0:
            raise NotImplementedError("cpc_audio_b200: speakerEmbedding > 0 is outside the accelerated hot path")
        if mode not in [None, "reverse"]:
            raise ValueError("Invalid mode")
        if mode == "reverse":
            raise NotImplementedError("cpc_audio_b200: cpc_mode='reverse' is outside the accelerated hot path")
        if nPredicts > 16:
            raise NotImplementedError("cpc_audio_b200: nPredicts must be <= 16")
        self.speakerEmb = None
        self.wPrediction = PredictionNetwork(nPredicts, dimOutputAR, dimOutputEncoder, rnnMode=rnnMode, dropout=dropout,
                                             sizeInputSeq=sizeInputSeq - nPredicts)
        self.nPredicts = nPredicts
        self.negativeSamplingExt = negativeSamplingExt
        self.mode = mode
        self.compute_dtype = compute_dtype or default_dtype()

    def sampleIndices(self, batchSize, windowSize, seqSize, device):
        """The two draws of criterion.py:181-189 (batchIdx first, then seqIdx), raw, int64, flat (B, N, W)."""
        n = self.negativeSamplingExt * windowSize * batchSize
        batchIdx = torch.randint(low=0, high=batchSize, size=(n,), device=device)
        seqIdx = torch.randint(low=1, high=seqSize, size=(n,), device=device)
        return batchIdx, seqIdx

    def dropoutMasks(self, batchSize, windowSize, device, p):
        """Train-mode dropout masks of the K transformer heads, head by head (att_k, ffn_k), as PredictionNetwork.forward
        (criterion.py:106-108) would draw them."""
        return draw_dropout_masks(self.nPredicts, batchSize, windowSize, T_HEADS, T_DFF, p, device)

    def extIndices(self, batchIdx, seqIdx, dims):
        """criterion.py:191-199 on the device: ext = ((seqIdx + w) mod S) + batchIdx * S, int32 (B, N, W)."""
        B, S, H, Har, K, N, dtype_code = dims
        lib = L.lib()
        dev = batchIdx.device
        ext = torch.empty(B, N, S - K, device=dev, dtype=torch.int32)
        d = _plan("crit", B, S * 160, H, Har, K, N, 1, dtype_code)[0]
        with L.device_guard(dev):
            L.check(lib.cpcb200_sample_ext_idx(d, L.ptr(batchIdx.contiguous()), L.ptr(seqIdx.contiguous()), L.ptr(ext),
                                               L.stream_ptr(dev)), "sample_ext_idx")
        return ext

    def forward(self, cFeature, encodedData, label):
        _require_cuda(cFeature, "CPCUnsupersivedCriterion")
        batchSize, seqSize, dimAR = cFeature.size()
        windowSize = seqSize - self.nPredicts
        if windowSize < 1:
            raise ValueError(f"sequence of {seqSize} frames is too short for nPredicts={self.nPredicts}")
        H = encodedData.size(2)
        dims = (batchSize, seqSize, H, dimAR, self.nPredicts, self.negativeSamplingExt, _dtype_code(self.compute_dtype))
        dev = encodedData.device
        ready = getattr(encodedData, "_cpcb200_ready", None)
        if ready is not None:
            # criterion.py:181-199 only needs the SHAPE of z: drawn on the side stream, beside the recurrence that is still
            # producing cFeature on this stream (the generator is consumed in host order: same draws either way)
            from .model import side_stream
            main, side = torch.cuda.current_stream(dev), side_stream(dev)
            side.wait_event(ready)
            with torch.cuda.stream(side):
                batchIdx, seqIdx = self.sampleIndices(batchSize, windowSize, seqSize, dev)
                ext = self.extIndices(batchIdx, seqIdx, dims)
                done = torch.cuda.Event()
                done.record(side)
            main.wait_event(done)
            for t in (batchIdx, seqIdx, ext):
                t.record_stream(main)
        else:
            batchIdx, seqIdx = self.sampleIndices(batchSize, windowSize, seqSize, dev)
            ext = self.extIndices(batchIdx, seqIdx, dims)
        if self.wPrediction.transformer:
            if windowSize != self.wPrediction.sizeInputSeq:
                raise ValueError(f"transformer heads were built for {self.wPrediction.sizeInputSeq} positions, got {windowSize}")
            masks = None
            p_drop = self.wPrediction.predictors[0][0].dropout
            if self.training and p_drop > 0:  # transformers.py:49,92: drawn AFTER the negatives (criterion.py:234 precedes :241)
                att, ffn = self.dropoutMasks(batchSize, windowSize, cFeature.device, p_drop)
                masks = (att, ffn, 1.0 / (1.0 - p_drop))
            params = [t for p in self.wPrediction.predictors for t in p[0].thead_tensors()]
            losses, acc = _CriterionTFn.apply(cFeature, encodedData, ext, dims, masks, *params)
            return losses.view(1, -1), acc.view(1, -1)
        w_flat = self.wPrediction.stacked()
        weights = [p.weight for p in self.wPrediction.predictors]
        losses, acc = _CriterionFn.apply(cFeature, encodedData, ext, dims, w_flat, *weights)
        return losses.view(1, -1), acc.view(1, -1)


# correctly spelled alias (BASELINE.json uses it)
CPCUnsupervisedCriterion = CPCUnsupersivedCriterion
