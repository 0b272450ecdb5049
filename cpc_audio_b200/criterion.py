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
from .model import _bytes, _dtype_code, _require_cuda, default_dtype


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


class PredictionNetwork(nn.Module):
    """cpc/criterion/criterion.py:44-95, linear heads.  Parameters are K views into one (K, H, Har) buffer so
    that the K projections run as a single GEMM; ``stacked()`` re-packs them if .to()/.cuda() split them."""

    def __init__(self, nPredicts, dimOutputAR, dimOutputEncoder, rnnMode=None, dropout=False, sizeInputSeq=116):
        super().__init__()
        if rnnMode in ("RNN", "LSTM", "ffd", "conv4", "conv8", "conv12", "transformer"):
            raise NotImplementedError(f"cpc_audio_b200: rnnMode={rnnMode!r} prediction heads are outside the accelerated "
                                      f"hot path of this build (linear heads only: pass --rnnMode linear)")
        if dropout:
            raise NotImplementedError("cpc_audio_b200: criterion dropout is outside the accelerated hot path")
        self.RESIDUAL_STD = 0.01
        self.dimOutputAR = dimOutputAR
        self.dropout = None
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
        c = c.contiguous().float()
        z = z.contiguous().float()
        d = L.make_dims(B, S * 160, H, Har, K, N, 1, dtype_code)
        losses = torch.empty(K, device=dev, dtype=torch.float32)
        acc = torch.empty(K, device=dev, dtype=torch.float32)
        save = _bytes(lib.cpcb200_criterion_save_bytes(d), dev)
        wsn = lib.cpcb200_criterion_ws_bytes(d, 0)
        ws = _bytes(wsn, dev)
        with torch.cuda.device(dev):
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
        d = L.make_dims(B, S * 160, H, Har, K, N, 1, dtype_code)
        dc = torch.empty_like(c)
        dz = torch.empty_like(z)
        dw = torch.empty(K, H, Har, device=dev, dtype=torch.float32)
        wsn = lib.cpcb200_criterion_ws_bytes(d, 1)
        ws = _bytes(wsn, dev)
        dlosses = dlosses.contiguous().float()
        with torch.cuda.device(dev):
            L.check(lib.cpcb200_criterion_bwd(d, L.ptr(c), L.ptr(z), L.ptr(w_flat), L.ptr(ext), L.ptr(dlosses), L.ptr(save),
                                              L.ptr(dc), L.ptr(dz), L.ptr(dw), L.ptr(ws), wsn, L.stream_ptr(dev)),
                    "criterion_bwd")
        from .optim import sinks_for
        sinks = sinks_for(ctx.weights)
        if sinks is not None and all(s.is_contiguous() and s.data_ptr() == sinks[0].data_ptr() + i * H * Har * 4
                                     for i, s in enumerate(sinks)):
            sinks[0].as_strided((K, H, Har), (H * Har, Har, 1)).add_(dw)  # one kernel into the bucket
            return (dc, dz, None, None, None, *([None] * K))
        return (dc, dz, None, None, None, *dw.unbind(0))


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

    def extIndices(self, batchIdx, seqIdx, dims):
        """criterion.py:191-199 on the device: ext = ((seqIdx + w) mod S) + batchIdx * S, int32 (B, N, W)."""
        B, S, H, Har, K, N, dtype_code = dims
        lib = L.lib()
        dev = batchIdx.device
        ext = torch.empty(B, N, S - K, device=dev, dtype=torch.int32)
        d = L.make_dims(B, S * 160, H, Har, K, N, 1, dtype_code)
        with torch.cuda.device(dev):
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
        batchIdx, seqIdx = self.sampleIndices(batchSize, windowSize, seqSize, encodedData.device)
        ext = self.extIndices(batchIdx, seqIdx, dims)
        w_flat = self.wPrediction.stacked()
        weights = [p.weight for p in self.wPrediction.predictors]
        losses, acc = _CriterionFn.apply(cFeature, encodedData, ext, dims, w_flat, *weights)
        return losses.view(1, -1), acc.view(1, -1)


# correctly spelled alias (BASELINE.json uses it)
CPCUnsupervisedCriterion = CPCUnsupersivedCriterion
