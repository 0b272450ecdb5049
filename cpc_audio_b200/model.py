"""Drop-in mirrors of the reference module surfaces cpc/model.py:{ChannelNorm,CPCEncoder,CPCAR,CPCModel}.

Same constructor arguments, attributes, state_dict keys and return shapes as the reference, so
``cpc/train.py`` and ``cpc/feature_loader.py`` run unchanged once ``cpc.model`` is patched
(``cpc_audio_b200.patch.install``).  All arithmetic runs in libcpc_b200.so (hand-written sm_100a CUDA) through
the C ABI of ``include/cpc_b200.h``; there is no eager/CPU fallback - CPU tensors raise.
"""
from __future__ import annotations

import functools
import os

import torch
import torch.nn as nn

from . import _lib as L

_DTYPES = {"bf16": L.BF16, "bfloat16": L.BF16, "f32": L.F32, "fp32": L.F32, "float32": L.F32}


def default_dtype() -> str:
    return os.environ.get("CPC_B200_DTYPE", "bf16")


def _dtype_code(name) -> int:
    if name not in _DTYPES:
        raise ValueError(f"compute dtype must be one of {sorted(_DTYPES)}, got {name!r}")
    return _DTYPES[name]


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"cpc_audio_b200.{what}: input is on {t.device}; this implementation is CUDA (sm_100a) only "
                           f"and has no CPU fallback")


def _grad_targets(params, dev):
    """Where a backward pass accumulates parameter gradients: the views of a live GradBucket (then autograd gets
    None for them) or one freshly zeroed flat buffer returned to autograd."""
    from .optim import sinks_for
    sinks = sinks_for(params)
    if sinks is not None:
        return sinks, True
    sizes = [p.numel() for p in params]
    flat = torch.zeros(sum(sizes), device=dev, dtype=torch.float32)
    return [g.view_as(p) for g, p in zip(flat.split(sizes), params)], False


def _bytes(n, device):
    return torch.empty(max(int(n), 256), dtype=torch.uint8, device=device)


@functools.lru_cache(maxsize=512)
def _plan(op, B, Lw, H, Har, K, N, nL, code, a=0, b=0):
    """(dims struct, save bytes, workspace bytes for mode 0 / 1 / 2) of one op at one shape: the size queries cross the
    C ABI once per shape, not once per step (the eager train.py-driven loop is host-bound)."""
    lib = L.lib()
    d = L.make_dims(B, Lw, H, Har, K, N, nL, code)
    if op == "enc":
        return d, lib.cpcb200_encoder_save_bytes(d), tuple(lib.cpcb200_encoder_ws_bytes(d, m) for m in (0, 1, 2))
    if op == "gru":
        return d, lib.cpcb200_gru_save_bytes(d), tuple(lib.cpcb200_gru_ws_bytes(d, m) for m in (0, 1))
    if op == "lstm":
        return d, lib.cpcb200_lstm_save_bytes(d), tuple(lib.cpcb200_lstm_ws_bytes(d, m) for m in (0, 1, 2))
    if op == "crit":
        return d, lib.cpcb200_criterion_save_bytes(d), tuple(lib.cpcb200_criterion_ws_bytes(d, m) for m in (0, 1))
    if op == "crit_t":
        return d, lib.cpcb200_criterion_t_save_bytes(d, a, b), tuple(lib.cpcb200_criterion_t_ws_bytes(d, a, b, m) for m in (0, 1))
    if op == "tlayer":
        return d, lib.cpcb200_tlayer_save_bytes(d, a, b), tuple(lib.cpcb200_tlayer_ws_bytes(d, a, b, m) for m in (0, 1))
    raise KeyError(op)


_CONV_GEOMETRY = ((10, 5, 3), (8, 4, 2), (4, 2, 1), (4, 2, 1), (4, 2, 1))  # (kernel, stride, padding), cpc/model.py:83-92


def frames_for(n_samples: int) -> int:
    """Output frames of the encoder for a window of `n_samples` (= n_samples // 160 when that divides; the reference's
    Conv1d stack accepts any length, cpc/feature_loader.py:228-269 feeds it chunks whose last one is arbitrary)."""
    n = int(n_samples)
    for k, s, p in _CONV_GEOMETRY:
        n = (n + 2 * p - k) // s + 1
        if n < 1:
            return 0
    return n


_SIDE_STREAMS = {}


def side_stream(device):
    """One library-owned side stream per device: work that does not depend on the recurrence of the context network (the
    negative-sample draws and their index arithmetic) runs on it WHILE the recurrence - 128 dependent steps on 32 of the
    148 SMs - runs on the caller's stream.  Fork and join are events, so the caller's stream order is preserved and the
    pattern is captured as parallel branches by a CUDA graph."""
    key = (device.type, device.index)
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return st


def overlap_enabled():
    # opt-in: measured on B200 (r2k) the fork saves 7 us of a 1.46 ms step (the draws and the index arithmetic are short
    # and already hidden by programmatic dependent launch)
    return os.environ.get("CPC_B200_OVERLAP", "0") == "1"


class ChannelNorm(nn.Module):
    """Parameter holder for cpc/model.py:25-58 (weight/bias of shape (1, C, 1)); the math is fused in the kernels."""

    def __init__(self, numFeatures, epsilon=1e-05, affine=True):
        super().__init__()
        if not affine or epsilon != 1e-05:
            raise NotImplementedError("cpc_audio_b200: ChannelNorm supports affine=True, epsilon=1e-5 only")
        self.weight = nn.Parameter(torch.ones(1, numFeatures, 1))
        self.bias = nn.Parameter(torch.zeros(1, numFeatures, 1))
        self.epsilon = epsilon
        self.affine = affine


class _EncoderFn(torch.autograd.Function):
    """x (B,1,L) -> z (B,S,H) channel-last.  cpc/model.py:99-105."""

    @staticmethod
    def forward(ctx, x, dtype_code, *params):
        _require_cuda(x, "CPCEncoder")
        lib = L.lib()
        B, one, Lw = x.shape
        H = params[0].shape[0]
        dev = x.device
        x = L.f32c(x)
        params = L.cparams(params)
        d, save_n, wsn = _plan("enc", B, Lw, H, H, 1, 1, 1, dtype_code)
        z = torch.empty(B, frames_for(Lw), H, device=dev, dtype=torch.float32)
        save = _bytes(save_n, dev)
        ws = _bytes(wsn[0], dev)
        with L.device_guard(dev):
            L.check(lib.cpcb200_encoder_fwd(d, L.ptr(x), _encoder_params(params), L.ptr(z), L.ptr(save), L.ptr(ws), wsn[0],
                                            L.stream_ptr(dev)), "encoder_fwd")
        ctx.save_for_backward(x, save)
        ctx.params = params
        ctx.dims = (B, Lw, H, dtype_code)
        return z

    @staticmethod
    def backward(ctx, dz):
        lib = L.lib()
        x, save = ctx.saved_tensors
        params = ctx.params
        B, Lw, H, dtype_code = ctx.dims
        dev = x.device
        d, _, wsn = _plan("enc", B, Lw, H, H, 1, 1, 1, dtype_code)
        grads, sunk = _grad_targets(params, dev)
        ws = _bytes(wsn[1], dev)
        dz = L.f32c(dz)
        with L.device_guard(dev):
            L.check(lib.cpcb200_encoder_bwd(d, L.ptr(x), _encoder_params(params), L.ptr(dz), L.ptr(save),
                                            _encoder_params(grads), L.ptr(ws), wsn[1], L.stream_ptr(dev)), "encoder_bwd")
        return (None, None, *([None] * len(grads) if sunk else grads))


def _encoder_infer(x, dtype_code, params):
    """no_grad forward (valStep cpc/train.py:141-142, feature extraction feature_loader.py:33): nothing is saved for a
    backward pass - the intermediate activations ping-pong inside one workspace, pre-norm rows are never written."""
    _require_cuda(x, "CPCEncoder")
    lib = L.lib()
    B, one, Lw = x.shape
    H = params[0].shape[0]
    dev = x.device
    x = L.f32c(x)
    params = L.cparams([p.detach() for p in params])
    d, _, wsn = _plan("enc", B, Lw, H, H, 1, 1, 1, dtype_code)
    z = torch.empty(B, frames_for(Lw), H, device=dev, dtype=torch.float32)
    ws = _bytes(wsn[2], dev)
    with L.device_guard(dev):
        L.check(lib.cpcb200_encoder_fwd(d, L.ptr(x), _encoder_params(params), L.ptr(z), None, L.ptr(ws), wsn[2], L.stream_ptr(dev)),
                "encoder_fwd")
    return z


def _encoder_params(ts):
    ep = L.EncoderParams()
    for i in range(5):
        ep.conv_w[i] = ts[4 * i].data_ptr()
        ep.conv_b[i] = ts[4 * i + 1].data_ptr()
        ep.norm_w[i] = ts[4 * i + 2].data_ptr()
        ep.norm_b[i] = ts[4 * i + 3].data_ptr()
    return ep


class CPCEncoder(nn.Module):
    """cpc/model.py:61-105.  Returns (B, H, S) like the reference (a permuted view of the channel-last result)."""

    def __init__(self, sizeHidden=512, normMode="layerNorm", compute_dtype=None):
        super().__init__()
        validModes = ["batchNorm", "instanceNorm", "ID", "layerNorm"]
        if normMode not in validModes:
            raise ValueError(f"Norm mode must be in {validModes}")
        if normMode != "layerNorm":
            raise NotImplementedError(f"cpc_audio_b200: normMode={normMode!r} is outside the accelerated hot path "
                                      f"(only 'layerNorm' = ChannelNorm); no fallback is provided")
        if sizeHidden % 64 != 0 or not 64 <= sizeHidden <= 512:
            raise NotImplementedError("cpc_audio_b200: hiddenEncoder must be a multiple of 64 in [64, 512]")
        self.dimEncoded = sizeHidden
        # nn.Conv1d modules are parameter holders (same keys and default init as the reference); never called
        self.conv0 = nn.Conv1d(1, sizeHidden, 10, stride=5, padding=3)
        self.batchNorm0 = ChannelNorm(sizeHidden)
        self.conv1 = nn.Conv1d(sizeHidden, sizeHidden, 8, stride=4, padding=2)
        self.batchNorm1 = ChannelNorm(sizeHidden)
        self.conv2 = nn.Conv1d(sizeHidden, sizeHidden, 4, stride=2, padding=1)
        self.batchNorm2 = ChannelNorm(sizeHidden)
        self.conv3 = nn.Conv1d(sizeHidden, sizeHidden, 4, stride=2, padding=1)
        self.batchNorm3 = ChannelNorm(sizeHidden)
        self.conv4 = nn.Conv1d(sizeHidden, sizeHidden, 4, stride=2, padding=1)
        self.batchNorm4 = ChannelNorm(sizeHidden)
        self.DOWNSAMPLING = 160
        self.compute_dtype = compute_dtype or default_dtype()

    def getDimOutput(self):
        return self.conv4.out_channels

    def _params(self):
        out = []
        for i in range(5):
            conv, norm = getattr(self, f"conv{i}"), getattr(self, f"batchNorm{i}")
            out += [conv.weight, conv.bias, norm.weight, norm.bias]
        return out

    def forward_channel_last(self, x):
        if x.dim() != 3 or x.size(1) != 1 or frames_for(x.size(2)) < 1:
            raise ValueError(f"CPCEncoder expects (B, 1, L) with L long enough for one output frame, got {tuple(x.shape)}")
        params = self._params()
        if not torch.is_grad_enabled() or not (x.requires_grad or any(p.requires_grad for p in params)):
            return _encoder_infer(x, _dtype_code(self.compute_dtype), params)
        return _EncoderFn.apply(x, _dtype_code(self.compute_dtype), *params)

    def forward(self, x):
        return self.forward_channel_last(x).permute(0, 2, 1)


class _GruFn(torch.autograd.Function):
    """z (B,S,H), h0 (nL,B,Har)|None -> c (B,S,Har), hT (nL,B,Har).  torch.nn.GRU(batch_first=True) semantics."""

    @staticmethod
    def forward(ctx, z, h0, dtype_code, n_layers, *params):
        _require_cuda(z, "CPCAR")
        lib = L.lib()
        B, S, H = z.shape
        Har = params[1].shape[1]
        dev = z.device
        z = L.f32c(z)
        h0c = L.f32c(h0) if h0 is not None else None
        params = L.cparams(params)
        d, save_n, wsn = _plan("gru", B, S * 160, H, Har, 1, 1, n_layers, dtype_code)
        c = torch.empty(B, S, Har, device=dev, dtype=torch.float32)
        hT = torch.empty(n_layers, B, Har, device=dev, dtype=torch.float32)
        save = _bytes(save_n, dev)
        ws = _bytes(wsn[0], dev)
        with L.device_guard(dev):
            L.check(lib.cpcb200_gru_fwd(d, L.ptr(z), L.ptr(h0c), _gru_params(params, n_layers), L.ptr(c), L.ptr(hT),
                                        L.ptr(save), L.ptr(ws), wsn[0], L.stream_ptr(dev)), "gru_fwd")
        ctx.save_for_backward(z, c, save)
        ctx.params = params
        ctx.h0 = h0c
        ctx.dims = (B, S, H, Har, n_layers, dtype_code)
        ctx.mark_non_differentiable(hT)
        return c, hT

    @staticmethod
    def backward(ctx, dc, _dhT):
        lib = L.lib()
        z, c, save = ctx.saved_tensors
        params = ctx.params
        B, S, H, Har, n_layers, dtype_code = ctx.dims
        dev = z.device
        d, _, wsn = _plan("gru", B, S * 160, H, Har, 1, 1, n_layers, dtype_code)
        grads, sunk = _grad_targets(params, dev)
        dz = torch.empty_like(z)
        ws = _bytes(wsn[1], dev)
        dc = L.f32c(dc)
        with L.device_guard(dev):
            L.check(lib.cpcb200_gru_bwd(d, L.ptr(z), L.ptr(ctx.h0), _gru_params(params, n_layers), L.ptr(c), L.ptr(dc),
                                        L.ptr(save), L.ptr(dz), _gru_params(grads, n_layers), L.ptr(ws), wsn[1],
                                        L.stream_ptr(dev)), "gru_bwd")
        return (dz, None, None, None, *([None] * len(grads) if sunk else grads))


def _gru_params(ts, n_layers):
    gp = L.GruParams()
    for l in range(n_layers):
        gp.w_ih[l] = ts[4 * l].data_ptr()
        gp.w_hh[l] = ts[4 * l + 1].data_ptr()
        gp.b_ih[l] = ts[4 * l + 2].data_ptr()
        gp.b_hh[l] = ts[4 * l + 3].data_ptr()
    return gp


class _LstmFn(torch.autograd.Function):
    """z (B,S,H), h0/c0 (nL,B,Har)|None -> out (B,S,Har), hT, cT.  torch.nn.LSTM(batch_first=True) semantics."""

    @staticmethod
    def forward(ctx, z, h0, c0, dtype_code, n_layers, train, *params):
        _require_cuda(z, "CPCAR")
        lib = L.lib()
        B, S, H = z.shape
        Har = params[1].shape[1]
        dev = z.device
        z = L.f32c(z)
        h0c = L.f32c(h0) if h0 is not None else None
        c0c = L.f32c(c0) if c0 is not None else None
        params = L.cparams(params)
        d, save_n, wsns = _plan("lstm", B, S * 160, H, Har, 1, 1, n_layers, dtype_code)
        out = torch.empty(B, S, Har, device=dev, dtype=torch.float32)
        hT = torch.empty(n_layers, B, Har, device=dev, dtype=torch.float32)
        cT = torch.empty(n_layers, B, Har, device=dev, dtype=torch.float32)
        save = _bytes(save_n, dev) if train else None
        wsn = wsns[0 if train else 2]
        ws = _bytes(wsn, dev)
        with L.device_guard(dev):
            L.check(lib.cpcb200_lstm_fwd(d, L.ptr(z), L.ptr(h0c), L.ptr(c0c), _gru_params(params, n_layers), L.ptr(out), L.ptr(hT),
                                         L.ptr(cT), L.ptr(save), L.ptr(ws), wsn, L.stream_ptr(dev)), "lstm_fwd")
        if train:
            ctx.save_for_backward(z, out, save)
        ctx.params = params
        ctx.h0, ctx.c0 = h0c, c0c
        ctx.dims = (B, S, H, Har, n_layers, dtype_code)
        ctx.mark_non_differentiable(hT, cT)
        return out, hT, cT

    @staticmethod
    def backward(ctx, dout, _dhT, _dcT):
        lib = L.lib()
        z, out, save = ctx.saved_tensors
        params = ctx.params
        B, S, H, Har, n_layers, dtype_code = ctx.dims
        dev = z.device
        d, _, wsns = _plan("lstm", B, S * 160, H, Har, 1, 1, n_layers, dtype_code)
        grads, sunk = _grad_targets(params, dev)
        dz = torch.empty_like(z)
        wsn = wsns[1]
        ws = _bytes(wsn, dev)
        dout = L.f32c(dout)
        with L.device_guard(dev):
            L.check(lib.cpcb200_lstm_bwd(d, L.ptr(z), L.ptr(ctx.h0), L.ptr(ctx.c0), _gru_params(params, n_layers), L.ptr(out),
                                         L.ptr(dout), L.ptr(save), L.ptr(dz), _gru_params(grads, n_layers), L.ptr(ws), wsn,
                                         L.stream_ptr(dev)), "lstm_bwd")
        return (dz, None, None, None, None, None, *([None] * len(grads) if sunk else grads))


class CPCAR(nn.Module):
    """cpc/model.py:155-204: GRU (--arMode GRU) and LSTM (--arMode LSTM, the reference default) context networks.
    RNN / reverse are outside the accelerated path and raise."""

    def __init__(self, dimEncoded, dimOutput, keepHidden, nLevelsGRU, mode="GRU", reverse=False, compute_dtype=None):
        super().__init__()
        self.RESIDUAL_STD = 0.1
        if mode == "RNN":
            raise NotImplementedError("cpc_audio_b200: arMode='RNN' is outside the accelerated hot path (GRU, LSTM and "
                                      "transformer context networks are implemented)")
        if reverse:
            raise NotImplementedError("cpc_audio_b200: cpc_mode='reverse' is outside the accelerated hot path")
        if dimOutput % 64 != 0 or not 64 <= dimOutput <= 512 or not 1 <= nLevelsGRU <= L.MAX_GRU_LAYERS:
            raise NotImplementedError("cpc_audio_b200: hiddenGar must be a multiple of 64 in [64, 512], nLevelsGRU in [1, 4]")
        # parameter holder: same keys (gAR.baseNet.weight_ih_l0 ...) and default init as the reference (model.py:171-179:
        # anything that is not 'LSTM' / 'RNN' builds a GRU)
        self.lstm = mode == "LSTM"
        cls = nn.LSTM if self.lstm else nn.GRU
        self.baseNet = cls(dimEncoded, dimOutput, num_layers=nLevelsGRU, batch_first=True)
        self.hidden = None
        self.keepHidden = keepHidden
        self.reverse = reverse
        self.compute_dtype = compute_dtype or default_dtype()

    def getDimOutput(self):
        return self.baseNet.hidden_size

    def _params(self):
        out = []
        for l in range(self.baseNet.num_layers):
            out += [getattr(self.baseNet, f"{n}_l{l}") for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
        return out

    def forward(self, x):
        params = self._params()
        code, nl = _dtype_code(self.compute_dtype), self.baseNet.num_layers
        if self.lstm:
            h0, c0 = self.hidden if self.hidden is not None else (None, None)
            train = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))
            c, hT, cT = _LstmFn.apply(x, h0, c0, code, nl, train, *params)
            if self.keepHidden:
                self.hidden = (hT.detach(), cT.detach())  # model.py:194-196: a tuple for the LSTM
            return c
        c, hT = _GruFn.apply(x, self.hidden, code, nl, *params)
        if self.keepHidden:
            self.hidden = hT.detach()  # model.py:194-198
        return c


class CPCModel(nn.Module):
    """cpc/model.py:276-289.  forward(batchData, label) -> (cFeature (B,S,Har), encodedData (B,S,H), label)."""

    def __init__(self, encoder, AR):
        super().__init__()
        self.gEncoder = encoder
        self.gAR = AR

    def forward(self, batchData, label):
        if isinstance(self.gEncoder, CPCEncoder):
            encodedData = self.gEncoder.forward_channel_last(batchData)
        else:
            encodedData = self.gEncoder(batchData).permute(0, 2, 1)
        if encodedData.is_cuda and overlap_enabled():
            # 'z is ready': the criterion forks its generator draws / index arithmetic from here, beside the recurrence
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(encodedData.device))
            encodedData._cpcb200_ready = ev
        cFeature = self.gAR(encodedData)
        return cFeature, encodedData, label
