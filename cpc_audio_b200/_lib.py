"""ctypes binding of libcpc_b200.so (C ABI declared in include/cpc_b200.h).

The product path has NO fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcpc_b200.so")

F32, BF16 = 0, 1
MAX_GRU_LAYERS = 4


class Dims(C.Structure):
    _fields_ = [("B", C.c_int32), ("L", C.c_int32), ("H", C.c_int32), ("Har", C.c_int32), ("K", C.c_int32),
                ("N", C.c_int32), ("nLayers", C.c_int32), ("dtype", C.c_int32)]


class EncoderParams(C.Structure):
    _fields_ = [("conv_w", C.c_void_p * 5), ("conv_b", C.c_void_p * 5), ("norm_w", C.c_void_p * 5),
                ("norm_b", C.c_void_p * 5)]


class GruParams(C.Structure):
    _fields_ = [("w_ih", C.c_void_p * MAX_GRU_LAYERS), ("w_hh", C.c_void_p * MAX_GRU_LAYERS),
                ("b_ih", C.c_void_p * MAX_GRU_LAYERS), ("b_hh", C.c_void_p * MAX_GRU_LAYERS)]


class THeadParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("wq", "wk", "wv", "wo", "krelpos", "ln1_w", "ln1_b", "w1", "b1", "w2", "b2",
                                          "ln2_w", "ln2_b")] + [("dff", C.c_int32), ("nheads", C.c_int32),
                                                                ("att_keep", C.c_void_p), ("ffn_keep", C.c_void_p),
                                                                ("keep_scale", C.c_float)]


THEAD_FIELDS = ("wq", "wk", "wv", "wo", "krelpos", "ln1_w", "ln1_b", "w1", "b1", "w2", "b2", "ln2_w", "ln2_b")

_P, _SZ, _I = C.c_void_p, C.c_size_t, C.c_int
class Peers(C.Structure):
    """cpcb200_peers: peer-mapped gradient buckets and signal words of the GPUs of one node."""
    _fields_ = [("grads", C.c_void_p * 8), ("signals", C.c_void_p * 8), ("rank", C.c_int32), ("world", C.c_int32),
                ("grads_mc", C.c_void_p), ("timeout_ns", C.c_int64)]


_DP = C.POINTER(Dims)

# name -> (restype, argtypes).  Must list every symbol include/cpc_b200.h declares (tests/test_abi.py checks).
SIGNATURES = {
    "cpcb200_version": (C.c_int, []),
    "cpcb200_last_error": (C.c_char_p, []),
    "cpcb200_launch_count": (C.c_uint64, []),
    "cpcb200_prof_enable": (_I, [_I]),
    "cpcb200_prof_report": (_I, [C.c_char_p, _SZ]),
    "cpcb200_gather_windows": (_I, [_P, C.c_int64, _P, _I, _I, _P, _P, _I, _P, _P, _P]),
    "cpcb200_encoder_save_bytes": (_SZ, [_DP]),
    "cpcb200_encoder_ws_bytes": (_SZ, [_DP, _I]),
    "cpcb200_encoder_fwd": (_I, [_DP, _P, C.POINTER(EncoderParams), _P, _P, _P, _SZ, _P]),
    "cpcb200_encoder_bwd": (_I, [_DP, _P, C.POINTER(EncoderParams), _P, _P, C.POINTER(EncoderParams), _P, _SZ, _P]),
    "cpcb200_gru_save_bytes": (_SZ, [_DP]),
    "cpcb200_gru_ws_bytes": (_SZ, [_DP, _I]),
    "cpcb200_gru_fwd": (_I, [_DP, _P, _P, C.POINTER(GruParams), _P, _P, _P, _P, _SZ, _P]),
    "cpcb200_gru_bwd": (_I, [_DP, _P, _P, C.POINTER(GruParams), _P, _P, _P, _P, C.POINTER(GruParams), _P, _SZ, _P]),
    "cpcb200_lstm_save_bytes": (_SZ, [_DP]),
    "cpcb200_lstm_ws_bytes": (_SZ, [_DP, _I]),
    "cpcb200_lstm_fwd": (_I, [_DP, _P, _P, _P, C.POINTER(GruParams), _P, _P, _P, _P, _P, _SZ, _P]),
    "cpcb200_lstm_bwd": (_I, [_DP, _P, _P, _P, C.POINTER(GruParams), _P, _P, _P, _P, C.POINTER(GruParams), _P, _SZ, _P]),
    "cpcb200_tlayer_save_bytes": (_SZ, [_DP, _I, _I]),
    "cpcb200_tlayer_ws_bytes": (_SZ, [_DP, _I, _I, _I]),
    "cpcb200_tlayer_fwd": (_I, [_DP, _P, C.POINTER(THeadParams), _P, _P, _P, _SZ, _P]),
    "cpcb200_tlayer_bwd": (_I, [_DP, _P, C.POINTER(THeadParams), _P, _P, _P, C.POINTER(THeadParams), _P, _SZ, _P]),
    "cpcb200_sample_ext_idx": (_I, [_DP, _P, _P, _P, _P]),
    "cpcb200_criterion_save_bytes": (_SZ, [_DP]),
    "cpcb200_criterion_ws_bytes": (_SZ, [_DP, _I]),
    "cpcb200_criterion_fwd": (_I, [_DP, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "cpcb200_criterion_bwd": (_I, [_DP, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "cpcb200_criterion_t_save_bytes": (_SZ, [_DP, _I, _I]),
    "cpcb200_criterion_t_ws_bytes": (_SZ, [_DP, _I, _I, _I]),
    "cpcb200_criterion_t_fwd": (_I, [_DP, _P, _P, C.POINTER(THeadParams), _P, _P, _P, _P, _P, _SZ, _P]),
    "cpcb200_criterion_t_bwd": (_I, [_DP, _P, _P, C.POINTER(THeadParams), _P, _P, _P, _P, _P, C.POINTER(THeadParams), _P, _SZ, _P]),
    "cpcb200_adam_step": (_I, [_P, _P, _P, _P, _SZ, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32, _P]),
    "cpcb200_encoder_bwd_set_event": (_I, [_P, _P]),
    "cpcb200_allreduce_adam_step": (_I, [_P, _P, _P, _P, _SZ, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _P, C.c_int,
                                         _P, C.c_int, _P]),
    "cpcb200_peer_reduce_range": (_I, [_P, _P, C.c_int, _P, _P]),
    "cpcb200_adam_step_dev": (_I, [_P, _P, _P, _P, _SZ, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _P, C.c_int, _P]),
    "cpcb200_test_gemm_nt": (_I, [_I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "cpcb200_test_gemm_tn": (_I, [_I, _I, _I, _I, _P, _P, _P, _P]),
    "cpcb200_test_gemm_nt_act": (_I, [_I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "cpcb200_debug_gemm_timeline": (_I, [_P]),
}

_lib = None


def lib():
    """Load the shared library once; raise loudly if it is not built (no CPU / eager fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"cpc_audio_b200: {LIB_PATH} is missing - run `python -c 'import __graft_entry__ as g; "
                               f"g.build()'` (nvcc, sm_100a). There is no fallback path.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(status: int, what: str):
    if status != 0:
        msg = lib().cpcb200_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"cpc_audio_b200: {what} failed with status {status}: {msg}")


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def stream_ptr(device):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class device_guard:
    """``with torch.cuda.device(dev)`` without its cost when `dev` is already the current device (the usual case: one
    process per GPU); DataParallel worker threads still get their own device set."""
    __slots__ = ("idx", "prev")

    def __init__(self, device):
        self.idx = device.index if device.index is not None else 0
        self.prev = None

    def __enter__(self):
        import torch
        cur = torch.cuda.current_device()
        if cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)
        return self

    def __exit__(self, *exc):
        if self.prev is not None:
            import torch
            torch.cuda.set_device(self.prev)
        return False


def f32c(t):
    """`t` as a contiguous fp32 tensor, without a dispatcher round trip when it already is one."""
    import torch
    if t.dtype is torch.float32 and t.is_contiguous():
        return t
    return t.contiguous().float()


def cparams(params):
    return tuple(p if p.is_contiguous() else p.contiguous() for p in params)


def make_dims(B, L, H, Har, K, N, nLayers, dtype):
    return Dims(int(B), int(L), int(H), int(Har), int(K), int(N), int(nLayers), int(dtype))
