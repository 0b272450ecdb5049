"""GPU micro-benchmark of the NT GEMM building block at the shapes the hot path uses (dense A, bf16 in / bf16 out).

    python tools/gemm_probe.py            # prints us per launch, TFLOP/s and the per-CTA cycle timeline

Reads the debug timeline hook (cpcb200_debug_gemm_timeline).  Development tool, not part of the product path.
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpc_audio_b200 import _lib  # noqa: E402

SHAPES = [  # (name, M, N, K)
    ("conv1 fwd", 65536, 256, 2048), ("conv2 fwd", 32768, 256, 1024), ("conv3 fwd", 16384, 256, 1024),
    ("conv4 fwd", 8192, 256, 1024), ("conv1 dgrad", 65536, 1024, 512), ("conv2 dgrad", 32768, 512, 512),
    ("conv4 dgrad", 8192, 512, 512), ("heads", 7424, 3072, 256), ("gru in", 8192, 768, 256),
    ("long-K", 148 * 128, 256, 16384),
]


def main():
    lib = _lib.lib()
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tl = np.zeros(148 * 8, dtype=np.uint64)
    for name, M, N, K in SHAPES:
        a = (torch.randn(M, K, device=dev) * 0.1).to(torch.bfloat16)
        b = (torch.randn(N, K, device=dev) * 0.1).to(torch.bfloat16)
        c = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        def run():
            r = lib.cpcb200_test_gemm_nt_act(1, M, N, K, _lib.ptr(a), _lib.ptr(b), None, _lib.ptr(c), C.c_void_p(st))
            _lib.check(r, 'gemm_nt_act')
        for _ in range(3):
            run()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run()
        e1.record(); torch.cuda.synchronize()
        warm = e0.elapsed_time(e1) * 1e3 / 20
        us = float(np.median(ts))
        lib.cpcb200_debug_gemm_timeline(tl.ctypes.data_as(C.c_void_p))
        t = tl.reshape(148, 8).astype(np.float64)
        act = t[:, 6] > 0
        med = np.median(t[act], axis=0)
        ref = torch.matmul(a[:256].float(), b.float().t())
        err = (c[:256].float() - ref).abs().max().item() / ref.abs().max().item()
        print(f"{name:12s} M={M:6d} N={N:5d} K={K:5d}  cold {us:7.1f} us  warm {warm:7.1f} us  {2.0 * M * N * K / warm / 1e6:7.1f} TF/s  relerr {err:.1e}")
        print("    cycles: prologue %d | 1st TMA %d | 1st data %d | last MMA %d | 1st acc %d | last acc %d | last epi %d | end %d"
              % (med[0], med[1], med[2], med[3], med[4], med[7], med[5], med[6]))


if __name__ == "__main__":
    main()
