"""Every dense product of one training step (BASELINE config 2, B = 64) as a plain GEMM of the same shape: this library's
tcgen05 kernels (through the cpcb200_test_gemm_* hooks) against torch.matmul (cuBLAS) in bf16 on the same GPU.
Prints a markdown table (committed as profiles/r2_gemm_shapes.md).  Development / evidence tool."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpc_audio_b200 import _lib as L  # noqa: E402

lib = L.lib()
dev = torch.device("cuda", 0)
B, H, S, W, K = 64, 256, 128, 116, 12
NT = [  # (what, M, N, Kd)
    ("conv1 fwd (k=8)", B * 1024, H, 8 * H), ("conv2 fwd (k=4)", B * 512, H, 4 * H), ("conv3 fwd", B * 256, H, 4 * H),
    ("conv4 fwd", B * 128, H, 4 * H),
    ("dgrad1 (s=4)", B * 1025, 4 * H, 2 * H), ("dgrad2 (s=2)", B * 513, 2 * H, 2 * H), ("dgrad3", B * 257, 2 * H, 2 * H),
    ("dgrad4", B * 129, 2 * H, 2 * H),
    ("GRU input proj", B * S, 3 * H, H), ("GRU d(input)", B * S, H, 3 * H),
    ("heads fwd", B * W, K * H, H), ("heads dc", B * W, H, K * H),
]
TN = [  # (what, M rows reduced, N1, N2)
    ("wgrad conv1", B * 1024, H, 8 * H), ("wgrad conv2", B * 512, H, 4 * H), ("wgrad conv3", B * 256, H, 4 * H),
    ("wgrad conv4", B * 128, H, 4 * H), ("wgrad GRU W_ih", B * S, 3 * H, H), ("wgrad GRU W_hh", B * (S - 1), 3 * H, H),
    ("wgrad heads", B * W, K * H, H),
]


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)  # 256 MB > L2
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e3  # us


g = torch.Generator(device=dev).manual_seed(0)
print("| product | M | N | K | ours (us) | ours TF/s | torch.matmul bf16 (us) | cuBLAS TF/s | ours / cuBLAS |")
print("|---|---|---|---|---|---|---|---|---|")
tot_o = tot_t = 0.0
for what, M, N, Kd in NT:
    A = torch.randn(M, Kd, device=dev, generator=g).bfloat16()
    Bm = torch.randn(N, Kd, device=dev, generator=g).bfloat16()
    C = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    st = L.stream_ptr(dev)
    ours = timeit(lambda: L.check(lib.cpcb200_test_gemm_nt_act(L.BF16, M, N, Kd, L.ptr(A), L.ptr(Bm), None, L.ptr(C), st), "nt"))
    ref = timeit(lambda: torch.matmul(A, Bm.t()))
    fl = 2.0 * M * N * Kd
    tot_o += ours; tot_t += ref
    print(f"| {what} | {M} | {N} | {Kd} | {ours:.1f} | {fl / ours / 1e6:.0f} | {ref:.1f} | {fl / ref / 1e6:.0f} | {ref / ours:.2f} |")
for what, M, N1, N2 in TN:
    A = torch.randn(M, N1, device=dev, generator=g).bfloat16()
    Bm = torch.randn(M, N2, device=dev, generator=g).bfloat16()
    C = torch.zeros(N1, N2, device=dev)
    st = L.stream_ptr(dev)
    ours = timeit(lambda: L.check(lib.cpcb200_test_gemm_tn(L.BF16, M, N1, N2, L.ptr(A), L.ptr(Bm), L.ptr(C), st), "tn"))
    ref = timeit(lambda: torch.matmul(A.t(), Bm))
    fl = 2.0 * M * N1 * N2
    tot_o += ours; tot_t += ref
    print(f"| {what} (A^T B, fp32 += ) | {M} | {N1} | {N2} | {ours:.1f} | {fl / ours / 1e6:.0f} | {ref:.1f} | {fl / ref / 1e6:.0f} | {ref / ours:.2f} |")
print(f"\nsum over the {len(NT) + len(TN)} products: ours {tot_o:.0f} us, torch.matmul {tot_t:.0f} us "
      f"(cold L2: a 256 MB buffer is rewritten before every timed call; in the step the conv products are implicit GEMMs over "
      f"a strided row view with ChannelNorm + ReLU in the epilogue and the weight gradients run as ONE grouped launch - this table "
      f"isolates the GEMM kernels on plain operands)")
