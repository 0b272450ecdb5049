"""Time the two Adam entry points in isolation (CUDA events, 200 launches each).  Development tool."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpc_audio_b200 import _lib as L

lib = L.lib()
dev = torch.device("cuda:0")
n = 2498304
p, g, m, v = (torch.randn(n, device=dev) * 0.01 for _ in range(4))
v = v.abs()
state = torch.zeros(8, dtype=torch.int32, device=dev)
st = L.stream_ptr(dev)


def timeit(fn, reps=200):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


step = [0]
def a():
    step[0] += 1
    L.check(lib.cpcb200_adam_step(L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), n, 2e-4, 0.9, 0.999, 1e-8, 0.0, step[0], st), "adam")
def b():
    L.check(lib.cpcb200_adam_step_dev(L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), n, 2e-4, 0.9, 0.999, 1e-8, 0.0, L.ptr(state), 0, st), "adam_dev")
def c():
    L.check(lib.cpcb200_adam_step_dev(L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), n, 2e-4, 0.9, 0.999, 1e-8, 0.0, L.ptr(state), 1, st), "adam_dev")

print(f"adam_step          {timeit(a):7.2f} us")
print(f"adam_step_dev z=0  {timeit(b):7.2f} us   steps={state[0].item()}")
print(f"adam_step_dev z=1  {timeit(c):7.2f} us   steps={state[0].item()}")
g.normal_()
print(f"adam_step_dev z=0 (fresh g)  {timeit(b):7.2f} us")
