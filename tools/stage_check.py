"""Run each stage of the bf16 path in its own subprocess with a timeout, so that a hanging or faulting kernel
is localised in one GPU call.   python tools/stage_check.py [B]"""
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGES = ["encoder_fwd", "encoder_bwd", "gru_fwd", "gru_bwd", "crit_fwd", "crit_bwd", "full_f32_vs_bf16"]

CHILD = r'''
import sys, time, torch
sys.path.insert(0, %r)
stage, B = sys.argv[1], int(sys.argv[2])
from oracle import cpc_oracle as O
from tests import helpers as Hh
import cpc_audio_b200 as M
d = O.Dims(B=B, L=20480, H=256, Har=256, K=12, N=128, nLayers=1)
mp, cp = O.make_params(d, seed=3, pred_scale=30.0)
x, label = O.make_batch(d, seed=5)
bi, si = O.make_raw_indices(d, seed=6)
def mods(dt):
    m, c = Hh.build_modules(d, mp, cp, dt)
    c.sampleIndices = lambda B, W, S, device: (bi.to(device), si.to(device))
    return m, c
t0 = time.time()
def done(msg=""):
    torch.cuda.synchronize(); print(f"  {stage}: ok {time.time()-t0:.2f}s {msg}", flush=True)
m, c = mods("bf16")
xc = x.cuda()
if stage == "encoder_fwd":
    z = m.gEncoder.forward_channel_last(xc); done(f"z norm {z.norm().item():.3f}")
elif stage == "encoder_bwd":
    z = m.gEncoder.forward_channel_last(xc); z.square().sum().backward(); done(f"gw1 {m.gEncoder.conv1.weight.grad.norm().item():.3e}")
elif stage in ("gru_fwd", "gru_bwd"):
    z = torch.randn(B, 128, 256, device="cuda", requires_grad=True)
    cc = m.gAR(z)
    if stage == "gru_bwd": cc.square().sum().backward()
    done(f"c norm {cc.norm().item():.3f}")
elif stage in ("crit_fwd", "crit_bwd"):
    z = torch.randn(B, 128, 256, device="cuda").relu().requires_grad_(True); cc = torch.randn(B, 128, 256, device="cuda").tanh().requires_grad_(True)
    l, a = c(cc, z, label.cuda())
    if stage == "crit_bwd": l.sum().backward()
    done(f"loss {l.flatten()[:3].tolist()}")
else:
    o16 = Hh.run_modules(m, c, x, label, bi, si)
    m32, c32 = mods("f32")
    o32 = Hh.run_modules(m32, c32, x, label, bi, si)
    dl = (o16["losses"] - o32["losses"]).abs().max().item()
    worst = min(torch.nn.functional.cosine_similarity(o16["grads"][k].flatten().double(), o32["grads"][k].flatten().double(), dim=0).item() for k in o32["grads"])
    done(f"max|dloss| {dl:.3e} worst grad cosine {worst:.4f} z rel {Hh.rel_err(o16['z'], o32['z']):.2e} c rel {Hh.rel_err(o16['c'], o32['c']):.2e}")
''' % REPO

if __name__ == "__main__":
    B = sys.argv[1] if len(sys.argv) > 1 else "4"
    for st in STAGES:
        try:
            r = subprocess.run([sys.executable, "-c", CHILD, st, B], capture_output=True, text=True, timeout=120)
            out = (r.stdout + r.stderr).strip().splitlines()
            print(f"[{st}] rc={r.returncode}")
            for line in out[-6:]:
                print("   ", line)
        except subprocess.TimeoutExpired:
            print(f"[{st}] TIMEOUT (hang)")
