"""Where does the host time of the train.py-driven loop go?  python tools/train_py_probe.py [B]
Runs the reference's trainStep over the B200 modules at batch B (default 64) and at B = 1 (GPU time negligible: the
loop time is then the host's enqueue time), then prints a torch.profiler table of the host side.  Development tool."""
import contextlib
import io
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import ref_driver as R  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
FLAT = "--flat-adam" in sys.argv
dev = torch.device("cuda", 0)
args = R.default_args(arMode="GRU", rnnMode="linear")
ref = R.reference_or_none()


def run(b, steps=30, prof=False):
    import cpc_audio_b200.patch as patch
    if FLAT:
        patch.install_adam()
    model, crit, mdp, cdp, opt = R.build(args, b200=True, seed=0, device=dev)
    patch.uninstall_adam()
    R.use_b200_modules(False)
    x = (torch.randn(b, 1, 20480) * 0.1).pin_memory()
    lab = torch.zeros(b, dtype=torch.long)
    with contextlib.redirect_stdout(io.StringIO()):
        ref.train.trainStep([(x, lab)] * 5, mdp, cdp, opt, None, 10 ** 9)
        torch.cuda.synchronize()
        if prof:
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as p:
                ref.train.trainStep([(x, lab)] * 10, mdp, cdp, opt, None, 10 ** 9)
                torch.cuda.synchronize()
            return p
        t0 = time.perf_counter()
        ref.train.trainStep([(x, lab)] * steps, mdp, cdp, opt, None, 10 ** 9)
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3


print(f"flat_adam={FLAT}  B={B}: {run(B):.3f} ms/step   B=1 (host-bound): {run(1):.3f} ms/step")
p = run(B, prof=True)
print(p.key_averages().table(sort_by="self_cpu_time_total", row_limit=60, max_name_column_width=60))
