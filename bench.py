#!/usr/bin/env python
"""bench.py - CPC training-step throughput on B200 (audio-seconds/sec on 20480-sample windows).

    python bench.py --gpus 1 --steps 20 --warmup 5                 # our arm (N=1)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W                  # our arm, one rank per GPU over NCCL
    python bench.py --impl reference --steps 3 --warmup 1          # reference arm: CPU path on the host cores

A step = CPCModel.forward + CPCUnsupersivedCriterion.forward + backward (+ gradient all-reduce when N > 1)
+ Adam step + zero_grad, exactly the body of cpc/train.py:83-91, on BASELINE.json config 2
(CPC default: H=Har=256, 1-layer GRU, K=12, 128 negatives, 64 windows of 20480 samples per GPU, bf16 storage /
fp32 accumulation).  Weak scaling: every rank processes its own 64 windows; the only collective is one NCCL
all-reduce (sum, cpc/train.py:85 semantics) over a flat fp32 gradient bucket.

One JSON line is printed by rank 0 (keys: see the driver contract in the task statement) with
  value     : whole-job audio-s/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e       : same through the public nn.Module API with HOST (pinned) inputs: H2D of the batch and D2H of the
              losses inside the timed region
  roofline  : dominant kernel of the step - algorithmic bytes (or FLOPs) / CUDA-event duration vs MEASURED_PEAKS
  cpu_baseline : the reference algorithm (oracle port, same torch CPU ops as the reference) on the host cores
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

WINDOW = 20480
SR = 16000.0
L2_RED_CEILING_GBS = 5400.0  # fp32 scatter-add into an 8 MB L2-resident buffer, measured with tools/redbench.cu (B200)
T0 = time.time()


def log(msg):
    if int(os.environ.get("RANK", "0")) == 0:
        print(f"[bench +{time.time() - T0:6.1f}s] {msg}", file=sys.stderr, flush=True)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="windows per GPU")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--optimizer", default="fused", choices=["fused", "torch"])
    ap.add_argument("--heads", default="linear", choices=["linear", "transformer"],
                    help="prediction heads: 'linear' = BASELINE config 2/3 (default), 'transformer' = config 4 (eval-mode heads)")
    ap.add_argument("--cpu-batch", type=int, default=8, help="windows per step of the CPU baseline sample")
    ap.add_argument("--ar", default="GRU", choices=["GRU", "LSTM", "transformer"],
                    help="context network (--arMode of cpc/train.py; the reference default is LSTM, BASELINE configs use the GRU)")
    ap.add_argument("--preset", default="default", choices=["default", "config5"],
                    help="'config5' = BASELINE config 5 (hiddenEncoder = hiddenGar = 512, 2-level GRU, K = 16, 256 negatives, "
                         "81920-sample windows; use --batch 8); its per-kernel rooflines are not tabulated")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train-py", action="store_true", help="skip the leg that drives the reference's trainStep loop")
    ap.add_argument("--no-torch-gpu", action="store_true", help="skip the informational leg: unmodified reference modules on this GPU")
    ap.add_argument("--launch", default="graph", choices=["graph", "eager"],
                    help="'graph' = the step is captured once as a CUDA graph (cpc_audio_b200.graph.GraphedTrainStep) and "
                         "replayed; 'eager' = ~45 launches per step through the nn.Module surfaces, as cpc/train.py issues them")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm on the host cores (oracle port; the reference itself is Python and is not
# present on the GPU box).  Bounded sample: `cpu_batch` windows per step.
# ---------------------------------------------------------------------------------------------------------------
def cpu_reference_throughput(cpu_batch, steps, warmup):
    """The reference's CPU path on the host cores.  kind = "reference": the UNMODIFIED reference package (baseline/_ref, a
    pip --target install of /root/reference that travels with the snapshot; /root/reference in the authoring container) -
    its own factories build the modules and its own ``trainStep`` (cpc/train.py:64-119) runs the steps, with ``Tensor.cuda``
    neutralised (the one device-specific call of that loop).  kind = "port": the oracle restatement (same torch CPU ops),
    used only when the reference package is absent."""
    import torch
    threads = pick_cpu_threads()
    torch.set_num_threads(threads)
    ref = None
    try:
        from tests import ref_driver as R
        ref = R.reference_or_none()
    except Exception as e:  # noqa: BLE001
        log(f"cpu arm: reference package not importable ({type(e).__name__}: {e}); timing the oracle port")
    log(f"cpu arm: {threads} threads, {cpu_batch} windows/step, kind={'reference' if ref else 'port'}")
    times = []
    t_begin = time.perf_counter()
    if ref is not None:
        args = R.default_args(arMode="GRU", rnnMode="linear")
        model, crit, model_dp, crit_dp, opt = R.build(args, b200=False, seed=0, device="cpu")
        g = torch.Generator().manual_seed(1)
        x = torch.randn(cpu_batch, 1, WINDOW, generator=g) * 0.1
        label = torch.zeros(cpu_batch, dtype=torch.long)
        real_cuda = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self
        try:
            for i in range(warmup + steps):
                t0 = time.perf_counter()
                R.train_steps(model_dp, crit_dp, opt, [(x, label)])
                if i >= warmup:
                    times.append(time.perf_counter() - t0)
                log(f"cpu arm: step {i} {time.perf_counter() - t0:.2f}s")
                if times and time.perf_counter() - t_begin > 60:  # bounded sample
                    break
        finally:
            torch.Tensor.cuda = real_cuda
        kind, what = "reference", "the unmodified reference (cpc/train.py:trainStep over cpc.model / cpc.criterion, torch CPU fp32)"
    else:
        from oracle import cpc_oracle as O
        d = O.Dims(B=cpu_batch, L=WINDOW, H=256, Har=256, K=12, N=128, nLayers=1)
        mp, cp = O.make_params(d, seed=0)
        x, _ = O.make_batch(d, seed=1)
        bi, si = O.make_raw_indices(d, seed=2)
        params = [v.requires_grad_(True) for v in list(cp.values()) + list(mp.values())]
        opt = torch.optim.Adam(params, lr=2e-4, betas=(0.9, 0.999), eps=1e-8)
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.train_step_reference_style(x, mp, cp, bi, si, d)
            opt.step()
            opt.zero_grad()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
            log(f"cpu arm: step {i} {time.perf_counter() - t0:.2f}s")
            if times and time.perf_counter() - t_begin > 60:  # bounded sample
                break
        kind, what = "port", "the oracle port (fwd+bwd+Adam, fp32, torch CPU ops as the reference uses them)"
    steps = len(times)
    times.sort()
    med = times[len(times) // 2]
    return dict(value=cpu_batch * WINDOW / SR / med, unit="audio-s/s", cores=threads, kind=kind,
                sample=f"{steps} steps of {cpu_batch} windows x {WINDOW} samples of {what}, median step {med * 1e3:.0f} ms",
                ms_per_step=med * 1e3)


def usable_cores():
    """Cores this process may really use: affinity mask, capped by the cgroup CPU quota when there is one."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(float(txt[0]) / float(txt[1]) + 0.5)))
            else:
                q = int(txt[0])
                if q > 0:
                    per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                    n = min(n, max(1, int(q / per + 0.5)))
        except Exception:
            pass
    return n


def pick_cpu_threads():
    """Use all the host threads that help: time one small step at a few thread counts, keep the fastest."""
    import torch
    from oracle import cpc_oracle as O
    top = usable_cores()
    cands = sorted({c for c in (top, top // 2, top // 4, 32, 16, 8) if 1 <= c <= top}, reverse=True)
    d = O.Dims(B=2, L=WINDOW, H=256, Har=256, K=12, N=128, nLayers=1)
    mp, cp = O.make_params(d, seed=0)
    for v in list(mp.values()) + list(cp.values()):
        v.requires_grad_(True)
    x, _ = O.make_batch(d, seed=1)
    bi, si = O.make_raw_indices(d, seed=2)
    best, best_t = cands[-1], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        ts = []
        for _ in range(2):
            t0 = time.perf_counter()
            O.train_step_reference_style(x, mp, cp, bi, si, d)
            ts.append(time.perf_counter() - t0)
            if ts[-1] > 20:
                break
        log(f"cpu arm: calibration {c} threads -> {min(ts):.2f}s per 2-window step")
        if min(ts) < best_t:
            best, best_t = c, min(ts)
    return best


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_throughput(a.cpu_batch, max(1, a.steps), max(1, a.warmup))
    line = {"impl": "reference", "metric": "audio-seconds/sec", "value": r["value"], "unit": "audio-s/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "CPC default (H=256, 1-layer GRU, K=12, 128 negatives), 20480-sample windows, "
                                   f"{a.cpu_batch} windows per CPU step", "global_batch": a.cpu_batch, "seq_len": WINDOW},
            "cpu_baseline": {"value": r["value"], "unit": "audio-s/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.th.join(timeout=2)
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(hbm=j["hbm_gbs"], tf_burst=j["bf16_tflops"], tf_sust=j.get("bf16_tflops_sustained", j["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


def algorithmic_work(B, bf16):
    """Per-launch algorithmic bytes / FLOPs of each kernel at default dims (DESIGN.md 'Rooflines')."""
    es = 2 if bf16 else 4
    H, S, W, K, N, L0 = 256, 128, 116, 12, 128, 4096
    w = {}
    # HBM-bound: bytes each launch must move once
    w["conv0_fwd"] = ("hbm", B * (WINDOW * 4 + L0 * H * es))
    w["conv0_bwd_du"] = ("hbm", B * (WINDOW * 4 + 2 * L0 * H * es))
    w["conv0_wgrad"] = ("hbm", B * (WINDOW * 4 + L0 * H * es))
    w["conv0_fwd_mma"] = w["conv0_fwd"]
    w["conv0_bwd_du_mma"] = w["conv0_bwd_du"]
    w["conv0_bwd2_mma"] = ("hbm", B * (WINDOW * 4 + L0 * H * es))   # one read of dy0 (+ the waveform); du0 never leaves the SM
    w["conv0_wgrad_mma"] = w["conv0_wgrad"]
    w["cnorm_relu_fwd"] = ("hbm", sum(B * lo * H * es * 2 for lo in (1024, 512, 256, 128)) + B * 128 * H * 4)
    w["cnorm_relu_bwd"] = ("hbm", sum(B * lo * H * es * 3 for lo in (1024, 512, 256, 128)))
    w["gru_rec_fwd"] = ("hbm", B * S * (3 * H * es + H * 4 + 5 * H * es) + 3 * H * H * 4)
    w["gru_rec_bwd"] = ("hbm", B * S * (2 * H * 4 + 4 * H * es + 6 * H * es) + 3 * H * H * 4)
    w["gru_rec_fwd_mma"] = w["gru_rec_fwd"]
    w["gru_rec_bwd_mma"] = w["gru_rec_bwd"]
    w["score_fwd"] = ("hbm", B * W * (K * H * es + N * 4 + K * (N + 1) * 4) + B * S * H * es)
    w["score_bwd"] = ("hbm", B * W * (2 * K * H * es + N * 4 + K * (N + 1) * 4) + B * S * H * (es + 4))
    w["score_fwd_mma"] = ("hbm", B * W * (K * H * es + N * 4 + 3 * K * 4) + B * S * H * es)
    w["score_bwd_mma"] = ("hbm", B * W * (2 * K * H * es + N * 4 + K * 4) + B * S * H * (es + 4))
    # tensor-bound: FLOPs per step summed over that kernel's launches (reported per launch by dividing)
    conv = [(1024, 8), (512, 4), (256, 4), (128, 4)]
    f_fwd = sum(2.0 * B * lo * H * k * H for lo, k in conv)           # conv1-4 forward (ChannelNorm+ReLU in the epilogue)
    w["gemm_nt_cnorm_tc2"] = ("tensor", f_fwd)
    f_nt = sum(2.0 * B * lo * H * k * H for lo, k in conv)            # dgrad
    f_nt += 2.0 * B * S * 3 * H * H * 2                               # GRU input projection fwd + d(input)
    f_nt += 2.0 * B * W * K * H * H * 2                               # heads fwd + dc
    f_tn = sum(2.0 * B * lo * H * k * H for lo, k in conv) + 2.0 * B * S * 3 * H * H * 2 + 2.0 * B * W * K * H * H
    w["gemm_nt_tc"] = ("tensor", f_nt)
    w["gemm_tn_tc"] = ("tensor", f_tn)
    w["gemm_nt_tc2"] = ("tensor", f_nt)
    w["gemm_tn_tc2"] = ("tensor", f_tn)
    w["gemm_nt_simt"] = ("tensor", f_nt)
    w["gemm_tn_simt"] = ("tensor", f_tn)
    return w


def _trainstep_like(loader, model, crit, opt):
    """The body of cpc/train.py:64-119, statement for statement (used only when the reference package is absent)."""
    import numpy as np
    model.train()
    crit.train()
    logs = {}
    it = 0
    for step, (batchData, label) in enumerate(loader):
        batchData = batchData.cuda(non_blocking=True)
        label = label.cuda(non_blocking=True)
        c_feature, encoded_data, label = model(batchData, label)
        allLosses, allAcc = crit(c_feature, encoded_data, label)
        totLoss = allLosses.sum()
        totLoss.backward()
        opt.step()
        opt.zero_grad()
        if "locLoss_train" not in logs:
            logs["locLoss_train"] = np.zeros(allLosses.size(1))
            logs["locAcc_train"] = np.zeros(allLosses.size(1))
        it += 1
        logs["locLoss_train"] += (allLosses.mean(dim=0)).detach().cpu().numpy()
        logs["locAcc_train"] += (allAcc.mean(dim=0)).cpu().numpy()
    return logs


def _drive_trainstep(a, dev, x_host, label, b200, steps, warm, patch_adam=False, feeder=False, patch_dp=False):
    """Build model / criterion / torch.optim.Adam as cpc/train.py:307-337 + 372-375 does and run `steps` steps of trainStep.
    patch_adam: cpc_audio_b200.patch.install(adam=True) - train.py's torch.optim.Adam(...) call then builds the flat fused
    optimizer.  feeder: the dataLoader is a cpc_audio_b200.feeder.WindowFeeder over an HBM-resident pack (the .cuda() calls
    of the loop become no-ops) instead of a list of pinned host batches."""
    import contextlib
    import io
    import torch
    import cpc_audio_b200 as M
    import cpc_audio_b200.patch as patch
    R = None
    try:
        from tests import ref_driver as R_
        if R_.reference_or_none() is not None:
            R = R_
    except Exception:  # noqa: BLE001
        R = None
    if R is None and not b200:
        return None
    label_host = label.cpu()
    Bsz = x_host.shape[0]
    if patch_adam:
        patch.install_adam()
    if patch_dp:
        patch.install_dataparallel()
    try:
        if R is not None:
            args = R.default_args(arMode="GRU", rnnMode=a.heads)
            os.environ["CPC_B200_DTYPE"] = a.dtype
            model, crit, model_dp, crit_dp, opt = R.build(args, b200=b200, seed=0, device=dev)
            R.use_b200_modules(False)
            ref = R.reference_or_none()
            loop = lambda loader: ref.train.trainStep(loader, model_dp, crit_dp, opt, None, 10 ** 9)  # noqa: E731
            how = "cpc/train.py:trainStep (unmodified, imported from the reference package)"
        else:
            torch.manual_seed(0)
            model = M.CPCModel(M.CPCEncoder(256, "layerNorm", compute_dtype=a.dtype),
                               M.CPCAR(256, 256, False, 1, mode="GRU", reverse=False, compute_dtype=a.dtype)).to(dev)
            crit = M.CPCUnsupersivedCriterion(12, 256, 256, 128, mode=None, rnnMode=a.heads, dropout=False, speakerEmbedding=0,
                                              nSpeakers=0, sizeInputSeq=WINDOW // 160, compute_dtype=a.dtype).to(dev)
            opt = torch.optim.Adam(list(crit.parameters()) + list(model.parameters()), lr=2e-4, betas=(0.9, 0.999), eps=1e-8)
            model_dp = torch.nn.DataParallel(model, device_ids=[dev.index]).to(dev)
            crit_dp = torch.nn.DataParallel(crit, device_ids=[dev.index]).to(dev)
            loop = lambda loader: _trainstep_like(loader, model_dp, crit_dp, opt)  # noqa: E731
            how = "a statement-for-statement copy of the loop of cpc/train.py:64-119 (reference package absent on this box)"
    finally:
        if patch_adam:
            patch.uninstall_adam()
        if patch_dp:
            patch.uninstall_dataparallel()

    if feeder:
        import itertools
        from cpc_audio_b200.feeder import ResidentPack, WindowFeeder
        n_win = Bsz * 8 + 1
        g = torch.Generator().manual_seed(3)
        host_pack = torch.randn(n_win * WINDOW, generator=g) * 0.1
        bounds = [0, host_pack.numel()]
        pack = ResidentPack.from_host(host_pack, bounds, bounds, device=dev)
        feed = WindowFeeder(pack, Bsz, WINDOW, sampling="uniform", random_offset=True)

        def batches(n):  # n batches: as many passes over the resident pack as needed (a pass = 8 batches here)
            return itertools.islice(itertools.chain.from_iterable(iter(feed) for _ in range(n)), n)
    else:
        def batches(n):
            return [(x_host, label_host)] * n
    with contextlib.redirect_stdout(io.StringIO()):
        loop(batches(warm))
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        logs = loop(batches(steps))
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
    if hasattr(opt, "bucket"):
        opt.bucket.detach()
    return dt / steps, how, float(logs["locLoss_train"].mean()), type(opt).__name__


def train_py_leg(a, dev, x_host, label):
    from cpc_audio_b200 import _lib as L
    steps = max(20, a.steps)
    sec = x_host.shape[0] * WINDOW / SR
    out = None
    for name, kw in (("stock", {}), ("flat_adam", {"patch_adam": True}), ("fast", {"patch_adam": True, "patch_dp": True}),
                     ("fast_feeder", {"patch_adam": True, "patch_dp": True, "feeder": True})):
        n0 = L.lib().cpcb200_launch_count()
        per_step, how, loss, opt_name = _drive_trainstep(a, dev, x_host, label, True, steps, 5, **kw)
        launches = (L.lib().cpcb200_launch_count() - n0) / (steps + 5)
        log(f"train.py-driven [{name}]: {per_step * 1e3:.3f} ms/step ({launches:.0f} library launches/step, optimizer {opt_name})")
        ent = {"value": sec / per_step, "unit": "audio-s/s", "ms_per_step": per_step * 1e3, "kernel_launches_per_step": launches,
               "optimizer_class": opt_name, "mean_loss": loss}
        if out is None:
            out = dict(ent)
            out.update({"steps": steps, "h2d_bytes_per_step": x_host.numel() * 4 + label.numel() * 8, "d2h_bytes_per_step": 2 * 12 * 4,
                        "driver": how, "optimizer": "torch.optim.Adam (as cpc/train.py:335 builds it)",
                        "timing": "host wall clock around the call, device synchronised on both sides (the loop itself reads the loss every step)",
                        "variants": {}})
        else:
            ent["what"] = {"flat_adam": "cpc_audio_b200.patch.install(adam=True): train.py's torch.optim.Adam(...) call builds the flat fused "
                                        "optimizer (one kernel per step); host batches as above",
                           "fast": "python -m cpc_audio_b200.patch --fast train.py: flat_adam + torch.nn.DataParallel(device_ids=[0]) "
                                   "becomes a true pass-through (no scatter / gather for one device); host batches as above",
                           "fast_feeder": "the same + the dataLoader is a WindowFeeder over an HBM-resident pack (SURVEY 8f N2): no "
                                          "per-step host->device copy, batches cut by cpcb200_gather_windows"}[name]
            out["variants"][name] = ent
    return out


def torch_gpu_leg(a, dev, x_host, label):
    """Informational (SURVEY 8(d) optional row): the UNMODIFIED reference modules on this same B200 through torch / cuDNN /
    cuBLAS in fp32 (the reference's own GPU path), driven by its own trainStep.  Not the baseline the contract names (that
    is the CPU arm) - the honest same-box GPU comparator."""
    import torch
    steps = 5
    r = _drive_trainstep(a, dev, x_host, label, False, steps, 2)
    if r is None:
        return {"unavailable": "reference package not present on this box (baseline/_ref)"}
    per_step, how, loss, _ = r
    torch.cuda.empty_cache()
    sec = x_host.shape[0] * WINDOW / SR
    log(f"torch-on-GPU reference: {per_step * 1e3:.1f} ms/step")
    return {"value": sec / per_step, "unit": "audio-s/s", "ms_per_step": per_step * 1e3, "steps": steps, "dtype": "f32 (TF32 convs: torch default)",
            "driver": how, "mean_loss": loss}


def run_ours(a):
    import torch
    import torch.distributed as dist
    import cpc_audio_b200 as M
    from cpc_audio_b200 import _lib as L
    from cpc_audio_b200.optim import FlatAdam, GradBucket, PeerAdam

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = L.lib()
    B = a.batch

    global WINDOW
    HID, NLEV, KP, NNEG = 256, 1, 12, 128
    if a.preset == "config5":
        HID, NLEV, KP, NNEG, WINDOW = 512, 2, 16, 256, 81920
        a.no_train_py = a.no_torch_gpu = a.no_cpu_baseline = True
    torch.manual_seed(0)  # identical initial parameters on every rank (replicas)
    arnet = (M.buildTransformerAR(HID, 1, WINDOW // 160, False, compute_dtype=a.dtype) if a.ar == "transformer" else
             M.CPCAR(HID, HID, False, NLEV, mode=a.ar, reverse=False, compute_dtype=a.dtype))
    model = M.CPCModel(M.CPCEncoder(HID, "layerNorm", compute_dtype=a.dtype), arnet).to(dev)
    if a.ar != "GRU":
        a.no_train_py = a.no_torch_gpu = True
    crit = M.CPCUnsupersivedCriterion(KP, HID, HID, NNEG, mode=None, rnnMode=a.heads, dropout=False, speakerEmbedding=0,
                                      nSpeakers=0, sizeInputSeq=WINDOW // 160, compute_dtype=a.dtype).to(dev)
    model.train()
    crit.train()  # cpc/train.py:71-72: the transformer heads then apply the reference's dropout 0.1 (masks from torch's generator)
    params = list(crit.parameters()) + list(model.parameters())  # cpc/train.py:332 order
    use_graph = a.launch == "graph"
    fused_ar = False  # all-reduce + Adam + zero_grad as ONE kernel over peer memory (PeerAdam) instead of NCCL + Adam
    if a.optimizer == "fused":
        opt = None
        force_peer = world == 1 and os.environ.get("CPC_B200_FORCE_PEER", "0") == "1"  # debug: PeerAdam on one GPU (gloo group)
        if force_peer:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29533")
            dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
        if (world > 1 or force_peer) and os.environ.get("CPC_B200_FUSED_AR", "1") != "0":
            try:
                enc = model.gEncoder
                # early-range overlap is OFF by default: measured on B200 (profiles/r2_multi_gpu.md) the side-stream kernel slows
                # the dgrad GEMM / conv0 backward it runs beside by more than the exchange it hides
                overlap = os.environ.get("CPC_B200_PEER_OVERLAP", "0") != "0"
                opt = PeerAdam(params, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, fuse_zero_grad=True, overlap=overlap,
                               late_params=[enc.conv0.weight, enc.conv0.bias, enc.batchNorm0.weight, enc.batchNorm0.bias])
                fused_ar = True
            except Exception as e:  # noqa: BLE001 - symmetric memory unavailable: NCCL all-reduce + fused Adam instead
                log(f"peer-memory all-reduce unavailable ({type(e).__name__}: {e}); using NCCL")
                opt = None
        if opt is None:
            opt = FlatAdam(params, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, capturable=use_graph, fuse_zero_grad=use_graph)
        bucket = opt.bucket
    else:
        opt = torch.optim.Adam(params, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, capturable=use_graph)
        bucket = GradBucket(params)

    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    torch.cuda.manual_seed(4321 + rank)  # negative-sample draws differ per rank, like DataParallel replicas
    x_dev = torch.randn(B, 1, WINDOW, device=dev, generator=gen) * 0.1
    label = torch.zeros(B, dtype=torch.long, device=dev)
    x_host = x_dev.cpu().pin_memory()
    loss_host = torch.empty(1, KP).pin_memory()

    if world > 1 and not fused_ar and os.environ.get("CPC_B200_AR_OVERLAP", "0") != "0":  # measured slower: see DESIGN.md 5
        enc = model.gEncoder  # all-reduce everything but conv0's gradients while the backward pass is still running
        bucket.setup_overlap([enc.conv0.weight, enc.conv0.bias, enc.batchNorm0.weight, enc.batchNorm0.bias])

    def step_eager(x):
        c, z, _ = model(x, label)
        losses, acc = crit(c, z, label)
        if world > 1 and not fused_ar:
            bucket.arm_overlap()
        elif fused_ar:
            opt.arm_overlap()
        losses.sum().backward()
        if world > 1 and not fused_ar:
            bucket.allreduce()
        opt.step()
        opt.zero_grad()
        return losses

    gstep, launch_mode, launches_per_replay = None, "eager", None
    if use_graph:
        try:
            from cpc_audio_b200.graph import GraphedTrainStep
            n_before = lib.cpcb200_launch_count()
            nccl = world > 1 and not fused_ar
            gstep = GraphedTrainStep(model, crit, opt, x_dev, label, allreduce=bucket.allreduce if nccl else None, warmup=3,
                                     before_backward=bucket.arm_overlap if nccl else (opt.arm_overlap if fused_ar else None))
            launches_per_replay = (lib.cpcb200_launch_count() - n_before) // 4  # 3 warm-up steps + the captured one
            launch_mode = "cuda-graph replay (GraphedTrainStep)"
        except Exception as e:  # noqa: BLE001 - report and run the eager loop instead (same kernels, same work)
            log(f"CUDA-graph capture failed ({type(e).__name__}: {e}); running eager launches")
            gstep = None
            torch.cuda.synchronize(dev)

    def step(x):
        if gstep is not None:
            return gstep(x)[0]
        return step_eager(x)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def timed(fn, n):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        sync_all()
        return ms.item()

    W_ = max(3, a.warmup)
    log("modules built; warm-up")
    # clocks / throttle reasons are sampled (nvidia-smi, 100 ms period) from the warm-up to the end of the end-to-end pass:
    # the GPU is under the same load throughout, and a 20-step timed region alone (30 ms) would hold a single sample
    clocks = ClockSampler(local) if rank == 0 else None
    if clocks:
        clocks.start()
    for _ in range(W_):
        step(x_dev)
    torch.cuda.synchronize(dev)
    log("warm-up done; timed region")
    n0 = lib.cpcb200_launch_count()
    ms = timed(lambda: step(x_dev), a.steps)
    launches = lib.cpcb200_launch_count() - n0 if gstep is None else launches_per_replay * a.steps
    log(f"timed: {ms / a.steps:.3f} ms/step; e2e pass")

    # End to end: the batch lives in pinned host memory.  As in a training loop with a prefetching loader, the
    # copy of step i+1's batch runs on a side stream while step i computes; every step still pays one full H2D of its own
    # inputs and one D2H + host synchronisation for its loss (train.py:98 reads it every step).
    copy_stream = torch.cuda.Stream(device=dev)
    x_bufs = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    state = {"i": 0}

    def prefetch(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])      # the step that last read this buffer is done
            x_bufs[slot].copy_(x_host, non_blocking=True)
            ready[slot].record(copy_stream)

    for sl in range(2):
        consumed[sl].record(torch.cuda.current_stream(dev))
    prefetch(0)

    def e2e_step():
        slot = state["i"] & 1
        state["i"] += 1
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(ready[slot])
        losses = step(x_bufs[slot])
        consumed[slot].record(cur)
        loss_host.copy_(losses.detach(), non_blocking=True)
        prefetch(slot ^ 1)                               # next step's batch: copied while this step computes
        cur.synchronize()  # the loss is read on the host every step (train.py:98)

    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, a.steps)
    # keep the same load up for ~0.5 s more (a few sampler periods); ms_e2e is the max over ranks, so every rank runs the
    # same number of extra steps (they contain the gradient exchange)
    for _ in range(min(2000, int(0.5 / max(ms_e2e / a.steps * 1e-3, 1e-5)))):
        e2e_step()
    torch.cuda.synchronize(dev)
    clk = clocks.stop() if clocks else None

    # The drop-in path itself (north_star: "cpc/train.py is unchanged"): the reference's OWN trainStep loop (cpc/train.py:64-119)
    # - .cuda() of a pinned host batch, forward, backward, torch.optim.Adam.step/zero_grad, loss read back every step - over
    # freshly built B200 modules behind DataParallel(device_ids=[0]), eager launches, no graph, no fused optimizer.
    train_py = None
    if world == 1 and not a.no_train_py:
        try:
            train_py = train_py_leg(a, dev, x_host, label)
        except Exception as e:  # noqa: BLE001
            log(f"train.py-driven leg failed: {type(e).__name__}: {e}")
            train_py = {"error": f"{type(e).__name__}: {e}"}
    torch_gpu = None
    if world == 1 and not a.no_torch_gpu and a.heads == "linear":
        try:
            torch_gpu = torch_gpu_leg(a, dev, x_host, label)
        except Exception as e:  # noqa: BLE001
            log(f"torch-on-GPU reference leg failed: {type(e).__name__}: {e}")
            torch_gpu = {"error": f"{type(e).__name__}: {e}"}

    # per-kernel timing pass (CUDA events on the launching stream, same workload, after the timed region)
    roof = None
    log(f"e2e: {ms_e2e / a.steps:.3f} ms/step; per-kernel pass")
    if rank == 0:
        lib.cpcb200_prof_enable(1)
    os.environ["CPC_B200_OVERLAP"] = "0"   # single stream: the per-kernel events are recorded on the launching stream
    for _ in range(min(a.steps, 10)):  # every rank runs these steps (they contain the all-reduce); eager launches: the
        step_eager(x_dev)              # event after every kernel serialises them (no programmatic overlap, no graph)
    torch.cuda.synchronize(dev)
    if rank == 0:
        buf = ctypes.create_string_buffer(1 << 16)
        L.check(lib.cpcb200_prof_report(buf, len(buf)), "prof_report")
        lib.cpcb200_prof_enable(0)
        nst = min(a.steps, 10)
        rows = {}
        for line in buf.value.decode().splitlines():
            name, cnt, tot = line.split()
            rows[name] = (int(cnt), float(tot))
        pk = peaks()
        work = algorithmic_work(B, a.dtype == "bf16") if a.preset == "default" else {}
        per_step = {k: v[1] / nst for k, v in rows.items()}
        # the dominant kernel = the library kernel with the largest time per step among those with a stated roofline
        top = max((k for k in per_step if k in work), key=per_step.get, default=max(per_step, key=per_step.get))
        kernels = {}
        for k, (cnt, tot) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
            ent = {"launches_per_step": cnt / nst, "ms_per_step": round(tot / nst, 4)}
            if k in work:
                kind, amount = work[k]
                if kind == "hbm":  # `amount` = bytes of all launches of this kernel in one step
                    ach = amount / (tot / nst * 1e-3) / 1e9
                    ent.update(bound="hbm", achieved_gbs=round(ach, 1), frac=round(ach / pk["hbm"], 4))
                else:
                    ach = amount / (tot / nst * 1e-3) / 1e12
                    ent.update(bound="tensor", achieved_tflops=round(ach, 1), frac=round(ach / pk["tf_sust"], 4))
            if k == "score_bwd_mma":
                # the kernel's real bound: fp32 scatter-adds of the negatives' / positives' gradient rows into the L2-resident
                # dz (B*W*(N+K) rows of H floats per launch); ceiling measured by tools/redbench.cu on this GPU type
                # (profiles/r1e_redbench_cpu_bound.txt: 5.4 TB/s for TMA bulk reductions and for red.global.v4 alike)
                red = B * 116 * (128 + 12) * 256 * 4 / (tot / nst * 1e-3) / 1e9
                ent.update(l2_scatter_add_gbs=round(red, 1), l2_scatter_add_ceiling_gbs=L2_RED_CEILING_GBS,
                           frac_of_l2_scatter_add_ceiling=round(red / L2_RED_CEILING_GBS, 4))
            kernels[k] = ent
        t = kernels[top]
        if t.get("bound") == "hbm":
            roof = {"kernel": top, "bound": "hbm", "achieved": t["achieved_gbs"], "peak": pk["hbm"], "unit": "GB/s",
                    "frac": t["frac"], "traffic": None, "peak_source": pk["src"] + " (MEASURED_PEAKS.json hbm_gbs)"}
        elif t.get("bound") == "tensor":
            roof = {"kernel": top, "bound": "tensor", "achieved": t["achieved_tflops"], "peak": pk["tf_sust"], "unit": "TFLOP/s",
                    "frac": t["frac"], "traffic": None, "peak_source": pk["src"] + " (bf16_tflops_sustained: kernel timed inside a long step)"}
        else:
            roof = {"kernel": top, "bound": "hbm", "achieved": None, "peak": pk["hbm"], "unit": "GB/s", "frac": None, "traffic": None}
        # DRAM traffic per launch of that kernel from the committed `ncu --set full` capture (profiles/), not measured live
        try:
            tj = json.load(open(os.path.join(REPO, "profiles", "r2f_traffic.json")))
            roof["traffic"] = tj["kernels"][top]["dram_bytes_per_launch"]
            roof["traffic_source"] = "profiles/r2f_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, round 2)"
            for k, ent in kernels.items():
                if k in tj["kernels"]:
                    ent["ncu_dram_bytes_per_launch"] = tj["kernels"][k]["dram_bytes_per_launch"]
        except Exception:
            pass
        if "l2_scatter_add_gbs" in kernels[top]:  # the dominant kernel's real limiter, next to the contract's HBM column
            roof["l2_scatter_add"] = {"achieved": kernels[top]["l2_scatter_add_gbs"], "ceiling": L2_RED_CEILING_GBS, "unit": "GB/s",
                                      "frac": kernels[top]["frac_of_l2_scatter_add_ceiling"],
                                      "note": "fp32 row reductions into the L2-resident dz; ceiling measured with tools/redbench.cu"}
        roof["kernels"] = kernels
        roof["kernel_ms_per_step_total"] = round(sum(per_step.values()), 4)

    if world > 1 and fused_ar and os.environ.get("CPC_B200_PEER_STAMPS", "0") == "1":
        for _ in range(3):
            step(x_dev)
        torch.cuda.synchronize(dev)
        sg = opt._sig.cpu()
        rel = sg[40:45].tolist()
        ab = sg[48:56].view(torch.int64).tolist()
        print(f"[stamps rank {rank}] early kernel: start 0, peers arrived +{(ab[1] - ab[0]) / 1e3:.1f} us, done +{(ab[2] - ab[0]) / 1e3:.1f} us; "
              f"step kernel starts +{(ab[3] - ab[0]) / 1e3:.1f} us after the early kernel started; inside the step kernel (ns): "
              f"barrier1 {rel[0]}, reduced {rel[1]}, grid sync {rel[2]}, barrier2 {rel[3]}, adam done {rel[4]}", file=sys.stderr, flush=True)
    if world > 1:
        dist.barrier()
    if rank == 0:
        cpu = None
        if world == 1 and not a.no_cpu_baseline:
            r = cpu_reference_throughput(a.cpu_batch, 3, 1)
            cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        sec = B * world * WINDOW / SR
        line = {"metric": "audio-seconds/sec", "value": sec / (ms / a.steps * 1e-3), "unit": "audio-s/s", "n_gpus": world,
                "steps": a.steps, "warmup": W_, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
                "config": {"workload": ("BASELINE config 5: hiddenEncoder=512, hiddenGar=512, nLevelsGRU=2, K=16, 256 negatives " if a.preset == "config5" else
                                        "BASELINE config 2: CPC default (hiddenEncoder=256, 1-layer GRU, K=12, 128 negatives) "
                                        if a.heads == "linear" else
                                        "BASELINE config 4: --rnnMode transformer prediction heads (train mode, dropout 0.1), GRU context net, K=12, 128 negatives ")
                                       + (f"[context network: {a.ar}] " if a.ar != "GRU" else "")
                                       + f"batch={B}/GPU seq={WINDOW}, white-noise 16 kHz windows, random-init weights",
                           "global_batch": B * world, "seq_len": WINDOW, "parallelism": f"dp{world}", "optimizer": a.optimizer, "launch": launch_mode,
                           "gradient_exchange": ("none (1 GPU)" if world == 1 else
                                                 ("peer-memory all-reduce of the early gradients on a side stream during the backward tail "
                                                  "(cpcb200_peer_reduce_range) + one kernel: late-gradient exchange + Adam + zero_grad "
                                                  if getattr(opt, "overlap", False) else "one kernel: peer-memory all-reduce + Adam + zero_grad ")
                                                 + "(cpcb200_allreduce_adam_step"
                                                 + (", NVSwitch multimem reduction)" if getattr(opt, "multicast", False) else ", peer loads/stores)")
                                                 if fused_ar else "NCCL all-reduce + fused Adam"),
                           "l2": "no explicit flush: one step streams > 1 GB of activations (> 126 MB L2) between reuses"},
                "e2e": {"value": sec / (ms_e2e / a.steps * 1e-3), "unit": "audio-s/s", "h2d_bytes_per_step": x_host.numel() * 4,
                        "d2h_bytes_per_step": loss_host.numel() * 4, "ms_per_step": ms_e2e / a.steps},
                "e2e_train_py": train_py, "torch_gpu_reference": torch_gpu,
                "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        # a captured graph holds NCCL work: release it before the communicator goes away, and do not let a slow
        # communicator teardown hold the job (every rank has finished and printed by now)
        if gstep is not None:
            gstep.graph.reset()
        torch.cuda.synchronize(dev)
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)
