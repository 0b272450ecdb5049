"""GPU parity tests (run with `-m gpu` on the B200 box): CUDA path through the C ABI vs the CPU oracle and the
committed reference fixtures.  Tolerances (stated per dtype, see DESIGN.md "Parity"):

  f32 path : z, c max-rel <= 1e-4 ; per-step loss |d| <= 1e-4 (+1e-5 rel) ; acc equal except near-tie rows ;
             every parameter gradient rel-L2 <= 5e-4.
  bf16 path: z, c rel-L2 <= 1.5e-2 (measured 4e-3 .. 8e-3) ; loss |d| <= 3e-3 (reference init) / 1% + 1e-2 (x30-scaled heads) ;
             acc within 1% abs at default widths (0.5% at the benchmarked shape) ; gradient cosine per parameter tensor >=
             tests/helpers.py:bf16_cos_floor - 0.999 except the named, measured exceptions (encoder tensors of tiny
             batches, conv0.weight, transformer-layer parameters).  Measured values: profiles/parity_r2.json.
  indices  : bit-exact.
"""
import math

import os

import numpy as np
import pytest
import torch

from oracle import cpc_oracle as O
from tests import helpers as Hh

pytestmark = pytest.mark.gpu


def _cos(a, b):
    a, b = a.detach().double().flatten().cpu(), b.detach().double().flatten().cpu()
    return (a @ b / (a.norm() * b.norm() + 1e-30)).item()


@pytest.mark.parametrize("name", Hh.CASES)
def test_ext_indices_bit_exact(name, built_lib):
    import cpc_audio_b200 as M
    g, d, mp, cp, x, label, bi, si = Hh.load_case(name)
    crit = M.CPCUnsupersivedCriterion(d.K, d.Har, d.H, d.N, rnnMode="linear", sizeInputSeq=d.S).cuda()
    dims = (d.B, d.S, d.H, d.Har, d.K, d.N, 0)
    ext = crit.extIndices(bi.cuda(), si.cuda(), dims).cpu().numpy()
    assert ext.dtype == np.int32 and np.array_equal(ext, g["ext_idx"])
    assert np.array_equal(ext, O.ext_indices_np(bi.numpy(), si.numpy(), d.B, d.N, d.W, d.S))


@pytest.mark.parametrize("name", Hh.CASES)
def test_parity_f32(name, built_lib):
    g, d, mp, cp, x, label, bi, si = Hh.load_case(name)
    ref = Hh.oracle_run(d, mp, cp, x, bi, si)
    model, crit = Hh.build_modules(d, mp, cp, "f32")
    out = Hh.run_modules(model, crit, x, label, bi, si)
    assert out["z"].shape == (d.B, d.S, d.H) and out["c"].shape == (d.B, d.S, d.Har)
    assert out["losses"].shape == (1, d.K) and out["acc"].shape == (1, d.K)
    assert Hh.max_rel(out["z"], ref["z"]) <= 1e-4
    assert Hh.max_rel(out["c"], ref["c"]) <= 1e-4
    # against the committed reference fixture
    np.testing.assert_allclose(Hh.subsample(out["z"]), g["z_sub"], rtol=0, atol=1e-4 * np.abs(g["z_sub"]).max())
    np.testing.assert_allclose(out["losses"].cpu().numpy(), g["losses"], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(out["losses"].cpu().numpy(), ref["losses"].numpy(), rtol=1e-5, atol=1e-4)
    tol = Hh.acc_tolerance(ref["logits"], d)
    assert ((out["acc"].cpu() - ref["acc"]).abs() <= tol).all()
    assert ((out["acc"].cpu() - torch.from_numpy(g["acc"])).abs() <= tol).all()
    rep, wc, wr = Hh.grad_report(out["grads"], ref["grads"])
    Hh.record(f"parity:{name}:f32", z_maxrel=Hh.max_rel(out["z"], ref["z"]), c_maxrel=Hh.max_rel(out["c"], ref["c"]),
              dloss=(out["losses"].cpu() - ref["losses"]).abs().max().item(), worst_cos=[wc[0], wc[1][0]], worst_rel=[wr[0], wr[1][1]])
    for k, gr in ref["grads"].items():
        e = Hh.rel_err(out["grads"][k], gr)
        assert e <= 5e-4, (k, e)
        sub = g[f"gsub.{k}"]
        assert np.abs(Hh.subsample(out["grads"][k], 512) - sub).max() <= 1e-3 * (np.abs(sub).max() + 1e-12), k


@pytest.mark.parametrize("name", Hh.CASES)
def test_parity_bf16(name, built_lib):
    g, d, mp, cp, x, label, bi, si = Hh.load_case(name)
    ref = Hh.oracle_run(d, mp, cp, x, bi, si)
    model, crit = Hh.build_modules(d, mp, cp, "bf16")
    out = Hh.run_modules(model, crit, x, label, bi, si)
    assert Hh.rel_err(out["z"], ref["z"]) <= 1.5e-2     # measured 5.0e-3 .. 6.5e-3
    assert Hh.rel_err(out["c"], ref["c"]) <= 1.5e-2     # measured 4.4e-3 .. 7.8e-3
    scaled = float(g["pred_scale"]) != 1.0
    dl = (out["losses"].cpu() - ref["losses"]).abs()
    if scaled:
        assert (dl <= 0.01 * ref["losses"].abs() + 1e-2).all(), dl
    else:
        assert (dl <= 3e-3).all(), dl
    assert ((out["acc"].cpu() - ref["acc"]).abs() <= (0.01 if d.H >= 256 else 0.03) + Hh.acc_tolerance(ref["logits"], d)).all()
    rep, wc, wr = Hh.grad_report(out["grads"], ref["grads"])
    Hh.record(f"parity:{name}:bf16", z_rel=Hh.rel_err(out["z"], ref["z"]), c_rel=Hh.rel_err(out["c"], ref["c"]), dloss=dl.max().item(),
              dacc=(out["acc"].cpu() - ref["acc"]).abs().max().item(), worst_cos=[wc[0], wc[1][0]], worst_rel=[wr[0], wr[1][1]],
              grads={k: [round(v[0], 6), float(f"{v[1]:.3e}")] for k, v in rep.items()})
    for k, gr in ref["grads"].items():
        cs = _cos(out["grads"][k], gr)
        assert cs >= Hh.bf16_cos_floor(k, d), (k, cs)


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_gemm_building_blocks(dtype, built_lib):
    from cpc_audio_b200 import _lib as L
    lib = L.lib()
    code = L.F32 if dtype == "f32" else L.BF16
    tdt = torch.float32 if dtype == "f32" else torch.bfloat16
    gen = torch.Generator(device="cuda").manual_seed(0)
    for (M_, N_, K_) in [(200, 64, 64), (1024, 256, 2048), (7424, 3072, 256), (8192, 768, 256), (333, 128, 512)]:
        A = torch.randn(M_, K_, device="cuda", generator=gen).to(tdt)
        Bm = torch.randn(N_, K_, device="cuda", generator=gen).to(tdt)
        bias = torch.randn(N_, device="cuda", generator=gen)
        C = torch.empty(M_, N_, device="cuda")
        L.check(lib.cpcb200_test_gemm_nt(code, M_, N_, K_, L.ptr(A), L.ptr(Bm), L.ptr(bias), L.ptr(C), L.stream_ptr(A.device)), "nt")
        ref = A.double() @ Bm.double().t() + bias.double()
        assert Hh.max_rel(C, ref) <= 2e-5, (M_, N_, K_, Hh.max_rel(C, ref))
    for (M_, N1, N2) in [(500, 64, 64), (65536, 256, 2048), (8192, 768, 256), (7424, 3072, 256), (1000, 128, 192)]:
        A = torch.randn(M_, N1, device="cuda", generator=gen).to(tdt)
        Bm = torch.randn(M_, N2, device="cuda", generator=gen).to(tdt)
        C = torch.zeros(N1, N2, device="cuda")
        L.check(lib.cpcb200_test_gemm_tn(code, M_, N1, N2, L.ptr(A), L.ptr(Bm), L.ptr(C), L.stream_ptr(A.device)), "tn")
        ref = A.double().t() @ Bm.double()
        assert Hh.max_rel(C, ref) <= 5e-5, (M_, N1, N2, Hh.max_rel(C, ref))


def test_gru_carried_hidden_state(built_lib):
    """keepHidden (cpc/model.py:194-198): the second chunk starts from the first chunk's final state."""
    import cpc_audio_b200 as M
    d = O.Dims(B=3, L=160 * 12, H=64, Har=128, nLayers=2)
    mp, _ = O.make_params(d, seed=11)
    z = torch.randn(d.B, 2 * d.S, d.H, generator=torch.Generator().manual_seed(3))
    ref, _ = O.gru_forward(z, mp, 2)
    ar = M.CPCAR(d.H, d.Har, True, 2, compute_dtype="f32")
    ar.load_state_dict({k.replace("gAR.", ""): v for k, v in mp.items() if k.startswith("gAR.")})
    ar = ar.cuda()
    zc = z.cuda().requires_grad_(True)
    c1 = ar(zc[:, :d.S])
    assert ar.hidden is not None and ar.hidden.shape == (2, d.B, d.Har) and not ar.hidden.requires_grad
    h_mid = ar.hidden.clone()
    c2 = ar(zc[:, d.S:])
    assert Hh.max_rel(torch.cat([c1, c2], 1), ref) <= 1e-4
    # gradient through the second chunk with a carried (detached) state
    h0 = h_mid.cpu()
    assert Hh.max_rel(h0, O.gru_forward(z[:, :d.S], mp, 2)[1]) <= 1e-4
    zr = z[:, d.S:].clone().requires_grad_(True)
    mpr = {k: v.clone().requires_grad_(True) for k, v in mp.items()}
    cr, _ = O.gru_forward(zr, mpr, 2, h0=O.gru_forward(z[:, :d.S], mp, 2)[1])
    wgt = torch.randn(cr.shape, generator=torch.Generator().manual_seed(4))
    (cr * wgt).sum().backward()
    ar.zero_grad()
    ar.hidden = h0.cuda()
    c2b = ar(zc[:, d.S:])
    (c2b * wgt.cuda()).sum().backward()
    assert Hh.rel_err(zc.grad[:, d.S:], zr.grad) <= 5e-4
    assert Hh.rel_err(ar.baseNet.weight_hh_l1.grad, mpr["gAR.baseNet.weight_hh_l1"].grad) <= 5e-4
    assert Hh.rel_err(ar.baseNet.weight_hh_l0.grad, mpr["gAR.baseNet.weight_hh_l0"].grad) <= 5e-4


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_full_size_properties(dtype, built_lib):
    """BASELINE config 2 shape (B=64, default dims): size-independent properties instead of an oracle run."""
    d = O.Dims(B=64, L=20480, H=256, Har=256, K=12, N=128, nLayers=1)
    mp, cp = O.make_params(d, seed=7, pred_scale=30.0)
    x, label = O.make_batch(d, seed=77)
    bi, si = O.make_raw_indices(d, seed=777)
    model, crit = Hh.build_modules(d, mp, cp, dtype)
    out = Hh.run_modules(model, crit, x, label, bi, si)
    assert torch.isfinite(out["losses"]).all() and all(torch.isfinite(g).all() for g in out["grads"].values())
    # (1) windows are independent through the model: permuting the batch permutes z and c
    perm = torch.randperm(d.B, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        c2, z2, _ = model(x[perm].cuda(), label.cuda())
    assert Hh.max_rel(z2, out["z"][perm.cuda()]) <= 1e-6 and Hh.max_rel(c2, out["c"][perm.cuda()]) <= 1e-6
    # (2) sub-batch consistency against the CPU oracle on 4 of the 64 windows (encoder + GRU are per-window)
    zo = O.encoder_forward(x[:4], mp).permute(0, 2, 1)
    co, _ = O.gru_forward(zo, mp, 1)
    tol = 1e-4 if dtype == "f32" else 3e-2
    assert Hh.rel_err(out["z"][:4], zo) <= tol and Hh.rel_err(out["c"][:4], co) <= tol
    # (3) zero heads: every logit is 0 -> loss = ln(N+1) exactly, accuracy 1 (class 0 wins ties, as torch.max)
    for p in crit.parameters():
        p.data.zero_()
    o0 = Hh.run_modules(model, crit, x, label, bi, si)
    assert (o0["losses"] - math.log(d.N + 1)).abs().max() <= 1e-5 and (o0["acc"] == 1).all()
    # (4) scoring against the oracle criterion fed with the GPU's own c, z (full size, einsum form)
    crit.load_state_dict(cp)
    lo, ao, lg = O.criterion_forward(out["c"].cpu(), out["z"].cpu(), cp, bi, si, d.K, d.N, materialize=False)
    dl = (out["losses"].cpu() - lo).abs()
    assert (dl <= (1e-4 if dtype == "f32" else 0.02 * lo.abs())).all(), dl
    assert ((out["acc"].cpu() - ao).abs() <= (0.0 if dtype == "f32" else 0.01) + Hh.acc_tolerance(lg, d)).all()


def test_training_reduces_loss_and_is_reproducible(built_lib):
    d = O.Dims(B=4, L=20480, H=256, Har=256, K=12, N=128, nLayers=1)
    mp, cp = O.make_params(d, seed=21)
    x, label = O.make_batch(d, seed=22)

    def run():
        torch.manual_seed(0)
        model, crit = Hh.build_modules(d, mp, cp, "bf16")
        opt = torch.optim.Adam(list(crit.parameters()) + list(model.parameters()), lr=2e-4)
        hist = []
        for _ in range(8):
            c, z, _ = model(x.cuda(), label.cuda())
            losses, acc = crit(c, z, label.cuda())
            losses.sum().backward()
            opt.step()
            opt.zero_grad()
            hist.append(losses.mean().item())
        return hist

    h1 = run()
    assert h1[-1] < h1[0], h1
    h2 = run()
    assert np.allclose(h1, h2, rtol=1e-3), (h1, h2)


def test_fused_adam_matches_torch_adam(built_lib):
    """cpcb200_adam_step vs torch.optim.Adam on identical, deterministic gradients (cpc/train.py:335-337)."""
    from cpc_audio_b200.optim import FlatAdam
    gen = torch.Generator(device="cuda").manual_seed(0)
    shapes = [(256, 256, 8), (768,), (1, 256, 1), (12, 256, 256)]
    p_t = [torch.nn.Parameter(torch.randn(s, device="cuda", generator=gen)) for s in shapes]
    p_f = [torch.nn.Parameter(p.detach().clone()) for p in p_t]
    o_t = torch.optim.Adam(p_t, lr=2e-4, betas=(0.9, 0.999), eps=1e-8)
    o_f = FlatAdam(p_f, lr=2e-4, betas=(0.9, 0.999), eps=1e-8)
    for it in range(6):
        for a, b in zip(p_t, p_f):
            g = torch.randn(a.shape, device="cuda", generator=gen) * (10.0 ** (it - 4))
            a.grad = g.clone()
            b.grad.copy_(g)
        o_t.step(); o_f.step()
        o_t.zero_grad(); o_f.zero_grad()
        assert all(b.grad.eq(0).all() for b in p_f)
    o_f.bucket.detach()
    for a, b in zip(p_t, p_f):
        assert (a - b).abs().max().item() <= 2e-6 * max(1.0, a.abs().max().item())


def test_capturable_adam_matches_torch_adam(built_lib):
    """cpcb200_adam_step_dev (device-side step count, gradient clearing fused) vs torch.optim.Adam."""
    from cpc_audio_b200.optim import FlatAdam
    gen = torch.Generator(device="cuda").manual_seed(1)
    shapes = [(256, 256, 4), (768,), (12, 256, 256)]
    p_t = [torch.nn.Parameter(torch.randn(s, device="cuda", generator=gen)) for s in shapes]
    p_f = [torch.nn.Parameter(p.detach().clone()) for p in p_t]
    o_t = torch.optim.Adam(p_t, lr=2e-4, betas=(0.9, 0.999), eps=1e-8)
    o_f = FlatAdam(p_f, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, capturable=True, fuse_zero_grad=True)
    for it in range(6):
        for a, b in zip(p_t, p_f):
            g = torch.randn(a.shape, device="cuda", generator=gen) * (10.0 ** (it - 4))
            a.grad = g.clone()
            b.grad.copy_(g)
        o_t.step(); o_f.step()
        assert all(b.grad.eq(0).all() for b in p_f)  # cleared by the step itself
        o_t.zero_grad(); o_f.zero_grad()
    assert o_f.steps == 6
    o_f.bucket.detach()
    for a, b in zip(p_t, p_f):
        assert (a - b).abs().max().item() <= 2e-6 * max(1.0, a.abs().max().item())


def test_graphed_step_matches_eager_steps(built_lib):
    """GraphedTrainStep (whole step captured as one CUDA graph, replayed) must walk the same trajectory as the eager
    loop of cpc/train.py:83-91: same losses step by step, same parameters at the end (bf16 path, B = 4)."""
    import cpc_audio_b200 as M
    from cpc_audio_b200.graph import GraphedTrainStep
    from cpc_audio_b200.optim import FlatAdam
    dev = torch.device("cuda:0")

    def build():
        torch.manual_seed(0)
        model = M.CPCModel(M.CPCEncoder(256, "layerNorm", compute_dtype="bf16"),
                           M.CPCAR(256, 256, False, 1, mode="GRU", reverse=False, compute_dtype="bf16")).to(dev)
        crit = M.CPCUnsupersivedCriterion(12, 256, 256, 128, mode=None, rnnMode="linear", dropout=False, speakerEmbedding=0,
                                          nSpeakers=0, sizeInputSeq=128, compute_dtype="bf16").to(dev)
        params = list(crit.parameters()) + list(model.parameters())
        opt = FlatAdam(params, lr=2e-4, capturable=True, fuse_zero_grad=True)
        return model, crit, opt

    gen = torch.Generator(device=dev).manual_seed(7)
    xs = [torch.randn(4, 1, 20480, device=dev, generator=gen) * 0.1 for _ in range(4)]
    label = torch.zeros(4, dtype=torch.long, device=dev)

    def eager_step(model, crit, opt, x):
        c, z, _ = model(x, label)
        losses, acc = crit(c, z, label)
        losses.sum().backward()
        opt.step()
        opt.zero_grad()
        return losses.detach().clone()

    # eager trajectory: 3 steps on xs[0] (what the graph's warm-up does) + 1 (the captured pass) then xs[1..3]
    model, crit, opt = build()
    torch.cuda.manual_seed(99)
    ref = [eager_step(model, crit, opt, xs[0]) for _ in range(4)] + [eager_step(model, crit, opt, x) for x in xs[1:]]
    ref_p = opt.flat_p.clone()
    opt.bucket.detach()

    model, crit, opt = build()
    torch.cuda.manual_seed(99)
    g = GraphedTrainStep(model, crit, opt, xs[0], label, warmup=3)   # 3 warm-up steps; the capture itself does not execute
    got = [g(xs[0])[0].clone()] + [g(x)[0].clone() for x in xs[1:]]
    assert opt.steps == 7
    # the negative-sample draws of the eager run and of the replays come from the same generator sequence, so the
    # trajectories agree up to the summation order of the fp32 atomics
    for a, b in zip(ref[3:], got):
        assert (a - b).abs().max().item() <= 2e-3, (a, b)
    # Adam turns the tiny run-to-run differences of near-zero gradients into +-lr updates: compare loosely
    assert Hh.rel_err(opt.flat_p, ref_p) <= 2e-2
    opt.bucket.detach()


def test_bucket_sinks_receive_the_same_gradients(built_lib):
    """With a GradBucket attached the backward kernels accumulate straight into the flat buffer; the result must
    equal the gradients autograd returns without a bucket (up to the atomics' summation order)."""
    from cpc_audio_b200.optim import GradBucket
    g, d, mp, cp, x, label, bi, si = Hh.load_case("cfg1_scaled")
    model, crit = Hh.build_modules(d, mp, cp, "f32")
    plain = Hh.run_modules(model, crit, x, label, bi, si)
    plain = {k: v.clone() for k, v in plain["grads"].items()}
    params = list(crit.parameters()) + list(model.parameters())
    bucket = GradBucket(params)
    crit.sampleIndices = lambda B, W, S, device: (bi.to(device), si.to(device))
    for rep in range(2):  # twice: zero() must really clear the bucket
        bucket.zero()
        c, z, _ = model(x.cuda(), label.cuda())
        losses, acc = crit(c, z, label.cuda())
        losses.sum().backward()
        named = {**{f"crit.{k}": v for k, v in crit.named_parameters()}, **{f"model.{k}": v for k, v in model.named_parameters()}}
        for k, p in named.items():
            assert p.grad.data_ptr() >= bucket.flat.data_ptr() and p.grad.data_ptr() < bucket.flat.data_ptr() + bucket.flat.numel() * 4
            assert Hh.rel_err(p.grad, plain[k]) <= 1e-5, (k, rep)
    bucket.detach()


@pytest.mark.parametrize("B,L", [(1, 20480), (5, 20480), (3, 10240), (9, 5120)])
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_ragged_shapes_against_oracle(B, L, dtype, built_lib):
    """Batch sizes that are not multiples of any tile (GRU tiles of 8 sequences, 128-row GEMM tiles, 16-frame conv0
    tiles, warps per position) and shorter windows (S = 64, 32), default widths so that the tensor-core kernels run."""
    d = O.Dims(B=B, L=L, H=256, Har=256, K=12, N=128, nLayers=1)
    mp, cp = O.make_params(d, seed=40 + B, pred_scale=30.0)
    x, label = O.make_batch(d, seed=50 + B)
    bi, si = O.make_raw_indices(d, seed=60 + B)
    ref = Hh.oracle_run(d, mp, cp, x, bi, si)
    model, crit = Hh.build_modules(d, mp, cp, dtype)
    out = Hh.run_modules(model, crit, x, label, bi, si)
    if dtype == "f32":
        assert Hh.max_rel(out["z"], ref["z"]) <= 1e-4 and Hh.max_rel(out["c"], ref["c"]) <= 1e-4
        np.testing.assert_allclose(out["losses"].cpu().numpy(), ref["losses"].numpy(), rtol=1e-5, atol=1e-4)
        for k, gr in ref["grads"].items():
            assert Hh.rel_err(out["grads"][k], gr) <= 5e-4, k
    else:
        assert Hh.rel_err(out["z"], ref["z"]) <= 1.5e-2 and Hh.rel_err(out["c"], ref["c"]) <= 1.5e-2
        assert ((out["losses"].cpu() - ref["losses"]).abs() <= 0.01 * ref["losses"].abs() + 1e-2).all()
        rep, wc, wr = Hh.grad_report(out["grads"], ref["grads"])
        Hh.record(f"ragged_batch:B{B}:L{L}:bf16", z_rel=Hh.rel_err(out["z"], ref["z"]), worst_cos=[wc[0], wc[1][0]], worst_rel=[wr[0], wr[1][1]])
        for k, gr in ref["grads"].items():
            assert _cos(out["grads"][k], gr) >= Hh.bf16_cos_floor(k, d), (k, _cos(out["grads"][k], gr))
    assert ((out["acc"].cpu() - ref["acc"]).abs() <= (0.0 if dtype == "f32" else 0.02) + Hh.acc_tolerance(ref["logits"], d)).all()


def test_config5_dims_bf16(built_lib):
    """BASELINE config 5 widths (hiddenEncoder=512, hiddenGar=512, 2-level GRU, K=16, 256 negatives) on a half-length
    window (S=256): exercises H=512 tiles, the 16-CTA clusters of the wide tensor-core GRU recurrence (gru_mma_wide.cu) and
    the warp-pair scoring kernels (H = 512, 16 chunks of negatives)."""
    d = O.Dims(B=2, L=40960, H=512, Har=512, K=16, N=256, nLayers=2)
    mp, cp = O.make_params(d, seed=70, pred_scale=30.0)
    x, label = O.make_batch(d, seed=71)
    bi, si = O.make_raw_indices(d, seed=72)
    ref = Hh.oracle_run(d, mp, cp, x, bi, si, materialize=False)
    model, crit = Hh.build_modules(d, mp, cp, "bf16")
    out = Hh.run_modules(model, crit, x, label, bi, si)
    assert Hh.rel_err(out["z"], ref["z"]) <= 1.5e-2 and Hh.rel_err(out["c"], ref["c"]) <= 1.5e-2
    assert ((out["losses"].cpu() - ref["losses"]).abs() <= 0.01 * ref["losses"].abs() + 1e-2).all()
    for k, gr in ref["grads"].items():
        assert _cos(out["grads"][k], gr) >= Hh.bf16_cos_floor(k, d), (k, _cos(out["grads"][k], gr))


@pytest.mark.parametrize("name", Hh.T_CASES)
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_transformer_heads_parity(name, dtype, built_lib):
    """rnnMode='transformer' (BASELINE config 4) against the reference fixture and the oracle, eval mode."""
    g, d, mp, cp, x, label, bi, si = Hh.load_case(name)
    ref = Hh.oracle_run(d, mp, cp, x, bi, si, heads="transformer")
    model, crit = Hh.build_modules(d, mp, cp, dtype, heads="transformer")
    out = Hh.run_modules(model, crit, x, label, bi, si)
    rep, wc, wr = Hh.grad_report(out["grads"], ref["grads"])
    Hh.record(f"theads_eval:{name}:{dtype}", dloss=(out["losses"].cpu() - ref["losses"]).abs().max().item(),
              worst_cos=[wc[0], wc[1][0]], worst_rel=[wr[0], wr[1][1]])
    if dtype == "f32":
        np.testing.assert_allclose(out["losses"].cpu().numpy(), g["losses"], rtol=1e-5, atol=2e-4)
        np.testing.assert_allclose(out["losses"].cpu().numpy(), ref["losses"].numpy(), rtol=1e-5, atol=2e-4)
        for k, gr in ref["grads"].items():
            e = Hh.rel_err(out["grads"][k], gr)
            assert e <= 2e-3, (k, e)
    else:
        assert ((out["losses"].cpu() - ref["losses"]).abs() <= 0.01 * ref["losses"].abs() + 1e-2).all()
        for k, gr in ref["grads"].items():
            assert _cos(out["grads"][k], gr) >= Hh.bf16_cos_floor(k, d), (k, _cos(out["grads"][k], gr))
    assert ((out["acc"].cpu() - ref["acc"]).abs() <= (0.0 if dtype == "f32" else (0.01 if d.H >= 256 else 0.03))
            + Hh.acc_tolerance(ref["logits"], d)).all()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs on one node")
def test_peer_adam_matches_nccl_allreduce_plus_adam(built_lib):
    """cpcb200_allreduce_adam_step (all-reduce over peer memory + Adam + zero_grad in one kernel) on 2 GPUs against an NCCL
    all-reduce + torch.optim.Adam on the same per-rank gradients (tools/peer_adam_check.py asserts <= 5e-6 relative)."""
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", os.path.join(repo, "tools", "peer_adam_check.py")],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "max rel param error" in r.stdout


# ---------------------------------------------------------------------------------------------------------------
# round 2: context networks, train-mode transformer layers (row T), arbitrary window lengths and the
# feature path (N3), gradients at the benchmarked shape, measured parity numbers (gpurun_out/parity.jsonl)
# ---------------------------------------------------------------------------------------------------------------
def _check_case(name, dtype, tag):
    g, d, mp, cp, x, label, bi, si = Hh.load_case(name)
    heads, ar = Hh.case_heads(g), Hh.case_ar(g)
    ar_masks, head_masks = Hh.case_masks(g, d)
    ref = Hh.oracle_run(d, mp, cp, x, bi, si, heads=heads, ar_masks=ar_masks, head_masks=head_masks)
    model, crit = Hh.build_modules(d, mp, cp, dtype, heads=heads, ar=ar)
    out = Hh.run_modules(model, crit, x, label, bi, si, ar_masks=ar_masks, head_masks=head_masks)
    rep, wc, wr = Hh.grad_report(out["grads"], ref["grads"])
    dl = (out["losses"].cpu() - ref["losses"]).abs().max().item()
    Hh.record(f"{tag}:{name}:{dtype}", z_rel=Hh.rel_err(out["z"], ref["z"]), c_rel=Hh.rel_err(out["c"], ref["c"]), dloss=dl,
              dacc=(out["acc"].cpu() - ref["acc"]).abs().max().item(), worst_cos=[wc[0], wc[1][0]], worst_rel=[wr[0], wr[1][1]],
              grads={k: [round(v[0], 6), float(f"{v[1]:.3e}")] for k, v in rep.items()})
    if dtype == "f32":
        assert Hh.max_rel(out["z"], ref["z"]) <= 1e-4 and Hh.max_rel(out["c"], ref["c"]) <= 2e-4
        np.testing.assert_allclose(out["losses"].cpu().numpy(), g["losses"], rtol=1e-5, atol=2e-4)   # the reference fixture
        np.testing.assert_allclose(out["losses"].cpu().numpy(), ref["losses"].numpy(), rtol=1e-5, atol=2e-4)
        for k, (cs, rl) in rep.items():
            # 2e-3, except where a ReLU pre-activation sits within fp32 rounding of zero: cfg4_train head 1 has a kept FFN
            # unit at 3.3e-7 (values are O(1)); the summation order decides its sign and the flipped unit alone moves that
            # head's attention gradients by < 1e-2 (cosine stays >= 0.9999) - a knife edge of the function, not of the kernels
            assert rl <= 2e-3 or (rl <= 1e-2 and cs >= 0.9999), (k, rl, cs)
    else:
        assert Hh.rel_err(out["z"], ref["z"]) <= 1.5e-2 and Hh.rel_err(out["c"], ref["c"]) <= 1.5e-2
        assert ((out["losses"].cpu() - ref["losses"]).abs() <= 0.01 * ref["losses"].abs() + 1e-2).all()
        for k, (cs, rl) in rep.items():
            assert cs >= Hh.bf16_cos_floor(k, d), (k, cs)
    assert ((out["acc"].cpu() - ref["acc"]).abs() <= (0.0 if dtype == "f32" else (0.01 if d.H >= 256 else 0.03))
            + Hh.acc_tolerance(ref["logits"], d)).all()


@pytest.mark.parametrize("name", Hh.AR_CASES)
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_context_networks_parity(name, dtype, built_lib):
    """--arMode LSTM (the reference default, cpc_default_config.py:74) and --arMode transformer (feature_loader.py:138-142)
    against the oracle and the fixtures generated from the unmodified reference."""
    _check_case(name, dtype, "ar")


@pytest.mark.parametrize("name", Hh.TRAIN_CASES)
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_train_mode_dropout_parity(name, dtype, built_lib):
    """Transformer layers in train() mode (dropout 0.1 on the attention probabilities and the FFN hidden, transformers.py:
    49, 92) with the keep-masks forced to the seeded ones of the fixture: loss, accuracy and every gradient."""
    _check_case(name, dtype, "train")


@pytest.mark.parametrize("L,B", [(20000, 2), (12345, 3), (7777, 1), (20480 + 163, 2)])
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_arbitrary_window_length(L, B, dtype, built_lib):
    """Window lengths that are not multiples of 160 (the Conv1d stack of model.py:83-92 floors at every layer): forward and
    backward against the oracle.  Exercises the partial conv0 tiles, the ragged transposed-conv rows and odd S."""
    d = O.Dims(B=B, L=L, H=256, Har=256, K=4, N=16, nLayers=1)
    S = O.encoder_forward(torch.zeros(1, 1, L), O.make_params(d, seed=1)[0]).shape[2]
    from cpc_audio_b200.model import frames_for
    assert frames_for(L) == S
    mp, cp = O.make_params(d, seed=80 + B, pred_scale=30.0)
    x, label = O.make_batch(d, seed=81)
    g = torch.Generator().manual_seed(82)
    W = S - d.K
    bi = torch.randint(0, B, (d.N * W * B,), generator=g)
    si = torch.randint(1, S, (d.N * W * B,), generator=g)
    ref = Hh.oracle_run(d, mp, cp, x, bi, si)
    model, crit = Hh.build_modules(d, mp, cp, dtype)
    crit.wPrediction  # noqa: B018
    out = Hh.run_modules(model, crit, x, label, bi, si)
    assert out["z"].shape == (B, S, 256)
    rep, wc, wr = Hh.grad_report(out["grads"], ref["grads"])
    Hh.record(f"ragged:L{L}:B{B}:{dtype}", z_rel=Hh.rel_err(out["z"], ref["z"]), worst_cos=[wc[0], wc[1][0]], worst_rel=[wr[0], wr[1][1]])
    if dtype == "f32":
        assert Hh.max_rel(out["z"], ref["z"]) <= 1e-4 and Hh.max_rel(out["c"], ref["c"]) <= 1e-4
        np.testing.assert_allclose(out["losses"].cpu().numpy(), ref["losses"].numpy(), rtol=1e-5, atol=1e-4)
        assert wr[1][1] <= 5e-4, wr
    else:
        assert Hh.rel_err(out["z"], ref["z"]) <= 1.5e-2 and Hh.rel_err(out["c"], ref["c"]) <= 1.5e-2
        assert wc[1][0] >= 0.985, wc      # measured 0.9931 .. 0.9949 (encoder tensors, B <= 3)
    # no_grad forward (nothing saved, activations ping-pong in the workspace) gives the same features
    with torch.no_grad():
        c2, z2, _ = model(x.cuda(), label.cuda())
    assert Hh.max_rel(z2, out["z"]) <= 1e-6 and Hh.max_rel(c2, out["c"]) <= 1e-6


@pytest.mark.parametrize("name", Hh.FEATURE_CASES)
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_feature_path_chunks_with_carried_state(name, dtype, built_lib):
    """feature_loader.py:228-269 semantics without the reference package: chunks of maxSizeSeq samples (the last one of
    arbitrary length) through the no_grad model with keepHidden, against oracle.feature_forward and the fixture written from
    the unmodified buildFeature."""
    g = np.load(f"{Hh.GOLDEN}/{name}.npz")
    nl, n, chunk, H = int(g["nLayers"]), int(g["n"]), int(g["chunk"]), int(g["H"])
    d = O.Dims(B=1, L=chunk, H=H, Har=H, nLayers=nl)
    mp, cp = O.make_params(d, seed=21, ar=str(g["ar"]))
    seq = torch.randn(n, generator=torch.Generator().manual_seed(22)) * 0.1
    want = O.feature_forward(seq, mp, nl, max_size_seq=chunk, keep_hidden=True)
    model, _ = Hh.build_modules(d, mp, cp, dtype, ar=str(g["ar"]), keep_hidden=True)
    model.eval()
    outs, start = [], 0
    with torch.no_grad():
        while start < n:
            sub = seq[start:min(n, start + chunk)].view(1, 1, -1).cuda()
            c, z, _ = model(sub, None)
            outs.append(c.cpu())
            start += chunk
    feat = torch.cat(outs, 1)
    assert feat.shape == want.shape == tuple(g["shape"])
    e = Hh.rel_err(feat, want)
    Hh.record(f"features:{name}:{dtype}", rel=e)
    assert e <= (2e-5 if dtype == "f32" else 3e-2), e
    np.testing.assert_allclose(Hh.subsample(feat), g["feat_sub"], rtol=0, atol=(2e-4 if dtype == "f32" else 5e-2) * np.abs(g["feat_sub"]).max())


_FULL = {}


def _full_size_oracle():
    if not _FULL:
        d = O.Dims(B=64, L=20480, H=256, Har=256, K=12, N=128, nLayers=1)
        mp, cp = O.make_params(d, seed=7, pred_scale=30.0)
        x, label = O.make_batch(d, seed=77)
        bi, si = O.make_raw_indices(d, seed=777)
        torch.set_num_threads(max(1, min(32, (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else 8))))
        _FULL["v"] = (d, mp, cp, x, label, bi, si, Hh.oracle_run(d, mp, cp, x, bi, si, materialize=False))
    return _FULL["v"]


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_full_size_gradients_against_oracle(dtype, built_lib):
    """BASELINE config 2 shape (B = 64 windows of 20480 samples, default widths): EVERY parameter gradient against the CPU
    oracle (einsum scoring, no 11.8 GB materialisation).  This is the shape whose one-wave split-K weight gradients, 8-cluster
    GRU tiling and 950 K-row scatter-add are benchmarked."""
    d, mp, cp, x, label, bi, si, ref = _full_size_oracle()
    model, crit = Hh.build_modules(d, mp, cp, dtype)
    out = Hh.run_modules(model, crit, x, label, bi, si)
    rep, wc, wr = Hh.grad_report(out["grads"], ref["grads"])
    dl = (out["losses"].cpu() - ref["losses"]).abs().max().item()
    Hh.record(f"full_size:{dtype}", z_rel=Hh.rel_err(out["z"], ref["z"]), c_rel=Hh.rel_err(out["c"], ref["c"]), dloss=dl,
              dacc=(out["acc"].cpu() - ref["acc"]).abs().max().item(), worst_cos=[wc[0], wc[1][0]], worst_rel=[wr[0], wr[1][1]],
              grads={k: [round(v[0], 6), float(f"{v[1]:.3e}")] for k, v in rep.items()})
    if dtype == "f32":
        assert dl <= 2e-4 and wr[1][1] <= 1e-3, (dl, wr)
    else:
        assert dl <= 0.01 * ref["losses"].abs().max().item() + 1e-2
        assert Hh.rel_err(out["z"], ref["z"]) <= 1e-2 and Hh.rel_err(out["c"], ref["c"]) <= 1e-2
        assert (out["acc"].cpu() - ref["acc"]).abs().max().item() <= 0.005    # +-0.5 % abs (measured 0.013 %)
        for k, (cs, rl) in rep.items():
            assert cs >= Hh.bf16_cos_floor(k, d), (k, cs, rl)   # >= 0.999 on every tensor but conv0.weight (0.995)


def test_config5_full_length_bf16(built_lib):
    """BASELINE config 5 at its full window (hiddenEncoder = hiddenGar = 512, 2-level GRU, K = 16, 256 negatives,
    81920 samples -> S = 512), B = 2: forward and every gradient against the oracle."""
    d = O.Dims(B=2, L=81920, H=512, Har=512, K=16, N=256, nLayers=2)
    mp, cp = O.make_params(d, seed=90, pred_scale=30.0)
    x, label = O.make_batch(d, seed=91)
    bi, si = O.make_raw_indices(d, seed=92)
    ref = Hh.oracle_run(d, mp, cp, x, bi, si, materialize=False)
    model, crit = Hh.build_modules(d, mp, cp, "bf16")
    out = Hh.run_modules(model, crit, x, label, bi, si)
    rep, wc, wr = Hh.grad_report(out["grads"], ref["grads"])
    Hh.record("config5:S512:bf16", z_rel=Hh.rel_err(out["z"], ref["z"]), c_rel=Hh.rel_err(out["c"], ref["c"]),
              dloss=(out["losses"].cpu() - ref["losses"]).abs().max().item(), worst_cos=[wc[0], wc[1][0]], worst_rel=[wr[0], wr[1][1]])
    assert Hh.rel_err(out["z"], ref["z"]) <= 1.5e-2 and Hh.rel_err(out["c"], ref["c"]) <= 1.5e-2
    assert ((out["losses"].cpu() - ref["losses"]).abs() <= 0.01 * ref["losses"].abs() + 1e-2).all()
    assert wc[1][0] >= 0.985, wc          # measured 0.9947 (gEncoder.conv0.weight, B = 2)


def test_wide_lstm_hidden512_bf16(built_lib):
    """--arMode LSTM at hiddenGar = 512, 2 levels, B = 3 (a partly filled batch tile): the 16-CTA-cluster wide LSTM recurrence
    (lstm_mma.cu) forward and every gradient against the oracle."""
    d = O.Dims(B=3, L=20480, H=512, Har=512, K=12, N=128, nLayers=2)
    mp, cp = O.make_params(d, seed=93, pred_scale=30.0, ar="LSTM")
    x, label = O.make_batch(d, seed=94)
    bi, si = O.make_raw_indices(d, seed=95)
    ref = Hh.oracle_run(d, mp, cp, x, bi, si, materialize=False)
    model, crit = Hh.build_modules(d, mp, cp, "bf16", ar="LSTM")
    out = Hh.run_modules(model, crit, x, label, bi, si)
    rep, wc, wr = Hh.grad_report(out["grads"], ref["grads"])
    Hh.record("lstm512:bf16", z_rel=Hh.rel_err(out["z"], ref["z"]), c_rel=Hh.rel_err(out["c"], ref["c"]),
              dloss=(out["losses"].cpu() - ref["losses"]).abs().max().item(), worst_cos=[wc[0], wc[1][0]], worst_rel=[wr[0], wr[1][1]])
    assert Hh.rel_err(out["z"], ref["z"]) <= 1.5e-2 and Hh.rel_err(out["c"], ref["c"]) <= 1.5e-2
    assert ((out["losses"].cpu() - ref["losses"]).abs() <= 0.01 * ref["losses"].abs() + 1e-2).all()
    for k, gr in ref["grads"].items():
        assert _cos(out["grads"][k], gr) >= Hh.bf16_cos_floor(k, d), (k, _cos(out["grads"][k], gr))


def test_optimizer_state_dict_interchanges_with_torch_adam(built_lib):
    """cpc/train.py:220 saves optimizer.state_dict(), :339-343 loads it: a FlatAdam checkpoint loads into torch.optim.Adam
    and vice versa, and both continue on the same trajectory; StepLR drives FlatAdam's learning rate (train.py:351-355)."""
    from cpc_audio_b200.optim import FlatAdam
    gen = torch.Generator(device="cuda").manual_seed(3)
    shapes = [(64, 64, 4), (192,), (1, 64, 1), (12, 64, 64)]

    def grads(it):
        g = torch.Generator(device="cuda").manual_seed(100 + it)
        return [torch.randn(s, device="cuda", generator=g) for s in shapes]

    p0 = [torch.randn(s, device="cuda", generator=gen) for s in shapes]
    for capturable in (False, True):
        pt = [torch.nn.Parameter(p.clone()) for p in p0]
        pf = [torch.nn.Parameter(p.clone()) for p in p0]
        ot = torch.optim.Adam(pt, lr=1e-3)
        of = FlatAdam(pf, lr=1e-3, capturable=capturable, fuse_zero_grad=capturable)
        st = torch.optim.lr_scheduler.StepLR(ot, 2, gamma=0.5)
        sf = torch.optim.lr_scheduler.StepLR(of, 2, gamma=0.5)
        for it in range(5):
            for a, b, gg in zip(pt, pf, grads(it)):
                a.grad = gg.clone()
                b.grad.copy_(gg)
            ot.step(); of.step(); ot.zero_grad(); of.zero_grad()
            st.step(); sf.step()
        assert of.param_groups[0]["lr"] == ot.param_groups[0]["lr"] == 1e-3 * 0.25
        for a, b in zip(pt, pf):
            assert (a - b).abs().max().item() <= 3e-6 * max(1.0, a.abs().max().item())
        # cross-load: torch -> flat, flat -> torch, then 3 more steps on each
        # (through torch.save / torch.load as train.py:220,339 does: Optimizer.load_state_dict does not copy tensors that
        # already have the right dtype and device, so loading a LIVE state_dict would alias the two optimizers' moments)
        import io
        sds = []
        for o in (ot, of):
            buf = io.BytesIO()
            torch.save(o.state_dict(), buf)
            buf.seek(0)
            sds.append(torch.load(buf, map_location="cpu"))
        sd_t, sd_f = sds
        assert set(sd_f["param_groups"][0]) >= {"lr", "betas", "eps", "weight_decay", "amsgrad", "params"}
        assert int(float(sd_f["state"][0]["step"])) == 5 and sd_f["state"][0]["exp_avg"].shape == pf[0].shape
        pt2 = [torch.nn.Parameter(p.detach().clone()) for p in pf]
        pf2 = [torch.nn.Parameter(p.detach().clone()) for p in pt]
        ot2 = torch.optim.Adam(pt2, lr=1e-3)
        of2 = FlatAdam(pf2, lr=1e-3, capturable=capturable, fuse_zero_grad=capturable)
        ot2.load_state_dict(sd_f)
        of2.load_state_dict(sd_t)
        assert of2.steps == 5 and of2.param_groups[0]["lr"] == 1e-3 * 0.25
        for it in range(5, 8):
            for a, b, c_, e, gg in zip(pt, pf, pt2, pf2, grads(it)):
                a.grad = gg.clone(); c_.grad = gg.clone()
                b.grad.copy_(gg); e.grad.copy_(gg)
            for o in (ot, of, ot2, of2):
                o.step(); o.zero_grad()
        for a, b, c_, e in zip(pt, pf, pt2, pf2):
            tol = 3e-6 * max(1.0, a.abs().max().item())
            assert (a - b).abs().max().item() <= tol, "FlatAdam vs torch.optim.Adam"
            assert (a - c_).abs().max().item() <= tol, "torch.optim.Adam resumed from a FlatAdam checkpoint"
            assert (a - e).abs().max().item() <= tol, "FlatAdam resumed from a torch.optim.Adam checkpoint"
        for o in (of, of2):
            o.bucket.detach()


def test_bucket_survives_default_zero_grad_of_torch_adam(built_lib):
    """torch.optim.Optimizer.zero_grad() defaults to set_to_none=True; a GradBucket must keep receiving the
    gradients (and must never all-reduce a stale buffer) when a stock torch.optim.Adam drives the loop."""
    from cpc_audio_b200.optim import GradBucket
    g, d, mp, cp, x, label, bi, si = Hh.load_case("small")
    model, crit = Hh.build_modules(d, mp, cp, "f32")
    crit.sampleIndices = lambda B, W, S, device: (bi.to(device), si.to(device))
    params = list(crit.parameters()) + list(model.parameters())
    bucket = GradBucket(params)
    opt = torch.optim.Adam(params, lr=1e-3)
    ref_model, ref_crit = Hh.build_modules(d, mp, cp, "f32")
    ref_crit.sampleIndices = crit.sampleIndices
    ref_params = list(ref_crit.parameters()) + list(ref_model.parameters())
    ref_opt = torch.optim.Adam(ref_params, lr=1e-3)
    for step in range(3):
        for m, c_, o in ((model, crit, opt), (ref_model, ref_crit, ref_opt)):
            cc, zz, _ = m(x.cuda(), label.cuda())
            losses, _ = c_(cc, zz, label.cuda())
            losses.sum().backward()
        # every gradient sits in the bucket again although zero_grad() set it to None after the previous step
        for p, v, q in zip(params, bucket.views, ref_params):
            assert p.grad is v and Hh.rel_err(p.grad, q.grad) <= 1e-5, step
        opt.step(); opt.zero_grad()          # set_to_none=True
        ref_opt.step(); ref_opt.zero_grad()
        assert all(p.grad is None for p in params)
    for p, q in zip(params, ref_params):
        assert Hh.rel_err(p, q) <= 1e-5
    bucket.detach()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs on one node")
def test_dataparallel_two_devices_in_process(built_lib):
    """cpc/train.py:372-375 with nGPU = 2: torch.nn.DataParallel replicates the modules and runs the replicas in two host
    threads, each bound to its own device (the C ABI must be callable concurrently).  The gradients on the
    master parameters must equal the SUM over the two halves computed one after the other (cpc/train.py:85 sums the
    gathered per-replica losses), with each replica drawing its negatives from its own device generator."""
    d = O.Dims(B=6, L=20480, H=256, Har=256, K=12, N=128, nLayers=1)
    mp, cp = O.make_params(d, seed=33, pred_scale=30.0)
    x, label = O.make_batch(d, seed=34)
    model, crit = Hh.build_modules(d, mp, cp, "f32", device="cuda:0")
    dp_model = torch.nn.DataParallel(model, device_ids=[0, 1])
    dp_crit = torch.nn.DataParallel(crit, device_ids=[0, 1])
    for dev in (0, 1):
        with torch.cuda.device(dev):
            torch.cuda.manual_seed(100 + dev)
    c, z, lab = dp_model(x.cuda(0), label.cuda(0))
    losses, acc = dp_crit(c, z, lab)
    assert losses.shape == (2, d.K) and acc.shape == (2, d.K)   # one row per replica, as train.py:85-99 expects
    losses.sum().backward()
    got = {**{f"model.{k}": v.grad.clone() for k, v in model.named_parameters()},
           **{f"crit.{k}": v.grad.clone() for k, v in crit.named_parameters()}}
    # the same two halves, sequentially, each on the device (and generator) its replica used
    want = None
    per_replica_losses = []
    for dev, sl in ((0, slice(0, 3)), (1, slice(3, 6))):
        m2, c2 = Hh.build_modules(d, mp, cp, "f32", device=f"cuda:{dev}")
        with torch.cuda.device(dev):
            torch.cuda.manual_seed(100 + dev)
            cc, zz, _ = m2(x[sl].cuda(dev), label[sl].cuda(dev))
            ll, _ = c2(cc, zz, label[sl].cuda(dev))
            ll.sum().backward()
        per_replica_losses.append(ll.detach().cpu())
        g = {**{f"model.{k}": v.grad.cpu() for k, v in m2.named_parameters()}, **{f"crit.{k}": v.grad.cpu() for k, v in c2.named_parameters()}}
        want = g if want is None else {k: want[k] + g[k] for k in g}
    assert (losses.detach().cpu() - torch.cat(per_replica_losses)).abs().max().item() <= 1e-5
    for k, v in want.items():
        assert Hh.rel_err(got[k], v) <= 2e-5, (k, Hh.rel_err(got[k], v))
