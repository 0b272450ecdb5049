"""The drop-in promise of north_star: ``cpc/train.py`` is unchanged.  These tests drive the reference's OWN code - its
factories (feature_loader.getEncoder/getAR, train.getCriterion), its ``trainStep`` loop (cpc/train.py:64-119), its
``torch.optim.Adam`` + ``DataParallel`` wrapping (train.py:335, 372-375), ``FeatureModule`` / ``buildFeature``
(feature_loader.py:15-38, 228-269) - over the B200 modules installed by ``cpc_audio_b200.patch``.

They need the reference package (``/root/reference`` or the pip --target copy in ``baseline/_ref``) and skip without it."""
import copy

import numpy as np
import pytest
import torch

from oracle import cpc_oracle as O
from tests import helpers as Hh
from tests import ref_driver as R

needs_ref = pytest.mark.skipif(R.reference_or_none() is None, reason="reference package not present (baseline/_ref)")

MODES = [("GRU", "linear"), ("LSTM", "transformer"), ("LSTM", "linear"), ("transformer", "linear")]


@needs_ref
@pytest.mark.parametrize("arMode,rnnMode", MODES)
def test_factories_build_b200_modules_with_reference_state_dict_keys(arMode, rnnMode):
    """train.py:307-311 through the patched package: same classes, attributes and state_dict keys / shapes as the reference
    builds for the same flags - for the reference's DEFAULT configuration (LSTM + transformer heads) too."""
    import cpc_audio_b200 as M
    args = R.default_args(arMode=arMode, rnnMode=rnnMode)
    rm, rc, *_ = R.build(args, b200=False, device="cpu")
    bm, bc, *_ = R.build(args, b200=True, device="cpu")
    R.use_b200_modules(False)
    assert isinstance(bm, M.CPCModel) and isinstance(bm.gEncoder, M.CPCEncoder) and isinstance(bc, M.CPCUnsupersivedCriterion)
    for ours, theirs in ((bm, rm), (bc, rc)):
        a, b = ours.state_dict(), theirs.state_dict()
        assert list(a) == list(b), (set(a) ^ set(b))
        assert all(a[k].shape == b[k].shape and a[k].dtype == b[k].dtype for k in a)
        ours.load_state_dict(b, strict=True)      # checkpoints interchange (feature_loader.py:180, 204-207)
        theirs.load_state_dict(a, strict=True)
    assert bm.gEncoder.DOWNSAMPLING == rm.gEncoder.DOWNSAMPLING == 160
    if arMode != "transformer":
        assert bm.gAR.getDimOutput() == rm.gAR.getDimOutput() and bm.gAR.keepHidden == rm.gAR.keepHidden


def _batches(n, B, L, dev, seed=5):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(B, 1, L, generator=g) * 0.1, torch.zeros(B, dtype=torch.long)) for _ in range(n)]


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("arMode,rnnMode", [("GRU", "linear"), ("LSTM", "linear")])
def test_reference_trainstep_over_b200_modules_tracks_the_reference_on_gpu(arMode, rnnMode, built_lib):
    """10 steps of the UNMODIFIED trainStep (train.py:64-119) + torch.optim.Adam + DataParallel(device_ids=[0]):
    (a) over the reference's own modules on the B200 (torch / cuDNN fp32), (b) over the patched B200 modules in fp32.
    Same seeds -> the same on-device torch.randint stream -> the same negatives: the loss trajectories agree."""
    args = R.default_args(arMode=arMode, rnnMode=rnnMode, hiddenEncoder=256, hiddenGar=256)
    batches = _batches(10, 4, 20480, "cuda")
    torch.backends.cudnn.allow_tf32 = False        # the reference arm must really be fp32 (cuDNN convs default to TF32)
    torch.backends.cuda.matmul.allow_tf32 = False
    rm, rc, rmd, rcd, ropt = R.build(args, b200=False, seed=3)
    state = (copy.deepcopy(rm.state_dict()), copy.deepcopy(rc.state_dict()))
    torch.cuda.manual_seed(77)
    ref_losses = R.train_steps(rmd, rcd, ropt, batches)
    import os
    os.environ["CPC_B200_DTYPE"] = "f32"
    try:
        bm, bc, bmd, bcd, bopt = R.build(args, b200=True, seed=3, state=state)
    finally:
        os.environ.pop("CPC_B200_DTYPE")
    torch.cuda.manual_seed(77)
    got = R.train_steps(bmd, bcd, bopt, batches)
    R.use_b200_modules(False)
    for i, (a, b) in enumerate(zip(ref_losses, got)):
        assert (a - b).abs().max().item() <= 2e-3, (i, a, b)
    assert ref_losses[-1].mean() < ref_losses[0].mean()
    # the UPDATE each parameter tensor received over the 10 Adam steps points the same way in both runs (Adam turns the
    # run-to-run noise of near-zero gradients into +-lr moves, so individual elements may differ; tensors that start at
    # zero - the ChannelNorm biases - are pure update)
    for (k, v), (kr, vr) in zip(state[0].items(), rm.state_dict().items()):
        if not v.is_floating_point() or v.numel() < 256:
            continue
        cs = Hh.cosine(bm.state_dict()[k].cpu() - v.cpu(), vr.cpu() - v.cpu())
        assert cs >= 0.9, (k, cs)


@needs_ref
@pytest.mark.gpu
def test_on_device_randint_stream_equals_reference_sampleClean(built_lib):
    """criterion.py:181-199 on CUDA: under the same torch.cuda.manual_seed the B200 criterion draws exactly the indices the
    reference's sampleClean draws (same two torch.randint calls, same order / shape / dtype / device) - bit-exact."""
    args = R.default_args(arMode="GRU", rnnMode="linear")
    rm, rc, *_ = R.build(args, b200=False, seed=1)
    bm, bc, *_ = R.build(args, b200=True, seed=1)
    R.use_b200_modules(False)
    B, S, H = 6, 128, 256
    z = torch.randn(B, S, H, device="cuda")
    rec = {}
    real = torch.randint

    def spy(tag):
        def f(*a, **k):
            out = real(*a, **k)
            rec.setdefault(tag, []).append(out.clone())
            return out
        return f

    try:
        torch.cuda.manual_seed(2024)
        torch.randint = spy("ref")
        rc.sampleClean(z, S - 12)
        torch.cuda.manual_seed(2024)
        torch.randint = spy("b200")
        bi, si = bc.sampleIndices(B, S - 12, S, z.device)
    finally:
        torch.randint = real
    assert len(rec["ref"]) == len(rec["b200"]) == 2
    for a, b in zip(rec["ref"], rec["b200"]):
        assert a.dtype == b.dtype == torch.int64 and a.device == b.device and torch.equal(a, b)
    ext = bc.extIndices(bi, si, (B, S, H, 256, 12, 128, 0)).cpu().numpy()
    assert np.array_equal(ext, O.ext_indices_np(bi.cpu().numpy(), si.cpu().numpy(), B, 128, S - 12, S))


def test_dropout_masks_follow_the_reference_call_order_cpu():
    """Host logic of row T in train(): ``draw_dropout_masks`` makes, per layer, one nn.Dropout-style draw for the attention
    probabilities (B*heads, W, W) and then one for the FFN hidden (B, W, dff) - the order of transformers.py:49 then :92 -
    and the stacked uint8 buffers hold exactly those draws (checked on the CPU generator)."""
    from cpc_audio_b200.criterion import draw_dropout_masks
    K, B, W, nh, dff, p = 3, 2, 7, 4, 16, 0.25
    torch.manual_seed(1234)
    att, ffn = draw_dropout_masks(K, B, W, nh, dff, p, torch.device("cpu"))
    assert att.shape == (K, B * nh, W, W) and ffn.shape == (K, B * W, dff) and att.dtype == torch.uint8 and ffn.dtype == torch.uint8
    torch.manual_seed(1234)
    for k in range(K):
        a = torch.nn.functional.dropout(torch.ones(B * nh, W, W), p, True) != 0
        f = torch.nn.functional.dropout(torch.ones(B, W, dff), p, True) != 0
        assert torch.equal(att[k].bool(), a) and torch.equal(ffn[k].bool(), f.view(B * W, dff)), k


@needs_ref
@pytest.mark.gpu
def test_train_mode_dropout_masks_equal_the_reference_draws(built_lib):
    """Row T in train(): the keep-masks drawn by ``draw_dropout_masks`` equal, element for element, the masks the reference's
    nn.Dropout layers apply under the same CUDA generator state (transformers.py:49, :92) - so a train()-mode step of
    the B200 criterion sees the reference's dropout pattern, not just its distribution."""
    from cpc_audio_b200.criterion import draw_dropout_masks
    ref = R.reference_or_none()
    B, W, D = 3, 116, 256
    layer = ref.transformers.TransformerLayer(sizeSeq=W, dmodel=D).cuda().train()
    seen = []
    real = torch.nn.Dropout.forward

    def spy(self, inp):
        out = real(self, inp)
        seen.append((out != 0, inp != 0))  # (kept, informative): a zero input says nothing about its mask bit
        return out

    x = torch.randn(B, W, D, device="cuda")
    try:
        torch.nn.Dropout.forward = spy
        torch.cuda.manual_seed(99)
        layer(x)
    finally:
        torch.nn.Dropout.forward = real
    torch.cuda.manual_seed(99)
    att, ffn = draw_dropout_masks(1, B, W, 8, 2048, 0.1, x.device)
    assert len(seen) == 2
    (a_kept, a_inf), (f_kept, f_inf) = seen
    a_kept, a_inf = a_kept.reshape(B * 8, W, W), a_inf.reshape(B * 8, W, W)
    f_kept, f_inf = f_kept.reshape(B * W, 2048), f_inf.reshape(B * W, 2048)
    # (probabilities above the diagonal are exactly 0 - causal mask - and about half of the ReLU outputs are 0)
    assert a_inf.float().mean().item() > 0.45 and f_inf.float().mean().item() > 0.3
    assert torch.equal(a_kept[a_inf], att[0].bool()[a_inf])
    assert torch.equal(f_kept[f_inf], ffn[0].bool()[f_inf])
    assert abs(att.float().mean().item() - 0.9) < 5e-3 and abs(ffn.float().mean().item() - 0.9) < 5e-3


@needs_ref
@pytest.mark.gpu
def test_reference_default_config_trains_through_the_patched_package(built_lib):
    """cpc_default_config.py defaults (arMode=LSTM, rnnMode=transformer - dropout 0.1 active in train()) through the
    unmodified trainStep: runs, is finite, and the loss goes down."""
    args = R.default_args()
    assert args.arMode == "LSTM" and args.rnnMode == "transformer"
    bm, bc, bmd, bcd, bopt = R.build(args, b200=True, seed=2)
    R.use_b200_modules(False)
    x = _batches(1, 4, 20480, "cuda")[0]
    torch.cuda.manual_seed(5)
    losses = R.train_steps(bmd, bcd, bopt, [x] * 12)
    assert all(torch.isfinite(l).all() for l in losses)
    assert losses[-1].mean() < losses[0].mean(), (losses[0], losses[-1])


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("name", Hh.FEATURE_CASES)
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_buildFeature_through_reference_FeatureModule(name, dtype, built_lib):
    """feature_loader.py:228-269 UNMODIFIED over the B200 model: chunks of maxSizeSeq samples whose last one has an
    arbitrary length, no_grad, hidden state carried by keepHidden; against the committed fixture of the reference run."""
    ref = R.reference_or_none()
    g = np.load(f"{Hh.GOLDEN}/{name}.npz")
    nl, n, chunk, H = int(g["nLayers"]), int(g["n"]), int(g["chunk"]), int(g["H"])
    d = O.Dims(B=1, L=chunk, H=H, Har=H, nLayers=nl)
    mp, cp = O.make_params(d, seed=21, ar=str(g["ar"]))
    seq = torch.randn(n, generator=torch.Generator().manual_seed(22)) * 0.1
    model, _ = Hh.build_modules(d, mp, cp, dtype, ar=str(g["ar"]), keep_hidden=True)
    model.eval()
    fm = ref.feature_loader.FeatureModule(model, False)
    fl = ref.feature_loader
    real = fl.torchaudio.load
    fl.torchaudio.load = lambda path: (seq.view(1, -1), 16000)
    try:
        feat = fl.buildFeature(fm, "seeded.wav", strict=False, maxSizeSeq=chunk)
    finally:
        fl.torchaudio.load = real
    assert list(feat.shape) == list(g["shape"])
    sub = Hh.subsample(feat)
    tol = 2e-4 if dtype == "f32" else 4e-2
    assert np.abs(sub - g["feat_sub"]).max() <= tol * np.abs(g["feat_sub"]).max()
