"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares,
size queries / argument validation work without a GPU, and the Python mirrors keep the reference's surface."""
import ctypes
import os
import re

import pytest
import torch

from tests.conftest import REPO


def header_symbols():
    src = open(os.path.join(REPO, "include", "cpc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cpcb200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built_lib):
    from cpc_audio_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 19
    raw = ctypes.CDLL(built_lib)
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in include/cpc_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in cpc_audio_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == syms
    assert _lib.lib().cpcb200_version() == 200


def test_size_queries_and_validation(built_lib):
    from cpc_audio_b200 import _lib as L
    lib = L.lib()
    d = L.make_dims(64, 20480, 256, 256, 12, 128, 1, L.BF16)
    assert lib.cpcb200_encoder_save_bytes(d) > 64 * 4096 * 256 * 2
    assert lib.cpcb200_encoder_ws_bytes(d, 1) > lib.cpcb200_encoder_ws_bytes(d, 0) > 0
    assert lib.cpcb200_gru_save_bytes(d) > 0 and lib.cpcb200_gru_ws_bytes(d, 0) > 0
    assert lib.cpcb200_criterion_save_bytes(d) > 64 * 116 * 12 * 129 * 4
    # any window long enough for one output frame is accepted (feature extraction feeds arbitrary chunk lengths) ...
    odd = L.make_dims(3, 20481, 256, 256, 12, 128, 1, L.BF16)
    assert lib.cpcb200_encoder_save_bytes(odd) > 0 and lib.cpcb200_encoder_ws_bytes(odd, 2) > 0
    assert lib.cpcb200_lstm_save_bytes(d) > 0 and lib.cpcb200_lstm_ws_bytes(d, 2) > lib.cpcb200_lstm_ws_bytes(d, 0) > 0
    assert lib.cpcb200_tlayer_save_bytes(d, 2048, 8) > 0 and lib.cpcb200_tlayer_ws_bytes(d, 2048, 8, 1) > 0
    # ... a window without a single frame is not
    bad = L.make_dims(64, 100, 256, 256, 12, 128, 1, L.BF16)
    assert lib.cpcb200_encoder_save_bytes(bad) == 0
    z = ctypes.c_void_p(16)
    st = lib.cpcb200_sample_ext_idx(bad, z, z, z, None)
    assert st == -1 and b"too short" in lib.cpcb200_last_error()
    bad2 = L.make_dims(2, 20480, 100, 256, 12, 128, 1, L.F32)
    assert lib.cpcb200_gru_ws_bytes(bad2, 0) == 0
    st = lib.cpcb200_adam_step(None, z, z, z, 4, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, None)
    assert st == -5 and b"NULL" in lib.cpcb200_last_error()
    st = lib.cpcb200_adam_step_dev(z, z, z, z, 4, 1e-3, 0.9, 0.999, 1e-8, 0.0, None, 1, None)
    assert st == -5 and b"NULL" in lib.cpcb200_last_error()


REF_MODEL_KEYS = [f"gEncoder.{n}{i}.{p}" for i in range(5) for n in ("conv", "batchNorm") for p in ("weight", "bias")] + \
                 [f"gAR.baseNet.{n}_l0" for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]


def test_module_surface_matches_reference():
    import cpc_audio_b200 as M
    enc = M.CPCEncoder(256, "layerNorm")
    ar = M.CPCAR(256, 256, False, 1, mode="GRU", reverse=False)
    model = M.CPCModel(enc, ar)
    sd = model.state_dict()
    assert sorted(sd) == sorted(REF_MODEL_KEYS)
    assert sd["gEncoder.conv1.weight"].shape == (256, 256, 8) and sd["gEncoder.batchNorm3.bias"].shape == (1, 256, 1)
    assert sd["gAR.baseNet.weight_hh_l0"].shape == (768, 256)
    assert enc.DOWNSAMPLING == 160 and enc.dimEncoded == 256 and enc.getDimOutput() == 256 and ar.getDimOutput() == 256
    assert ar.hidden is None and ar.keepHidden is False
    crit = M.CPCUnsupersivedCriterion(12, 256, 256, 128, mode=None, rnnMode="linear", dropout=False, speakerEmbedding=0,
                                      nSpeakers=0, sizeInputSeq=128)
    assert sorted(crit.state_dict()) == sorted(f"wPrediction.predictors.{k}.weight" for k in range(12))
    assert M.CPCUnsupervisedCriterion is M.CPCUnsupersivedCriterion
    assert crit.warmUp() is False and crit.update() is None
    n = sum(p.numel() for p in model.parameters()) + sum(p.numel() for p in crit.parameters())
    assert n == 2498304


def test_unsupported_modes_raise_instead_of_falling_back():
    import cpc_audio_b200 as M
    with pytest.raises(NotImplementedError):
        M.CPCAR(256, 256, False, 1, mode="RNN")
    with pytest.raises(NotImplementedError):
        M.CPCAR(256, 256, False, 1, mode="GRU", reverse=True)
    with pytest.raises(NotImplementedError):
        M.buildTransformerAR(256, 1, 128, True)   # abspos
    lstm = M.CPCAR(256, 256, True, 2, mode="LSTM")  # the reference default context network: same keys as nn.LSTM
    assert lstm.baseNet.weight_hh_l1.shape == (1024, 256) and lstm.getDimOutput() == 256
    with pytest.raises(NotImplementedError):
        M.CPCEncoder(256, "batchNorm")
    with pytest.raises(ValueError):
        M.CPCEncoder(256, "nope")
    with pytest.raises(NotImplementedError):
        M.CPCUnsupersivedCriterion(12, 256, 256, 128, rnnMode="conv4")
    with pytest.raises(NotImplementedError):
        M.CPCUnsupersivedCriterion(12, 128, 256, 128, rnnMode="transformer")  # needs hiddenGar == hiddenEncoder
    with pytest.raises(NotImplementedError):
        M.CPCUnsupersivedCriterion(12, 256, 256, 128, rnnMode="linear", speakerEmbedding=8, nSpeakers=4)
    model = M.CPCModel(M.CPCEncoder(64), M.CPCAR(64, 64, False, 1))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(torch.zeros(1, 1, 1600), torch.zeros(1, dtype=torch.long))


def test_prediction_weights_share_one_buffer():
    import cpc_audio_b200 as M
    crit = M.CPCUnsupersivedCriterion(4, 64, 128, 8, rnnMode="linear", sizeInputSeq=16)
    flat = crit.wPrediction.stacked()
    assert flat.shape == (4, 128, 64)
    crit.wPrediction.predictors[2].weight.data.fill_(3.0)
    assert crit.wPrediction.stacked()[2].eq(3.0).all()          # in-place updates (optimizer) stay visible
    sd = {k: torch.full_like(v, 7.0) for k, v in crit.state_dict().items()}
    crit.load_state_dict(sd)
    assert crit.wPrediction.stacked().eq(7.0).all()
    crit2 = M.CPCUnsupersivedCriterion(4, 64, 128, 8, rnnMode="linear", sizeInputSeq=16)
    crit2.double().float()                                      # _apply re-allocates every parameter
    f2 = crit2.wPrediction.stacked()
    assert all(p.weight.data_ptr() == f2[i].data_ptr() for i, p in enumerate(crit2.wPrediction.predictors))


def test_patch_install_swaps_reference_symbols(tmp_path, monkeypatch):
    """cpc_audio_b200.patch.install replaces the hot-path classes inside an importable `cpc` package (here a stub
    with the reference's module layout), which is all cpc/train.py needs to build the B200 modules."""
    import sys
    import types
    pkg = types.ModuleType("cpc"); pkg.__path__ = []
    model = types.ModuleType("cpc.model")
    crit_pkg = types.ModuleType("cpc.criterion"); crit_pkg.__path__ = []
    crit = types.ModuleType("cpc.criterion.criterion")
    tr = types.ModuleType("cpc.transformers")
    for name in ("ChannelNorm", "CPCEncoder", "CPCAR", "CPCModel"):
        setattr(model, name, object)
    for m in (crit_pkg, crit):
        m.CPCUnsupersivedCriterion = object
        m.PredictionNetwork = object
    tr.TransformerLayer = tr.buildTransformerAR = object
    for name, mod in (("cpc", pkg), ("cpc.model", model), ("cpc.criterion", crit_pkg), ("cpc.criterion.criterion", crit),
                      ("cpc.transformers", tr)):
        monkeypatch.setitem(sys.modules, name, mod)
    import cpc_audio_b200 as M
    from cpc_audio_b200 import patch
    patch.install(pkg)
    assert model.CPCEncoder is M.CPCEncoder and model.CPCAR is M.CPCAR and model.CPCModel is M.CPCModel
    assert crit.CPCUnsupersivedCriterion is M.CPCUnsupersivedCriterion
    assert crit_pkg.CPCUnsupersivedCriterion is M.CPCUnsupersivedCriterion
    assert tr.buildTransformerAR is M.buildTransformerAR
    # the constructor calls made by cpc/feature_loader.py:133-152 and cpc/train.py:31-40 work on the mirrors
    enc = model.CPCEncoder(256, "layerNorm")
    ar = model.CPCAR(256, 256, False, 1, mode="GRU", reverse=False)
    m = model.CPCModel(enc, ar)
    c = crit.CPCUnsupersivedCriterion(12, 256, 256, 128, mode=None, rnnMode="linear", dropout=False, nSpeakers=3,
                                      speakerEmbedding=0, sizeInputSeq=128)
    assert m.gEncoder.DOWNSAMPLING == 160 and c.nPredicts == 12


def test_transformer_heads_surface():
    """rnnMode='transformer' (the reference default, cpc_default_config.py:80): same state_dict keys / parameter count."""
    import cpc_audio_b200 as M
    crit = M.CPCUnsupersivedCriterion(12, 256, 256, 128, mode=None, rnnMode="transformer", dropout=False, speakerEmbedding=0,
                                      nSpeakers=0, sizeInputSeq=128)
    sd = crit.state_dict()
    pre = "wPrediction.predictors.3.0."
    for k, shape in {"multihead.Wq.weight": (256, 256), "multihead.Att.Krelpos": (32, 116), "multihead.Att.mask": (1, 116, 116),
                     "multihead.Att.z": (1, 116, 1), "ln_multihead.weight": (256,), "ffnetwork.lin1.weight": (2048, 256),
                     "ffnetwork.lin2.bias": (256,), "ln_ffnetwork.bias": (256,)}.items():
        assert tuple(sd[pre + k].shape) == shape, k
    assert len(sd) == 12 * 15 and sum(p.numel() for p in crit.parameters()) == 15813120
