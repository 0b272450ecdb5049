"""Pin the CPU oracle (oracle/cpc_oracle.py) against outputs of the UNMODIFIED reference modules.

tests/golden/*.npz were produced by oracle/make_golden.py in the authoring container, where /root/reference is
importable; they hold losses, accuracies, sub-sampled z / c / gradients and the exact negative-sample indices.
"""
import numpy as np
import pytest
import torch

from oracle import cpc_oracle as O
from tests import helpers as Hh


@pytest.mark.parametrize("name", Hh.CASES + Hh.T_CASES + Hh.AR_CASES + Hh.TRAIN_CASES + Hh.WIDE_CASES)
def test_oracle_matches_reference_fixture(name):
    g, d, mp, cp, x, label, bi, si = Hh.load_case(name)
    ar_masks, head_masks = Hh.case_masks(g, d)
    assert int(bi.sum()) == int(g["bi_sum"]) and int(si.sum()) == int(g["si_sum"]), "seeded draws changed"
    ext = O.ext_indices_np(bi.numpy(), si.numpy(), d.B, d.N, d.W, d.S)
    assert np.array_equal(ext.astype(np.int32), g["ext_idx"]), "negative-sample indices must be bit-exact"
    r = Hh.oracle_run(d, mp, cp, x, bi, si, heads=Hh.case_heads(g), ar_masks=ar_masks, head_masks=head_masks)
    np.testing.assert_allclose(r["losses"].numpy(), g["losses"], rtol=0, atol=2e-5)
    np.testing.assert_array_equal(r["acc"].numpy(), g["acc"])
    np.testing.assert_allclose(Hh.subsample(r["z"]), g["z_sub"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(Hh.subsample(r["c"]), g["c_sub"], rtol=0, atol=2e-5)
    assert abs(r["z"].double().norm().item() - float(g["z_norm"])) < 1e-3 * float(g["z_norm"])
    for k, gr in r["grads"].items():
        ref = g[f"gsub.{k}"]
        scale = np.abs(ref).max() + 1e-12
        assert np.abs(Hh.subsample(gr, 512) - ref).max() / scale < 2e-3, k
        assert abs(gr.double().norm().item() - float(g[f"gnorm.{k}"])) <= 2e-3 * float(g[f"gnorm.{k}"]) + 1e-9, k


def test_index_properties():
    d = O.Dims(B=5, L=160 * 20, H=64, Har=64, K=3, N=7)
    bi, si = O.make_raw_indices(d, seed=9)
    ext = O.ext_indices_np(bi.numpy(), si.numpy(), d.B, d.N, d.W, d.S)
    assert ext.shape == (d.B, d.N, d.W) and ext.min() >= 0 and ext.max() < d.B * d.S
    w = np.arange(d.W)[None, None, :]
    assert np.all(ext % d.S != w), "a negative never sits on the anchor's own time step (criterion.py:186-197)"


def test_gru_restatement_matches_torch_gru():
    d = O.Dims(B=3, L=160 * 9, H=64, Har=128, nLayers=2)
    mp, _ = O.make_params(d, seed=5)
    z = torch.randn(d.B, d.S, d.H, generator=torch.Generator().manual_seed(1))
    gru = torch.nn.GRU(d.H, d.Har, num_layers=2, batch_first=True)
    gru.load_state_dict({k.replace("gAR.baseNet.", ""): v for k, v in mp.items() if k.startswith("gAR.")})
    ref, hT = gru(z)
    out, h = O.gru_forward(z, mp, 2)
    assert (ref - out).abs().max() < 1e-5 and (hT - h).abs().max() < 1e-5


def test_einsum_scoring_equals_materialised():
    g, d, mp, cp, x, label, bi, si = Hh.load_case("small")
    c, z = O.model_forward(x, mp, d.nLayers)
    l1, a1, lg = O.criterion_forward(c, z, cp, bi, si, d.K, d.N, materialize=True)
    l2, a2, _ = O.criterion_forward(c, z, cp, bi, si, d.K, d.N, materialize=False)
    assert (l1 - l2).abs().max() < 1e-4
    # accuracies may differ only on near-tie rows (the true positive can be drawn as a negative: criterion.py:186-197)
    assert ((a1 - a2).abs() <= Hh.acc_tolerance(lg, d)).all()


@pytest.mark.parametrize("name", Hh.FEATURE_CASES)
def test_feature_path_restatement_matches_reference_buildFeature(name):
    """oracle.feature_forward vs the fixture written from the UNMODIFIED cpc/feature_loader.py:buildFeature + FeatureModule
    (chunks of maxSizeSeq samples, last chunk of arbitrary length, hidden state carried by keepHidden)."""
    g = np.load(f"{Hh.GOLDEN}/{name}.npz")
    nl, n, chunk, H = int(g["nLayers"]), int(g["n"]), int(g["chunk"]), int(g["H"])
    d = O.Dims(B=1, L=chunk, H=H, Har=H, nLayers=nl)
    mp, _ = O.make_params(d, seed=21, ar=str(g["ar"]))
    seq = torch.randn(n, generator=torch.Generator().manual_seed(22)) * 0.1
    feat = O.feature_forward(seq, mp, nl, max_size_seq=chunk, keep_hidden=True)
    assert list(feat.shape) == list(g["shape"])
    np.testing.assert_allclose(Hh.subsample(feat), g["feat_sub"], rtol=0, atol=2e-5)


def test_lstm_restatement_matches_torch_lstm():
    d = O.Dims(B=3, L=160 * 9, H=64, Har=128, nLayers=2)
    mp, _ = O.make_params(d, seed=5, ar="LSTM")
    z = torch.randn(d.B, d.S, d.H, generator=torch.Generator().manual_seed(1))
    lstm = torch.nn.LSTM(d.H, d.Har, num_layers=2, batch_first=True)
    lstm.load_state_dict({k.replace("gAR.baseNet.", ""): v for k, v in mp.items() if k.startswith("gAR.")})
    ref, (hT, cT) = lstm(z)
    out, (h, c) = O.lstm_forward(z, mp, 2)
    assert (ref - out).abs().max() < 1e-5 and (hT - h).abs().max() < 1e-5 and (cT - c).abs().max() < 1e-5
