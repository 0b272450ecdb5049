"""Input pipeline: the sampler restatements against the reference's own samplers (bit-exact batch index
lists under the same seeds, cpc/dataset.py:317-408) and, on the GPU, the window gather against AudioBatchData.__getitem__
(dataset.py:185-202) restated in numpy."""
import random

import numpy as np
import pytest
import torch

from cpc_audio_b200 import feeder as F
from tests import ref_driver as R


def _bounds(seed, n_intervals, window):
    g = torch.Generator().manual_seed(seed)
    sizes = (torch.randint(0, 9, (n_intervals,), generator=g) * window + torch.randint(0, window, (n_intervals,), generator=g)).tolist()
    return [0] + list(np.cumsum(sizes))


@pytest.mark.skipif(R.reference_or_none() is None, reason="reference package not present")
@pytest.mark.parametrize("offset", [0, 137])
def test_samplers_match_reference_bit_exact(offset):
    import cpc.dataset as D
    W, B = 640, 8
    bounds = [int(v) for v in _bounds(3, 25, W)]
    n = bounds[-1]
    torch.manual_seed(11); random.seed(12)
    ref = list(iter(D.SameSpeakerSampler(B, bounds, W, offset)))
    torch.manual_seed(11); random.seed(12)
    assert F.same_speaker_batches(B, bounds, W, offset) == ref
    torch.manual_seed(13)
    ref = list(iter(torch.utils.data.BatchSampler(D.UniformAudioSampler(n, W, offset), B, True)))
    torch.manual_seed(13)
    assert F.uniform_batches(B, n, W, offset) == ref
    assert F.sequential_batches(B, n, W, offset) == list(iter(D.SequentialSampler(n, W, offset, B)))


def test_sampler_properties():
    W, B = 320, 4
    bounds = [int(v) for v in _bounds(5, 12, W)]
    torch.manual_seed(1); random.seed(2)
    bl = F.same_speaker_batches(B, bounds, W, 50)
    seen = set()
    for b in bl:
        iv = {int(np.searchsorted(bounds, i, side="right")) - 1 for i in b}
        assert len(iv) == 1, "a batch never mixes speakers (dataset.py:385-396)"
        (k,) = iv
        assert all(bounds[k] <= i and i + W <= bounds[k + 1] for i in b)
        assert not (set(b) & seen)
        seen |= set(b)
    assert all(len(b) == B for b in F.uniform_batches(B, bounds[-1], W, 0))


@pytest.mark.gpu
@pytest.mark.parametrize("sampling", ["samespeaker", "uniform", "sequential"])
def test_feeder_batches_equal_getitem(sampling, built_lib):
    """Every batch of a pass equals the reference's __getitem__ + collate on the same indices: data[idx : idx + L] and the
    index of the speaker interval containing idx - bit-exact (it is a copy)."""
    W, B = 20480, 8
    bounds = [int(v) for v in _bounds(7, 9, W)]
    n = bounds[-1]
    host = torch.randn(n, generator=torch.Generator().manual_seed(8))
    pack = F.ResidentPack.from_host(host, bounds, bounds, chunk=100003)   # odd chunk size: several staged pieces
    feed = F.WindowFeeder(pack, B, W, sampling=sampling, random_offset=True, drop_ragged=False)
    torch.manual_seed(21); random.seed(22)
    want = feed.batches()
    torch.manual_seed(21); random.seed(22)
    got = list(feed)
    feed.check()
    assert len(got) == len(want) > 0
    data = host.numpy()
    for idxs, (x, label) in zip(want, got):
        assert x.shape == (len(idxs), 1, W) and x.is_cuda and label.dtype == torch.int64
        ref = np.stack([data[i:i + W] for i in idxs])[:, None, :]
        assert np.array_equal(x.cpu().numpy(), ref)
        assert label.cpu().tolist() == [int(np.searchsorted(bounds, i, side="right")) - 1 for i in idxs]


@pytest.mark.gpu
def test_out_of_range_start_is_flagged(built_lib):
    pack = F.ResidentPack.from_host(torch.zeros(50000), [0, 50000])
    starts = torch.tensor([0, 40000], dtype=torch.int64, device="cuda")
    _, _, err = pack.gather(starts, 20480)
    assert int(err.item()) == 1
