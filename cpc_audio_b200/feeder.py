"""Input pipeline for the B200 training step: batches cut on the GPU out of an HBM-resident pack.

The reference feeds ``trainStep`` from ``AudioBatchData.getDataLoader`` (cpc/dataset.py:227-262): a torch ``DataLoader`` with
``numWorkers=0`` (cpc/train.py:181-185) that slices every window on the host (``__getitem__``, dataset.py:185-202), collates
them and copies the batch with ``.cuda()`` (train.py:81) - a few hundred windows per second, against the > 40 000 windows/s
one B200 consumes.  Here the pack (``AudioBatchData.data``: one 1-D float tensor, <= 16 GB) is uploaded ONCE to HBM - in
pinned, double-buffered chunks on a side stream, so the upload of the next pack overlaps training on the current one - and
a batch is one small kernel (``cpcb200_gather_windows``) that gathers B windows by start index and looks up their speaker
labels.  The batch index lists come from restatements of the reference's samplers that consume ``torch`` / ``random``
generators exactly as the reference does, so the same seeds give the same batches (tests/test_feeder.py).

``WindowFeeder`` is an iterable of ``(batchData (B,1,L) cuda, label (B) cuda)`` - exactly what the loop of cpc/train.py:78-82
expects from its dataLoader (the ``.cuda()`` calls there become no-ops).
"""
from __future__ import annotations

import ctypes
import random

import torch

from . import _lib as L


# ---------------------------------------------------------------------------------------------------------------------
# samplers (host-side index arithmetic; restated from cpc/dataset.py so that the generator consumption is identical)
# ---------------------------------------------------------------------------------------------------------------------
def same_speaker_batches(batch_size, sampling_intervals, size_window, offset):
    """cpc/dataset.py:361-408 (SameSpeakerSampler): every batch holds windows of ONE interval (speaker / sequence).
    One ``torch.randperm`` per non-empty interval in interval order, then ``random.shuffle`` of the batch list."""
    if sampling_intervals[0] != 0:
        raise AttributeError("Sampling intervals should start at zero")
    n = len(sampling_intervals) - 1
    sizes = [(sampling_intervals[i + 1] - sampling_intervals[i]) // size_window for i in range(n)]
    if offset > 0:
        sizes = [max(0, x - 1) for x in sizes]
    order = [(i, torch.randperm(v).tolist()) for i, v in enumerate(sizes) if v > 0]
    batches = []
    for i, perm in order:
        for lo in range(0, sizes[i], batch_size):
            batches.append([offset + x * size_window + sampling_intervals[i] for x in perm[lo:lo + batch_size]])
    random.shuffle(batches)
    return batches


def uniform_batches(batch_size, data_size, size_window, offset):
    """cpc/dataset.py:317-335 (UniformAudioSampler) under ``BatchSampler(sampler, batchSize, drop_last=True)`` (:222-223)."""
    n = data_size // size_window - (1 if offset > 0 else 0)
    idx = (offset + size_window * torch.randperm(n)).tolist()
    return [idx[i:i + batch_size] for i in range(0, len(idx) - batch_size + 1, batch_size)]


def sequential_batches(batch_size, data_size, size_window, offset):
    """cpc/dataset.py:338-358 (SequentialSampler): batch element j walks the j-th of batchSize equal slices of the pack."""
    n = (data_size // size_window) // batch_size - (1 if offset > 0 else 0)
    starts = [x * (data_size // batch_size) for x in range(batch_size)]
    return [[offset + size_window * i + s for s in starts] for i in range(n)]


def make_batches(kind, batch_size, data_size, size_window, offset, speaker_bounds=None, seq_bounds=None):
    """AudioBatchData.getBaseSampler (cpc/dataset.py:211-223)."""
    if kind == "samespeaker":
        return same_speaker_batches(batch_size, speaker_bounds, size_window, offset)
    if kind == "samesequence":
        return same_speaker_batches(batch_size, seq_bounds, size_window, offset)
    if kind == "sequential":
        return sequential_batches(batch_size, data_size, size_window, offset)
    return uniform_batches(batch_size, data_size, size_window, offset)


# ---------------------------------------------------------------------------------------------------------------------
class ResidentPack:
    """One audio pack in HBM: ``data`` (n_samples fp32) + the speaker / sequence interval bounds (int64, first = 0)."""

    CHUNK = 16 << 20  # samples per staged chunk (64 MB of pinned memory per staging buffer)

    def __init__(self, n_samples, speaker_bounds, seq_bounds=None, device="cuda"):
        self.device = torch.device(device)
        self.n = int(n_samples)
        self.data = torch.empty(self.n, dtype=torch.float32, device=self.device)
        self.speaker_bounds_host = [int(v) for v in speaker_bounds]
        self.seq_bounds_host = [int(v) for v in seq_bounds] if seq_bounds is not None else None
        self.speaker_bounds = torch.tensor(self.speaker_bounds_host, dtype=torch.int64, device=self.device)
        self.ready = torch.cuda.Event()
        self.ready.record(torch.cuda.current_stream(self.device))

    @classmethod
    def from_host(cls, data, speaker_bounds, seq_bounds=None, device="cuda", stream=None, chunk=None):
        """Upload a host pack (``AudioBatchData.data``) through two pinned staging buffers on `stream` (a side stream by
        default): the host-side copy into pinned memory of chunk i+1 overlaps the DMA of chunk i, and the whole upload
        overlaps whatever the compute stream is doing.  ``pack.ready`` is recorded when the data is in HBM."""
        pack = cls(data.numel(), speaker_bounds, seq_bounds, device)
        dev = pack.device
        stream = stream or torch.cuda.Stream(device=dev)
        chunk = int(chunk or cls.CHUNK)
        src = data.reshape(-1)
        if src.is_cuda:
            with torch.cuda.stream(stream):
                pack.data.copy_(src, non_blocking=True)
                pack.ready.record(stream)
            return pack
        if src.is_pinned():
            with torch.cuda.stream(stream):
                pack.data.copy_(src, non_blocking=True)
                pack.ready.record(stream)
            return pack
        stage = [torch.empty(min(chunk, pack.n), dtype=torch.float32).pin_memory() for _ in range(2)]
        freed = [torch.cuda.Event(), torch.cuda.Event()]
        with torch.cuda.stream(stream):
            for i, lo in enumerate(range(0, pack.n, chunk)):
                hi = min(pack.n, lo + chunk)
                s = i & 1
                if i >= 2:
                    freed[s].synchronize()  # the DMA that last read this staging buffer has finished
                stage[s][:hi - lo].copy_(src[lo:hi])
                pack.data[lo:hi].copy_(stage[s][:hi - lo], non_blocking=True)
                freed[s].record(stream)
            pack.ready.record(stream)
        pack._stage = stage  # keep the pinned buffers alive until the copies have run
        return pack

    def gather(self, starts, size_window, out=None, labels=None, err=None):
        """(B,1,L) windows + (B) speaker labels for the int64 device tensor `starts` - cpcb200_gather_windows."""
        B = starts.numel()
        dev = self.device
        if out is None:
            out = torch.empty(B, 1, size_window, dtype=torch.float32, device=dev)
        if labels is None:
            labels = torch.empty(B, dtype=torch.int64, device=dev)
        if err is None:
            err = torch.zeros(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            L.check(L.lib().cpcb200_gather_windows(L.ptr(self.data), self.n, L.ptr(starts), B, int(size_window), L.ptr(out),
                                                   L.ptr(self.speaker_bounds), self.speaker_bounds.numel(), L.ptr(labels),
                                                   L.ptr(err), L.stream_ptr(dev)), "gather_windows")
        return out, labels, err


class WindowFeeder:
    """Iterable of ``(batchData (B,1,L), label (B))`` device tensors over one pass of a resident pack - the drop-in for the
    ``AudioLoader`` that cpc/train.py:78 iterates (dataset.py:282-314).

    The index lists of the WHOLE pass are built once on the host (a few thousand int64 per thousand batches), uploaded as one
    tensor, and every batch is one gather kernel on the compute stream: nothing on the host scales with the batch."""

    def __init__(self, pack, batch_size, size_window, sampling="samespeaker", random_offset=True, drop_ragged=True):
        self.pack, self.B, self.L = pack, int(batch_size), int(size_window)
        self.sampling, self.random_offset = sampling, random_offset
        self.drop_ragged = drop_ragged
        self._err = torch.zeros(1, dtype=torch.int32, device=pack.device)

    def batches(self):
        # AudioBatchData.getDataLoader.samplerCall (dataset.py:255-258): a fresh random offset per pass
        offset = random.randint(0, self.L // 2) if self.random_offset else 0
        bl = make_batches(self.sampling, self.B, self.pack.n, self.L, offset, self.pack.speaker_bounds_host, self.pack.seq_bounds_host)
        if self.drop_ragged:  # fixed-shape steps (CUDA graph replay): keep the full batches only
            bl = [b for b in bl if len(b) == self.B]
        return bl

    def __iter__(self):
        bl = self.batches()
        dev = self.pack.device
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(self.pack.ready)
        full = [b for b in bl if len(b) == self.B]
        table = torch.tensor(full, dtype=torch.int64).pin_memory().to(dev, non_blocking=True) if full else None
        k = 0
        for b in bl:
            if len(b) == self.B:
                starts = table[k]
                k += 1
            else:
                starts = torch.tensor(b, dtype=torch.int64, device=dev)
            x, label, _ = self.pack.gather(starts, self.L, err=self._err)
            yield x, label

    def __len__(self):
        return self.pack.n // (self.L * self.B)

    def check(self):
        """Raise if any window start of the pass was out of range (synchronises)."""
        if int(self._err.item()) != 0:
            raise RuntimeError("cpc_audio_b200.WindowFeeder: a window start was outside the resident pack")
