"""Drop-in mirror of cpc/transformers.py:{TransformerLayer, buildTransformerAR} (the ``--arMode transformer`` context
network, cpc/feature_loader.py:138-142, and the building block of the ``--rnnMode transformer`` prediction heads).

Same constructor arguments and state_dict keys (``0.multihead.Wq.weight`` ... ``0.ln_ffnetwork.bias``, buffers
``0.multihead.Att.z`` / ``.mask``) as the reference; the layer itself runs in libcpc_b200.so
(``cpcb200_tlayer_fwd / _bwd``).  ``abspos=True`` (StaticPositionEmbedding instead of relative positions) is outside
the accelerated path and raises.
"""
from __future__ import annotations

import torch.nn as nn

from .criterion import _TransformerLayer


class TransformerLayer(_TransformerLayer):
    """cpc/transformers.py:98-111."""

    def __init__(self, sizeSeq=32, dmodel=512, dff=2048, dropout=0.1, nheads=8, abspos=False, compute_dtype=None):
        if abspos:
            raise NotImplementedError("cpc_audio_b200: abspos=True (absolute position embedding) is outside the accelerated "
                                      "hot path; the reference default is relative positions (abspos=False)")
        if sizeSeq > 128:
            raise NotImplementedError("cpc_audio_b200: TransformerLayer supports sequences of at most 128 frames")
        super().__init__(sizeSeq, dmodel, dff=dff, nheads=nheads, dropout=dropout, compute_dtype=compute_dtype)


def buildTransformerAR(dimEncoded, nLayers, sizeSeq, abspos, compute_dtype=None):
    """cpc/transformers.py:129-139."""
    if abspos:
        raise NotImplementedError("cpc_audio_b200: abspos=True is outside the accelerated hot path")
    return nn.Sequential(*[TransformerLayer(sizeSeq=sizeSeq, dmodel=dimEncoded, abspos=abspos, compute_dtype=compute_dtype)
                           for _ in range(nLayers)])
