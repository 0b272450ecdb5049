"""cpc_audio_b200 - B200-native (sm_100a) CPC training-step hot path behind the reference's nn.Module surfaces.

    from cpc_audio_b200 import CPCEncoder, CPCAR, CPCModel, CPCUnsupersivedCriterion
    import cpc_audio_b200.patch; cpc_audio_b200.patch.install()   # then run the unmodified cpc/train.py

See DESIGN.md (path, kernels, rooflines) and INTEGRATION.md (how the reference binds to the C ABI).
"""
from .model import ChannelNorm, CPCEncoder, CPCAR, CPCModel  # noqa: F401
from .criterion import (BaseCriterion, PredictionNetwork, CPCUnsupersivedCriterion,  # noqa: F401
                        CPCUnsupervisedCriterion)
from .transformers import TransformerLayer, buildTransformerAR  # noqa: F401
from . import _lib  # noqa: F401

__all__ = ["ChannelNorm", "CPCEncoder", "CPCAR", "CPCModel", "BaseCriterion", "PredictionNetwork",
           "CPCUnsupersivedCriterion", "CPCUnsupervisedCriterion", "TransformerLayer", "buildTransformerAR"]
