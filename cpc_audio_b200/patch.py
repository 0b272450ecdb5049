"""Patch the reference package so that an UNMODIFIED ``cpc/train.py`` builds the B200 modules.

``cpc/train.py:307-311`` resolves ``CPCEncoder`` / ``CPCAR`` through ``cpc.feature_loader.getEncoder/getAR``
(``from .model import CPCEncoder`` at call time), ``model.CPCModel`` and ``cr.CPCUnsupersivedCriterion`` through
module attributes (train.py:31,311) - so replacing those attributes before ``cpc.train.main(argv)`` is enough.

    python -m cpc_audio_b200.patch /path/to/CPC_audio/cpc/train.py --arMode GRU --rnnMode linear ...
"""
from __future__ import annotations

import runpy
import sys


_TORCH_DP = None


def _single_device_dataparallel():
    """A ``torch.nn.DataParallel`` SUBCLASS (isinstance checks such as feature_loader.py:194 keep working, ``.module`` and the
    'module.'-prefixed state_dict keys are the parent's) whose forward skips scatter / replicate / gather when there is ONE
    device and the tensor arguments already live on it - which is exactly what the parent computes in that case
    (``return self.module(*inputs[0], **kwargs[0])`` after an identity scatter), minus ~0.27 ms of Python per training step
    and the Scatter / Gather autograd nodes (two 8 MB copies in the backward pass)."""
    import torch
    global _TORCH_DP
    if _TORCH_DP is None:
        _TORCH_DP = torch.nn.DataParallel
    base = _TORCH_DP

    class DataParallel(base):
        def forward(self, *inputs, **kwargs):
            if len(self.device_ids) == 1:
                dev = self.src_device_obj
                if all((not isinstance(t, torch.Tensor)) or t.device == dev for t in inputs) and \
                        all((not isinstance(t, torch.Tensor)) or t.device == dev for t in kwargs.values()):
                    return self.module(*inputs, **kwargs)
            return super().forward(*inputs, **kwargs)

    DataParallel.__qualname__ = "DataParallel"
    return DataParallel


def install_dataparallel():
    """Put the single-device fast path in place of torch.nn.DataParallel (see install(dataparallel=True))."""
    import torch
    dp = _single_device_dataparallel()
    torch.nn.DataParallel = dp
    torch.nn.parallel.DataParallel = dp


def uninstall_dataparallel():
    import torch
    if _TORCH_DP is not None:
        torch.nn.DataParallel = _TORCH_DP
        torch.nn.parallel.DataParallel = _TORCH_DP


def install(cpc_package=None, adam=None, dataparallel=None):
    """Replace the hot-path classes inside an importable ``cpc`` package; returns the patched modules.

    ``adam=True`` (or env ``CPC_B200_PATCH_ADAM=1``) additionally puts ``cpc_audio_b200.optim.Adam`` in place of
    ``torch.optim.Adam``: train.py:335 then builds the flat fused optimizer (one kernel per step instead of ~10 foreach
    launches over 36 tensors; no per-parameter AccumulateGrad nodes) - still without touching train.py.  Opt-in because
    it reaches outside the ``cpc`` package."""
    import os
    if adam is None:
        adam = os.environ.get("CPC_B200_PATCH_ADAM", "0") == "1"
    if adam:
        install_adam()
    if dataparallel is None:
        dataparallel = os.environ.get("CPC_B200_PATCH_DATAPARALLEL", "0") == "1"
    if dataparallel:  # train.py:372-375 with nGPU = 1 (one process per GPU): the wrapper becomes a true pass-through
        install_dataparallel()
    from . import criterion as our_crit
    from . import model as our_model
    if cpc_package is None:
        import cpc as cpc_package  # noqa: F401
    import importlib
    ref_model = importlib.import_module(cpc_package.__name__ + ".model")
    ref_crit_pkg = importlib.import_module(cpc_package.__name__ + ".criterion")
    ref_crit = importlib.import_module(cpc_package.__name__ + ".criterion.criterion")
    for name in ("ChannelNorm", "CPCEncoder", "CPCAR", "CPCModel"):
        setattr(ref_model, name, getattr(our_model, name))
    for mod in (ref_crit_pkg, ref_crit):
        setattr(mod, "CPCUnsupersivedCriterion", our_crit.CPCUnsupersivedCriterion)
        setattr(mod, "PredictionNetwork", our_crit.PredictionNetwork)
    # --arMode transformer: feature_loader.getAR does `from .transformers import buildTransformerAR` at call time
    from . import transformers as our_tr
    ref_tr = importlib.import_module(cpc_package.__name__ + ".transformers")
    for name in ("TransformerLayer", "buildTransformerAR"):
        setattr(ref_tr, name, getattr(our_tr, name))
    return ref_model, ref_crit


def install_adam():
    """Put cpc_audio_b200.optim.Adam in place of torch.optim.Adam (see install(adam=True))."""
    import torch
    from . import optim as our_optim
    torch.optim.Adam = our_optim.Adam


def uninstall_adam():
    import torch
    from . import optim as our_optim
    torch.optim.Adam = our_optim._TORCH_ADAM


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit(__doc__)
    fast = False
    if argv[0] == "--fast":
        fast, argv = True, argv[1:]
    script = argv[0]
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(script))))
    install(adam=True if fast else None, dataparallel=True if fast else None)
    sys.argv = argv
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
