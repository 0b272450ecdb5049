"""Patch the reference package so that an UNMODIFIED ``cpc/train.py`` builds the B200 modules.

``cpc/train.py:307-311`` resolves ``CPCEncoder`` / ``CPCAR`` through ``cpc.feature_loader.getEncoder/getAR``
(``from .model import CPCEncoder`` at call time), ``model.CPCModel`` and ``cr.CPCUnsupersivedCriterion`` through
module attributes (train.py:31,311) - so replacing those attributes before ``cpc.train.main(argv)`` is enough.

    python -m cpc_audio_b200.patch /path/to/CPC_audio/cpc/train.py --arMode GRU --rnnMode linear ...
"""
from __future__ import annotations

import runpy
import sys


def install(cpc_package=None, adam=None):
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
    script = argv[0]
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(script))))
    install()
    sys.argv = argv
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
