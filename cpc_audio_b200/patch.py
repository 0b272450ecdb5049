"""Patch the reference package so that an UNMODIFIED ``cpc/train.py`` builds the B200 modules.

``cpc/train.py:307-311`` resolves ``CPCEncoder`` / ``CPCAR`` through ``cpc.feature_loader.getEncoder/getAR``
(``from .model import CPCEncoder`` at call time), ``model.CPCModel`` and ``cr.CPCUnsupersivedCriterion`` through
module attributes (train.py:31,311) - so replacing those attributes before ``cpc.train.main(argv)`` is enough.

    python -m cpc_audio_b200.patch /path/to/CPC_audio/cpc/train.py --arMode GRU --rnnMode linear ...
"""
from __future__ import annotations

import runpy
import sys


def install(cpc_package=None):
    """Replace the hot-path classes inside an importable ``cpc`` package; returns the patched modules."""
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
