"""Locate and import the UNMODIFIED reference package (facebookresearch/CPC_audio).  *** TEST INFRASTRUCTURE ONLY ***

Used by oracle/make_golden.py (fixture generation), the drop-in tests (the reference's own ``trainStep`` driving the B200
modules) and ``bench.py --impl reference`` (the reference's CPU path, ``cpu_baseline.kind = "reference"``).  The product
package never imports this module.

Where the reference lives: ``/root/reference`` in the authoring container; ``baseline/_ref`` (a pip --target install of
the same tree, git-ignored, shipped with the snapshot) on the GPU box.  Import recipe: stub the two
third-party modules the package imports but this path never calls (``progressbar``, ``soundfile``) and alias the bare
``import transformers`` of criterion.py:83 to ``cpc.transformers``.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from collections import namedtuple

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = ("/root/reference", os.path.join(REPO, "baseline", "_ref"))

Ref = namedtuple("Ref", "root model criterion transformers train feature_loader package")
_cached = None


def find_reference(explicit=None):
    for root in ([explicit] if explicit else []) + list(CANDIDATES):
        if root and os.path.isfile(os.path.join(root, "cpc", "model.py")) and os.path.isfile(os.path.join(root, "cpc", "train.py")):
            return root
    return None


def import_reference(explicit=None):
    """Returns a Ref namedtuple of the reference's modules, or None when the reference is not present."""
    global _cached
    if _cached is not None:
        return _cached
    root = find_reference(explicit)
    if root is None:
        return None
    for name in ("progressbar", "soundfile"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except ImportError:
                sys.modules[name] = types.ModuleType(name)
    if root not in sys.path:
        sys.path.insert(0, root)
    saved_tr = sys.modules.get("transformers")
    import cpc  # noqa: F401
    import cpc.model as ref_model
    import cpc.transformers as ref_tr
    sys.modules["transformers"] = ref_tr  # criterion.py:83 does a bare `from transformers import buildTransformerAR`
    import cpc.criterion as ref_crit
    import cpc.feature_loader as ref_fl
    import cpc.train as ref_train
    if saved_tr is not None:
        sys.modules["transformers"] = saved_tr
    _cached = Ref(root, ref_model, ref_crit, ref_tr, ref_train, ref_fl, cpc)
    return _cached
