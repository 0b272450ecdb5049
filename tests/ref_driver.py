"""Test / bench infrastructure: drive the UNMODIFIED reference training loop (cpc/train.py:trainStep, :64-119) and its own
factories (feature_loader.getEncoder / getAR, train.getCriterion) - over the reference modules themselves, or over the B200
modules after ``cpc_audio_b200.patch.install``.  Needs the reference package (``oracle.ref_import``: /root/reference in the
authoring container, baseline/_ref on the GPU box); callers skip when it is absent.  Never imported by the product."""
from __future__ import annotations

import argparse
import contextlib
import io

import torch

from oracle.ref_import import import_reference

_REF_CLASSES = {}


def reference_or_none():
    ref = import_reference()
    if ref is not None and not _REF_CLASSES:
        # remember the reference's own classes so that install() / uninstall() can be toggled inside one process
        _REF_CLASSES.update({("model", n): getattr(ref.model, n) for n in ("ChannelNorm", "CPCEncoder", "CPCAR", "CPCModel")})
        import cpc.criterion.criterion as crit_mod
        _REF_CLASSES.update({("crit", n): getattr(crit_mod, n) for n in ("CPCUnsupersivedCriterion", "PredictionNetwork")})
        _REF_CLASSES.update({("tr", n): getattr(ref.transformers, n) for n in ("TransformerLayer", "buildTransformerAR")})
    return ref


def use_b200_modules(on: bool):
    """install() the B200 classes into the reference package, or put the reference's own classes back."""
    ref = reference_or_none()
    import cpc.criterion.criterion as crit_mod
    if on:
        import cpc_audio_b200.patch as patch
        patch.install(ref.package)
        return
    for (where, name), cls in _REF_CLASSES.items():
        if where == "model":
            setattr(ref.model, name, cls)
        elif where == "crit":
            setattr(ref.criterion, name, cls)
            setattr(crit_mod, name, cls)
        else:
            setattr(ref.transformers, name, cls)


def default_args(**over):
    """The argparse namespace cpc/train.py would build (cpc_default_config.py + the train.py flags this path reads)."""
    ref = reference_or_none()
    from cpc.cpc_default_config import set_default_cpc_config
    args = set_default_cpc_config(argparse.ArgumentParser()).parse_args([])
    args.supervised, args.pathPhone, args.CTC = False, None, False
    for k, v in over.items():
        setattr(args, k, v)
    return args


def build(args, b200: bool, seed=0, device="cuda", state=None):
    """train.py:307-337: encoder, context network, CPCModel, criterion, torch.optim.Adam - through the reference's factories."""
    ref = reference_or_none()
    use_b200_modules(b200)
    torch.manual_seed(seed)
    enc = ref.feature_loader.getEncoder(args)
    ar = ref.feature_loader.getAR(args)
    # `model.CPCModel` resolved through the module attribute, exactly as train.py:311 does
    model = ref.train.model.CPCModel(enc, ar)
    crit = ref.train.getCriterion(args, model.gEncoder.DOWNSAMPLING, 0, 0)
    if state is not None:
        model.load_state_dict(state[0], strict=False)
        crit.load_state_dict(state[1], strict=False)
    crit.to(device)
    model.to(device)
    g_params = list(crit.parameters()) + list(model.parameters())
    opt = torch.optim.Adam(g_params, lr=args.learningRate, betas=(args.beta1, args.beta2), eps=args.epsilon)
    # train.py:372-375 (nGPU = 1: one process per GPU)
    dev_ids = [torch.device(device).index or 0] if str(device).startswith("cuda") else None
    if dev_ids is not None:
        model_dp = torch.nn.DataParallel(model, device_ids=dev_ids).to(device)
        crit_dp = torch.nn.DataParallel(crit, device_ids=dev_ids).to(device)
    else:
        model_dp, crit_dp = model, crit
    return model, crit, model_dp, crit_dp, opt


def train_steps(model_dp, crit_dp, opt, batches, scheduler=None):
    """One call of the reference's trainStep per batch (so that the per-step losses are observable); returns the list of
    per-step mean losses (K,) as trainStep logs them."""
    ref = reference_or_none()
    out = []
    for x, label in batches:
        with contextlib.redirect_stdout(io.StringIO()):
            logs = ref.train.trainStep([(x, label)], model_dp, crit_dp, opt, scheduler, 10 ** 9)
        out.append(torch.as_tensor(logs["locLoss_train"]).clone())
    return out
