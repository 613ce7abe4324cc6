"""Load the reference's OWN model graphs (``models/vit_quant.py``, ``models/swin_quant.py``, ``models/layers_quant.py``,
``models/model_utils.py``) unchanged on top of this package's operator mirror.

The reference reaches its operators through one package, ``models/quantization_utils`` (``__init__.py:1``; imported
relatively at ``vit_quant.py:15``, ``swin_quant.py:11``, ``layers_quant.py:10``, ``model_utils.py:2``).  That import is
the plugin boundary: ``load_reference_models`` registers ``ivit_b200.quantization_utils`` under the reference package's
name in ``sys.modules`` BEFORE executing the reference's ``models/__init__.py``, so every ``from .quantization_utils
import ...`` in the reference files binds the sm_100a-backed classes.  No reference source is edited or copied into
this package; the files are executed from wherever the checkout lives (``/root/reference/models`` in the build
container, the staged ``baseline/_ref/models`` on a GPU box -- tools/fetch_ref.py).

``model_utils.freeze_model``'s exact-type test ``type(m) in [QuantAct]`` (model_utils.py:9,28) passes because the
``QuantAct`` in that namespace IS the mirror's class.

``mirror=False`` loads the reference with its own quantization_utils instead (the literal fp32-carrier implementation):
bench.py times it on the host cores as the CPU baseline; the product never runs through it.
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import sys
import types

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(_ROOT, "baseline", "_ref", "models")


def find_reference_models(path: str = None) -> str:
    """Directory of the reference's ``models`` package: explicit path, $IVIT_REFERENCE/models, the staged copy
    (baseline/_ref/models), /root/reference/models -- first that exists."""
    cands = [path, os.path.join(os.environ["IVIT_REFERENCE"], "models") if os.environ.get("IVIT_REFERENCE") else None,
             STAGED, "/root/reference/models"]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "vit_quant.py")):
            return c
    raise FileNotFoundError("reference models/ not found (run tools/fetch_ref.py where the reference checkout exists); "
                            "looked in: %s" % ", ".join(str(c) for c in cands if c))


def _tkinter_stub():
    # models/swin_quant.py:2 does `from tkinter import X` (unused); tkinter is not installed on headless boxes
    try:
        import tkinter  # noqa: F401
    except Exception:
        sys.modules["tkinter"] = types.SimpleNamespace(X="x")


_loaded = {}


def load_reference_models(path: str = None, package: str = None, mirror: bool = True, auto_accelerate: bool = False):
    """Import the reference ``models`` package from ``path`` (see ``find_reference_models``) as ``package`` and return
    the module: ``m.deit_tiny_patch16_224``, ``m.swin_tiny_patch4_window7_224``, ``m.freeze_model``, ...

    mirror=True   its ``quantization_utils`` is this package's sm_100a-backed mirror (the drop-in);
    mirror=False  the reference's own quantization_utils (CPU baseline only).
    auto_accelerate=True  (mirror only) the package's ``freeze_model`` / ``unfreeze_model`` -- the names
                  ``quant_train.py`` calls around its validation loop (:325-326, :273) -- additionally switch a whole
                  VisionTransformer / SwinTransformer that lives on a CUDA device to the fused engine
                  (``engine.accelerate``) and back: the reference's evaluation code then runs the fused path without
                  a single edit.  Freezing a sub-module, or a model on the CPU, behaves exactly as in the reference."""
    models_dir = find_reference_models(path)
    if auto_accelerate and not mirror:
        raise ValueError("auto_accelerate needs the mirror")
    if package is None:
        package = ("ivit_ref_models" if mirror else "ivit_ref_models_literal") + ("_auto" if auto_accelerate else "")
    key = (os.path.realpath(models_dir), package, mirror, auto_accelerate)
    if key in _loaded:
        return _loaded[key]
    if package in sys.modules:
        raise RuntimeError("module name %r is already taken" % package)
    _tkinter_stub()
    if mirror:
        from . import quantization_utils as qu
        sys.modules[package + ".quantization_utils"] = qu                      # <- the plugin boundary
        sys.modules[package + ".quantization_utils.quant_modules"] = qu.ops
        sys.modules[package + ".quantization_utils.quant_utils"] = qu.primitives
    spec = importlib.util.spec_from_file_location(package, os.path.join(models_dir, "__init__.py"),
                                                  submodule_search_locations=[models_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[package] = mod
    try:
        spec.loader.exec_module(mod)                                           # the reference's own models/__init__.py
    except Exception:
        for k in [k for k in sys.modules if k == package or k.startswith(package + ".")]:
            del sys.modules[k]
        raise
    if mirror:
        mod.quantization_utils = qu
        assert mod.QuantAct is qu.QuantAct and sys.modules[package + ".vit_quant"].QuantLinear is qu.QuantLinear
        if auto_accelerate:
            _install_auto_accelerate(mod, package)
    _loaded[key] = mod
    return mod


def _install_auto_accelerate(mod, package):
    """Wrap the reference's freeze_model / unfreeze_model (model_utils.py:5-40) in the loaded package's namespace."""
    from .engine import accelerate
    mu = sys.modules[package + ".model_utils"]
    ref_freeze, ref_unfreeze = mu.freeze_model, mu.unfreeze_model
    roots = (sys.modules[package + ".vit_quant"].VisionTransformer, sys.modules[package + ".swin_quant"].SwinTransformer)

    def _restore(model):
        if hasattr(model, "_ivit_forward_operator_level"):
            model.forward = model._ivit_forward_operator_level
            del model._ivit_forward_operator_level
            model._ivit_engine = None

    def freeze_model(model):
        ref_freeze(model)
        if isinstance(model, roots):
            _restore(model)                                   # weights / ranges may have changed since the last freeze
            p = next(model.parameters(), None)
            if p is not None and p.is_cuda:
                accelerate(model, p.device)

    def unfreeze_model(model):
        if isinstance(model, roots):
            _restore(model)
        ref_unfreeze(model)

    freeze_model.__doc__, unfreeze_model.__doc__ = ref_freeze.__doc__, ref_unfreeze.__doc__
    for ns in (mod, mu):
        ns.freeze_model, ns.unfreeze_model = freeze_model, unfreeze_model


@contextlib.contextmanager
def cuda_calls_are_noops():
    """The reference hard-codes ``.cuda()`` at six sites (quant_modules.py:356,440,494; quant_utils.py:88,174-175).
    Running its literal implementation on HOST tensors (CPU baseline) therefore needs ``Tensor.cuda`` to be the
    identity for the duration of the call -- also on a box that has a GPU."""
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig
