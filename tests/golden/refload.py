"""Loader for the UNMODIFIED reference (zkkli/I-ViT) -- used only in this container to
generate / re-check golden vectors.  /root/reference does not exist on the GPU box, so
everything that imports this module is skipped there (``have_reference()``).

Shims (no source edits): a ``tkinter`` stub (models/swin_quant.py:2
imports ``from tkinter import X``) and, on a GPU-less host, ``Tensor.cuda`` -> identity
(six hard-coded ``.cuda()`` sites: quant_modules.py:356,440,494; quant_utils.py:88,174-175).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("IVIT_REFERENCE", "/root/reference")


def have_reference() -> bool:
    return os.path.isdir(os.path.join(REF, "models", "quantization_utils"))


_models = None


def load():
    """Import the reference ``models`` package (cached)."""
    global _models
    if _models is None:
        if not have_reference():
            raise RuntimeError("reference checkout not present at %s" % REF)
        sys.modules.setdefault("tkinter", types.SimpleNamespace(X="x"))
        if not torch.cuda.is_available():
            torch.Tensor.cuda = lambda self, *a, **k: self
        sys.path.insert(0, REF)
        try:
            import models  # noqa: the reference's package
        finally:
            sys.path.remove(REF)
        _models = models
    return _models


def add_exact_carrier_hooks(model):
    """SURVEY.md section 8c candidate 2: the reference's own modules, with the input of
    IntSoftmax / IntGELU / IntLayerNorm snapped to an exact fp64 carrier so that the
    module's own ``x / scaling_factor`` returns the exact integer."""
    m = load()
    qu = m.quantization_utils.quant_modules
    handles = []

    def pre(mod, inp):
        xq = (inp[0].double() / inp[1].double()).round()
        return (xq * inp[1].double(), inp[1]) + tuple(inp[2:])

    def post(mod, inp, out):
        if isinstance(mod, qu.IntSoftmax):
            return (out[0].float(), out[1])
        return None

    for mod in model.modules():
        if isinstance(mod, (qu.IntSoftmax, qu.IntGELU, qu.IntLayerNorm)):
            handles.append(mod.register_forward_pre_hook(pre))
            handles.append(mod.register_forward_hook(post))
    return handles


def capture_integers(model, x):
    """Run ``model(x)`` and return (logits, {module_name: int64 ndarray of its output}).
    Integers are read as round(out / sf) in fp64 (sf broadcast on the last dim, or
    (1,C,1,1) for the conv)."""
    m = load()
    qu = m.quantization_utils.quant_modules
    kinds = (qu.QuantAct, qu.QuantLinear, qu.QuantConv2d, qu.QuantMatMul,
             qu.IntLayerNorm, qu.IntSoftmax, qu.IntGELU)
    cap = {}
    handles = []

    def mk(name):
        def hook(mod, inp, out):
            t, sf = out
            q = (t.double() / sf.double()).round()
            cap[name] = q.to(torch.int64).numpy().copy()
        return hook

    for name, mod in model.named_modules():
        if isinstance(mod, kinds):
            handles.append(mod.register_forward_hook(mk(name)))
    with torch.no_grad():
        y = model(x)
    for h in handles:
        h.remove()
    return y, cap


def calibration_table(model):
    """{QuantAct module name: (min_val, max_val, bits)} as exact fp32 values."""
    m = load()
    qu = m.quantization_utils.quant_modules
    tab = {}
    for name, mod in model.named_modules():
        if isinstance(mod, qu.QuantAct):
            mn = float(torch.as_tensor(mod.min_val).float().reshape(-1)[0])
            mx = float(torch.as_tensor(mod.max_val).float().reshape(-1)[0])
            tab[name] = [np.float32(mn).item(), np.float32(mx).item(), int(mod.activation_bit)]
    return tab
