"""Synthetic calibrated models (BASELINE.json workload: random-init weights of the named
architecture, no checkpoints / datasets available offline).

Weights come from ``synth.synth_parameters`` (deterministic per parameter name); activation
ranges come from a committed calibration table ``tests/golden/calib_<model>.json`` that was
produced by one unfrozen forward of the REFERENCE implementation on the same weights
(tests/golden/make_golden.py), so the frozen parameters are identical to the reference's --
"identical INT8 inputs and quantization params" (BASELINE.json north_star).
"""
from __future__ import annotations

import json
import os

import torch

from . import deit, swin
from .quantization_utils import QuantAct
from .synth import synth_parameters

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CALIB_DIR = os.path.join(_ROOT, "tests", "golden")

FACTORIES = {
    "deit_tiny_patch16_224": deit.deit_tiny_patch16_224,
    "deit_small_patch16_224": deit.deit_small_patch16_224,
    "deit_base_patch16_224": deit.deit_base_patch16_224,
    "vit_large_patch16_224": deit.vit_large_patch16_224,
    "swin_tiny_patch4_window7_224": swin.swin_tiny_patch4_window7_224,
    "swin_base_patch4_window7_224": swin.swin_base_patch4_window7_224,
}


def load_calibration(name: str) -> dict:
    with open(os.path.join(CALIB_DIR, "calib_%s.json" % name)) as f:
        return json.load(f)


def apply_calibration(model: torch.nn.Module, ranges: dict, strict: bool = True):
    """Set every QuantAct's (min_val, max_val) from ``ranges[name] = [min, max, bits]`` and freeze."""
    seen = set()
    for name, mod in model.named_modules():
        if type(mod).__name__ == "QuantAct":
            if name not in ranges:
                if strict:
                    raise KeyError("no calibration entry for QuantAct %r" % name)
                continue
            mn, mx, bits = ranges[name]
            if int(bits) != int(mod.activation_bit):
                raise ValueError("%s: calibration table has %d bits, module has %d" % (name, bits, mod.activation_bit))
            if isinstance(mod, QuantAct):
                mod.set_range(mn, mx)
            else:                                   # a reference QuantAct object
                mod.min_val = torch.tensor(float(mn), dtype=torch.float32)
                mod.max_val = torch.tensor(float(mx), dtype=torch.float32)
                mod.running_stat = False
            seen.add(name)
    return seen


def build_synthetic(name: str, seed: int = 0, check_weights: bool = True) -> torch.nn.Module:
    """The synthetic, calibrated, frozen model ``name`` (CPU tensors; feed it to pack.export_deit)."""
    cal = load_calibration(name)
    if cal["seed"] != seed:
        raise ValueError("calibration table was made for seed %d" % cal["seed"])
    model = FACTORIES[name]().eval()
    sha = synth_parameters(model, seed)
    if check_weights and sha != cal["weights_sha256"]:
        raise RuntimeError("synthetic weights differ from the ones the calibration table was made for "
                           "(torch RNG drift?): %s != %s" % (sha[:16], cal["weights_sha256"][:16]))
    apply_calibration(model, cal["ranges"])
    return model
