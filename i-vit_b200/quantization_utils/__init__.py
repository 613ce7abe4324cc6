"""Drop-in replacement for the reference's ``models/quantization_utils`` package: the same seven
operator classes (reference ``models/quantization_utils/__init__.py:1``), backed by sm_100a kernels."""
from .ops import QuantLinear, QuantAct, QuantConv2d, QuantMatMul, IntLayerNorm, IntSoftmax, IntGELU
from .primitives import (SymmetricQuantFunction, batch_frexp, fixedpoint_mul, floor_ste, linear_quantize,
                         round_ste, symmetric_linear_quantization_params)

__all__ = ["QuantLinear", "QuantAct", "QuantConv2d", "QuantMatMul", "IntLayerNorm", "IntSoftmax", "IntGELU",
           "SymmetricQuantFunction", "batch_frexp", "fixedpoint_mul", "floor_ste", "linear_quantize", "round_ste",
           "symmetric_linear_quantization_params"]
