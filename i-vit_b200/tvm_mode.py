"""TVM-semantics compatibility mode.

The reference's deployment tree restates the three integer row operators in TVM Relay
(``TVM_benchmark/models/layers.py``): ``quantized_layernorm`` (:329-350), ``shift_exp`` (:353-369),
``quantized_softmax`` (:372-386) and ``quantized_gelu`` (:389-404).  Their numerics differ from the PyTorch operators the
engines reproduce (``quant_modules.py``): int32 wrapping arithmetic, truncating divisions, ``n = 16`` / ``23`` with the
``(r >> 1) - x0`` exponent, no clamp on the sums, an 8-bit softmax by a wrapping cast.  The functions below keep the names
and argument meaning of ``layers.py`` -- tensors instead of Relay expressions -- and run as sm_100a kernels
(``csrc/ivit_tvm.cu``) through the C ABI; there is no CPU fallback.  They exist to cross-check a deployed TVM build's
intermediate tensors against this library's; the engines never call them.

TVM itself is not available in this image.  The operators are checked against ``oracle/tvm_semantics.py`` and against
``tests/golden/tvm_ops.npz`` -- vectors produced by the reference's own ``layers.py`` (unmodified) executed on a numpy
stand-in for the relay primitives (``tests/golden/relay_shim.py``): the structure of each operator is pinned to the
reference source, the integer semantics of the relay primitives (wrap, truncating division, ...) are taken from Relay's
documentation and are not pinned to a TVM run.
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import call, context, ptr


def x0_of(input_scale) -> int:
    """``relay.const(-1.0 / input_scale - 1, 'int32')`` (layers.py:357): a float converted to int32 truncates."""
    return int(np.array(-1.0 / float(input_scale) - 1).astype("int32"))


def _as_i32_rows(data: torch.Tensor):
    if not data.is_cuda:
        raise RuntimeError("ivit_b200.tvm_mode: needs a CUDA tensor (no CPU fallback)")
    x = data.to(torch.int32).contiguous()                       # relay.cast(data, 'int32')
    cols = x.shape[-1]
    return x, x.numel() // cols, cols


def quantized_softmax(data: torch.Tensor, input_scale, n: int = 16) -> torch.Tensor:
    """layers.py:372-386.  Integer tensor [..., cols] -> int8 probabilities (scale 2^-7)."""
    x, rows, cols = _as_i32_rows(data)
    out = torch.empty(x.shape, dtype=torch.int8, device=x.device)
    call("ivit_tvm_softmax", context(x.device), ptr(x), rows, cols, x0_of(input_scale), n, ptr(out))
    return out


def quantized_gelu(pre_data: torch.Tensor, input_scale, n: int = 23) -> torch.Tensor:
    """layers.py:389-404.  Integer tensor [..., cols] -> int32 ``pre_data * sigmoid_int``."""
    x, rows, cols = _as_i32_rows(pre_data)
    out = torch.empty(x.shape, dtype=torch.int32, device=x.device)
    call("ivit_tvm_gelu", context(x.device), ptr(x), rows, cols, x0_of(float(input_scale) * 1.702), n, ptr(out))
    return out


def quantized_layernorm(data: torch.Tensor, bias_int: torch.Tensor) -> torch.Tensor:
    """layers.py:329-350.  Integer tensor [..., C] and int32 ``bias_int`` [C] -> int32."""
    x, rows, cols = _as_i32_rows(data)
    b = bias_int.to(device=x.device, dtype=torch.int32).contiguous()
    if b.numel() != cols:
        raise ValueError("quantized_layernorm: bias_int must hold one value per channel")
    out = torch.empty(x.shape, dtype=torch.int32, device=x.device)
    call("ivit_tvm_layernorm", context(x.device), ptr(x), rows, cols, ptr(b), ptr(out))
    return out
