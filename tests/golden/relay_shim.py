"""A numpy stand-in for the handful of ``tvm.relay`` functions that the reference's integer row operators are written
with (TVM_benchmark/models/layers.py:329-404), so that the reference's OWN expressions can be executed in this image,
where TVM is not installed.  Used by make_tvm_golden.py only (golden-vector generation); never imported by the product.

What it pins and what it does not: the STRUCTURE of the operators (which shifts, which constants, which order) comes from
the unmodified reference source; the MEANING of each Relay primitive on integer tensors comes from this file and is the
documented Relay / TIR behaviour -- two's-complement wrap of int32 add / sub / mul, ``a / b`` = truncating division for
signed and floor division for unsigned integers, arithmetic ``right_shift``, ``mean`` = ``sum / count`` in the tensor's
dtype, ``cast`` keeping the low bits, ``const(float, 'int32')`` converting through numpy (truncation).  Where the backend
decides (division by zero, shift counts >= 32) the conventions of oracle/tvm_semantics.py are used.
"""
import sys
import types

import numpy as np

_RANGE = {"int8": (8, True), "int16": (16, True), "int32": (32, True), "uint32": (32, False), "int64": (64, True)}


def _wrap(v, dtype):
    bits, signed = _RANGE[dtype]
    if bits == 64:
        return np.asarray(v, np.int64)
    v = np.asarray(v, np.int64) % (1 << bits)
    if signed:
        v = np.where(v >= (1 << (bits - 1)), v - (1 << bits), v)
    return v.astype(np.int64)


class Expr:
    def __init__(self, v, dtype):
        self.dtype = dtype
        self.v = _wrap(v, dtype)

    def _other(self, o):
        assert isinstance(o, Expr), "the reference only combines relay expressions"
        assert o.dtype == self.dtype, (self.dtype, o.dtype)
        return o.v

    def __add__(self, o): return Expr(self.v + self._other(o), self.dtype)
    def __sub__(self, o): return Expr(self.v - self._other(o), self.dtype)
    def __mul__(self, o): return Expr(self.v * self._other(o), self.dtype)
    def __neg__(self): return Expr(-self.v, self.dtype)

    def __truediv__(self, o):
        b = np.broadcast_to(self._other(o), np.broadcast(self.v, o.v).shape)
        a = np.broadcast_to(self.v, b.shape)
        safe = np.where(b == 0, 1, b)
        if _RANGE[self.dtype][1]:
            q = np.abs(a) // np.abs(safe) * np.sign(a) * np.sign(safe)      # truncation toward zero
        else:
            q = a // safe
        return Expr(np.where(b == 0, 0, q), self.dtype)


def const(value, dtype="int32"):
    return Expr(np.array(value).astype(dtype).astype(np.int64), dtype)


def cast(data, dtype):
    return Expr(data.v, dtype)


def right_shift(a, b):
    return Expr(a.v >> a._other(b), a.dtype)


def left_shift(a, b):
    s = np.broadcast_to(a._other(b), np.broadcast(a.v, b.v).shape)
    bits = _RANGE[a.dtype][0]
    out = (np.broadcast_to(a.v, s.shape) % (1 << bits)) << np.clip(s, 0, bits - 1)
    return Expr(np.where(s >= bits, 0, out), a.dtype)


def maximum(a, b):
    return Expr(np.maximum(a.v, a._other(b)), a.dtype)


def _reduce(fn):
    def f(data, axis=None, keepdims=False):
        return Expr(fn(data.v, axis=axis, keepdims=keepdims), data.dtype)
    return f


max = _reduce(np.max)          # noqa: A001  (relay.max)
sum = _reduce(np.sum)          # noqa: A001  (relay.sum)


def mean(data, axis=None, keepdims=False):
    s = sum(data, axis=axis, keepdims=keepdims)
    n = data.v.shape[axis] if axis is not None else data.v.size
    return s / const(n, data.dtype)


def install():
    """Register the stand-in as ``tvm`` / ``tvm.relay`` (+ the one deep import layers.py makes) in sys.modules."""
    this = sys.modules[__name__]
    tvm = types.ModuleType("tvm")
    tvm.relay = this
    op = types.ModuleType("tvm.relay.op")
    tensor = types.ModuleType("tvm.relay.op.tensor")
    tensor.exp = None                                           # imported by layers.py:7, never used
    op.tensor = tensor
    sys.modules.update({"tvm": tvm, "tvm.relay": this, "tvm.relay.op": op, "tvm.relay.op.tensor": tensor})
