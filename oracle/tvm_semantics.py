"""CPU restatement (numpy) of the TVM-semantics integer row operators.  TEST INFRASTRUCTURE: imported by tests/ only.

PARITY: PINNED TO THE REFERENCE'S EXPRESSIONS, NOT TO A TVM RUN.  The reference states these operators as TVM Relay
expressions (/root/reference/TVM_benchmark/models/layers.py:329-404); TVM is not installed in this image and the reference
ships no vectors for them.  tests/golden/tvm_ops.npz holds the outputs of the reference's own layers.py, loaded unmodified
and executed on a numpy stand-in for the dozen relay primitives it uses (tests/golden/relay_shim.py, generator
make_tvm_golden.py); this file reproduces those vectors bit for bit (tests/test_tvm_oracle.py), so the STRUCTURE of every
operator is the reference's.  What remains unpinned is the meaning of the primitives themselves, which both this file and
the stand-in take from Relay's documented integer semantics:

  * int32 tensors, two's-complement wrap on add / sub / mul (Relay arithmetic lowers to LLVM integer ops);
  * ``a / b`` on integers is truncating division (``relay.divide`` -> ``tir.truncdiv``);
  * ``right_shift`` is arithmetic;  ``relay.const(float, 'int32')`` truncates toward zero;
  * ``relay.mean`` of an integer tensor is ``sum / count`` in the tensor's dtype;
  * ``cast(int32 -> int8)`` / ``cast(int32 -> uint32)`` keep the low bits.

Where Relay leaves the result to the backend this file (and csrc/ivit_tvm.cu) fixes a convention: x / 0 = 0, a left shift
by 32 or more gives 0.  Neither is reachable for inputs in the value ranges the deployed model produces.
"""
import numpy as np


def _w(a):
    """wrap an int64 array to int32"""
    return ((np.asarray(a, np.int64) + 2 ** 31) % 2 ** 32 - 2 ** 31).astype(np.int64)


def _div(a, b):
    """truncating int32 division, x / 0 := 0, INT_MIN / -1 wraps"""
    a = np.asarray(a, np.int64)
    b = np.broadcast_to(np.asarray(b, np.int64), a.shape)
    safe = np.where(b == 0, 1, b)
    q = np.abs(a) // np.abs(safe) * np.sign(a) * np.sign(safe)
    return _w(np.where(b == 0, 0, q))


def _shl(a, s):
    a = np.asarray(a, np.int64)
    s = np.broadcast_to(np.asarray(s, np.int64), a.shape)
    sc = np.clip(s, 0, 31)
    return np.where(s >= 32, 0, _w((a % 2 ** 32) << sc))


def x0_of(input_scale):
    """layers.py:357  relay.const(-1.0 / input_scale - 1, 'int32')"""
    return int(np.array(-1.0 / float(input_scale) - 1).astype("int32"))


def shift_exp(data, x0, n):
    """layers.py:353-369"""
    d = _w(np.asarray(data, np.int64))
    d = _w(_w(d + (d >> 1)) - (d >> 4))
    d = np.maximum(d, _w(n * x0))
    q = _div(d, x0)
    r = _w(d - _w(q * x0))
    return _shl(_w((r >> 1) - x0), _w(n - q))


def quantized_softmax(data, input_scale, n=16):
    """layers.py:372-386 -> int8"""
    x0 = x0_of(input_scale)
    d = _w(np.asarray(data, np.int64))
    d = _w(d - d.max(axis=-1, keepdims=True))
    e = shift_exp(d, x0, n)
    s = _w(e.sum(axis=-1, keepdims=True))
    out = _w(_div(np.full_like(s, 2 ** 31 - 1), s) * e) >> 24
    return ((out + 128) % 256 - 128).astype(np.int8)


def quantized_gelu(pre_data, input_scale, n=23):
    """layers.py:389-404 -> int32"""
    x0 = x0_of(float(input_scale) * 1.702)
    pre = _w(np.asarray(pre_data, np.int64))
    mx = pre.max(axis=-1, keepdims=True)
    e = shift_exp(_w(pre - mx), x0, n)
    e_max = shift_exp(_w(-mx), x0, n)
    s = _w(e + e_max)
    sig = _w(_div(np.full_like(s, 2 ** 31 - 1), s) * e) >> 24
    return _w(pre * sig).astype(np.int32)


def quantized_layernorm(data, bias_int):
    """layers.py:329-350 -> int32"""
    x = _w(np.asarray(data, np.int64))
    C = x.shape[-1]
    mean = _div(_w(x.sum(axis=-1, keepdims=True)), C)
    d = _w(x - mean)
    var = (_w(d * d) % 2 ** 32).sum(axis=-1, keepdims=True) % 2 ** 32          # uint32
    std = np.full_like(var, 2 ** 16)
    for _ in range(10):
        std = ((std + var // std) % 2 ** 32) // 2
    std = _w(std)
    out = _div(_w(_div(np.full_like(std, 2 ** 31 - 1), std) * d), 2)
    return _w(out + np.asarray(bias_int, np.int64)).astype(np.int32)
