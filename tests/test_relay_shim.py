"""The numpy stand-in for the relay primitives (tests/golden/relay_shim.py) is what gives the TVM-semantics golden vectors
their meaning, so its own semantics are pinned here: two's-complement wrap, truncating signed / flooring unsigned division,
arithmetic right shift, wrapping casts, const conversion, reductions."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import relay_shim as R  # noqa: E402


def E(v, dt="int32"):
    return R.Expr(np.asarray(v, np.int64), dt)


def test_int32_arithmetic_wraps():
    assert (E([2 ** 31 - 1]) + E([1])).v.tolist() == [-2 ** 31]
    assert (E([-2 ** 31]) - E([1])).v.tolist() == [2 ** 31 - 1]
    assert (E([65536]) * E([65536])).v.tolist() == [0]
    assert (E([46341]) * E([46341])).v.tolist() == [46341 * 46341 - 2 ** 32]
    assert (-E([-2 ** 31])).v.tolist() == [-2 ** 31]


def test_division_truncates_for_signed_and_floors_for_unsigned():
    assert (E([-7, 7, -7, 7]) / E([2, 2, -2, -2])).v.tolist() == [-3, 3, 3, -3]
    assert (E([5]) / E([0])).v.tolist() == [0]                       # convention shared with the oracle
    u = R.Expr(np.array([2 ** 32 - 1]), "uint32")
    assert (u / R.const(2, "uint32")).v.tolist() == [2 ** 31 - 1]


def test_shifts():
    assert R.right_shift(E([-5, 5]), R.const(1, "int32")).v.tolist() == [-3, 2]      # arithmetic
    assert R.left_shift(E([3]), E([30])).v.tolist() == [-2 ** 30]                      # wraps
    assert R.left_shift(E([3]), E([32])).v.tolist() == [0]                             # convention


def test_const_and_cast():
    assert R.const(-21.9, "int32").v.tolist() == -21                 # numpy conversion truncates toward zero
    assert R.cast(E([200, -129, 127]), "int8").v.tolist() == [-56, 127, 127]
    assert R.cast(E([-1]), "uint32").v.tolist() == [2 ** 32 - 1]


def test_reductions():
    x = E([[1, -2, 3], [-7, -8, -9]])
    assert R.max(x, axis=-1, keepdims=True).v.tolist() == [[3], [-7]]
    assert R.sum(x, axis=1, keepdims=True).v.tolist() == [[2], [-24]]
    assert R.mean(x, axis=1, keepdims=True).v.tolist() == [[0], [-8]]          # sum / count, truncating
    assert R.maximum(E([1, -5]), E([0, 0])).v.tolist() == [1, 0]
