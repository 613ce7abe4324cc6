"""CPU checks of the TVM-semantics oracle (oracle/tvm_semantics.py): hand-computed cases, closeness to the functions the
operators stand for, and bit-equality with vectors produced by the reference's own TVM_benchmark/models/layers.py executed
on a numpy stand-in for the relay primitives (TVM itself is absent from this image)."""
import numpy as np

from oracle import tvm_semantics as T


def test_x0_truncates_toward_zero():
    assert T.x0_of(0.05) == -21          # -1/0.05 - 1 = -21.0
    assert T.x0_of(0.03) == -34          # -34.33 -> -34
    assert T.x0_of(0.3) == -4            # -4.33 -> -4


def test_shift_exp_by_hand():
    # x0 = -21, n = 16: d = -40 -> d + (d>>1) - (d>>4) = -40 - 20 + 3 = -57; q = trunc(-57 / -21) = 2; r = -57 + 42 = -15
    # exp = ((-15 >> 1) + 21) << (16 - 2) = (−8 + 21) << 14
    assert int(T.shift_exp(np.array([-40]), -21, 16)[0]) == 13 << 14
    assert int(T.shift_exp(np.array([0]), -21, 16)[0]) == 21 << 16
    # clamp at n * x0: everything below gives q = n, r = 0
    assert int(T.shift_exp(np.array([-100000]), -21, 16)[0]) == 21


def test_softmax_close_to_float_softmax():
    rng = np.random.default_rng(1)
    x = rng.integers(-128, 128, (16, 197))
    p = T.quantized_softmax(x, 0.05)
    assert p.dtype == np.int8 and p.min() >= 0
    f = np.exp((x - x.max(-1, keepdims=True)) * 0.05)
    f /= f.sum(-1, keepdims=True)
    assert np.abs(p / 128.0 - f).max() < 0.03


def test_gelu_close_to_float_gelu():
    x = np.arange(-128, 128)[None, :]
    g = T.quantized_gelu(x, 0.03)
    xf = x * 0.03
    ref = xf / (1 + np.exp(-1.702 * xf))
    assert np.abs(g * (0.03 / 128) - ref).max() < 0.25      # (factor / sum) is a coarse integer at the row maximum


def test_layernorm_close_to_float_layernorm_and_truncating_mean():
    rng = np.random.default_rng(2)
    h = rng.integers(-3000, 3000, (8, 768))
    b = rng.integers(-2 ** 20, 2 ** 20, 768)
    ln = T.quantized_layernorm(h, b)
    hf = (h - h.mean(-1, keepdims=True)) / h.std(-1, keepdims=True)
    assert np.abs((ln.astype(np.int64) - b) * (np.sqrt(768) / 2 ** 30) - hf).max() < 2e-3
    # mean truncates toward zero: row sum -5 over 4 channels -> mean -1 (floor would give -2)
    row = np.array([[-2, -1, -1, -1]])
    out = T.quantized_layernorm(row, np.zeros(4, np.int64))
    d = row + 1
    var = int((d * d).sum())
    std = 2 ** 16
    for _ in range(10):
        std = (std + var // std) // 2
    f = (2 ** 31 - 1) // std
    want = np.array([int(np.trunc(f * v / 2)) for v in d[0]])
    assert np.array_equal(out[0], want)


# ---- the reference's OWN Relay expressions, executed on a numpy stand-in for the relay primitives ----------------------
import os  # noqa: E402

import pytest  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tvm_ops.npz")


def test_oracle_equals_the_reference_expressions():
    """tests/golden/tvm_ops.npz was produced by loading TVM_benchmark/models/layers.py unmodified on
    tests/golden/relay_shim.py (make_tvm_golden.py): the structure of every operator is the reference's."""
    g = np.load(GOLD)
    for i in range(3):
        assert np.array_equal(T.quantized_softmax(g["sm%d_x" % i], float(g["sm%d_s" % i])), g["sm%d_y" % i]), i
    for i in range(4):
        assert np.array_equal(T.quantized_gelu(g["ge%d_x" % i], float(g["ge%d_s" % i])), g["ge%d_y" % i]), i
    for i in range(4):
        assert np.array_equal(T.quantized_layernorm(g["ln%d_x" % i], g["ln%d_b" % i]), g["ln%d_y" % i]), i
    for i in range(3):
        y = T.shift_exp(g["se%d_d" % i], T.x0_of(float(g["se%d_s" % i])), int(g["se%d_n" % i]))
        assert np.array_equal(y, g["se%d_y" % i]), i
    assert np.abs(g["ge2_y"]).max() > 0 and g["ln2_y"].dtype == np.int32


@pytest.mark.skipif(not os.path.exists("/root/reference/TVM_benchmark/models/layers.py"), reason="reference checkout absent")
def test_golden_vectors_regenerate_from_the_reference_source():
    import importlib.util
    import sys
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, here)
    try:
        spec = importlib.util.spec_from_file_location("make_tvm_golden", os.path.join(here, "make_tvm_golden.py"))
        mk = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mk)
        L = mk.load_layers()
        import relay_shim as R
        g = np.load(GOLD)
        y = L.quantized_softmax(R.Expr(g["sm0_x"].astype(np.int64), "int8"), float(g["sm0_s"]))
        assert np.array_equal(y.v.astype(np.int8), g["sm0_y"])
        y = L.quantized_layernorm(R.Expr(g["ln1_x"].astype(np.int64), "int32"), R.Expr(g["ln1_b"].astype(np.int64), "int32"))
        assert np.array_equal(y.v.astype(np.int32), g["ln1_y"])
    finally:
        sys.path.remove(here)
        for m in ("tvm", "tvm.relay", "tvm.relay.op", "tvm.relay.op.tensor"):
            sys.modules.pop(m, None)
