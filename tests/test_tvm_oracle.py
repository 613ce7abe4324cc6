"""CPU checks of the TVM-semantics oracle (oracle/tvm_semantics.py).  TVM is absent from this image and the reference
holds no vectors for TVM_benchmark/models/layers.py:329-404, so the restatement is UNPINNED; what can be checked here is
that it is the stated arithmetic (hand-computed cases) and that it approximates the functions it stands for."""
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
