"""TVM-semantics compatibility mode: the sm_100a kernels behind ivit_b200.tvm_mode against
oracle/tvm_semantics.py and against vectors of the reference's own expressions (tests/golden/tvm_ops.npz), bit for bit."""
import numpy as np
import pytest
import torch

from oracle import tvm_semantics as T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    assert torch.cuda.is_available(), "needs a CUDA device: the product path has no CPU fallback"
    import ivit_b200.tvm_mode as M
    return M


@pytest.mark.parametrize("rows,cols,s", [(12, 197, 0.05), (5, 49, 0.3), (7, 1000, 0.0039), (3, 8, 0.9), (300, 197, 0.02)])
def test_softmax(M, rows, cols, s):
    rng = np.random.default_rng(rows * cols)
    x = rng.integers(-128, 128, (rows, cols)).astype(np.int32)
    x[0, :] = 127                                                     # flat row
    if rows > 2:
        x[1, :] = -128
        x[1, cols // 2] = 127                                         # one-hot row
    got = M.quantized_softmax(torch.from_numpy(x).cuda(), s).cpu().numpy()
    assert got.dtype == np.int8
    assert np.array_equal(got, T.quantized_softmax(x, s))


def test_softmax_int8_input_and_batched_shape(M):
    rng = np.random.default_rng(3)
    x = rng.integers(-128, 128, (2, 3, 17, 64)).astype(np.int8)
    got = M.quantized_softmax(torch.from_numpy(x).cuda(), 0.07).cpu().numpy()
    assert got.shape == x.shape and np.array_equal(got, T.quantized_softmax(x, 0.07))


@pytest.mark.parametrize("rows,cols,s,lo,hi", [(9, 3072, 0.03, -128, 128), (4, 768, 0.0734, -128, 128), (6, 100, 0.12, -128, 0),
                                               (5, 64, 0.004, -128, 128), (3, 33, 0.3, -20000, 20000)])
def test_gelu(M, rows, cols, s, lo, hi):
    rng = np.random.default_rng(cols)
    x = rng.integers(lo, hi, (rows, cols)).astype(np.int32)
    got = M.quantized_gelu(torch.from_numpy(x).cuda(), s).cpu().numpy()
    assert got.dtype == np.int32
    assert np.array_equal(got, T.quantized_gelu(x, s))                # includes the wrapping regime (|x0| large, n = 23)


@pytest.mark.parametrize("rows,C,mag", [(10, 768, 3000), (4, 192, 30000), (3, 1024, 200000), (5, 96, 5), (2, 8, 0)])
def test_layernorm(M, rows, C, mag):
    rng = np.random.default_rng(C)
    x = rng.integers(-mag, mag + 1, (rows, C)).astype(np.int32)
    b = rng.integers(-2 ** 24, 2 ** 24, C).astype(np.int32)
    got = M.quantized_layernorm(torch.from_numpy(x).cuda(), torch.from_numpy(b).cuda()).cpu().numpy()
    assert np.array_equal(got, T.quantized_layernorm(x, b))           # mag = 200000: the uint32 variance wraps


def test_refuses_cpu_tensors(M):
    with pytest.raises(RuntimeError):
        M.quantized_softmax(torch.zeros(2, 8, dtype=torch.int32), 0.1)


def test_kernels_equal_the_reference_expressions(M):
    """The CUDA kernels against tests/golden/tvm_ops.npz: vectors produced by the reference's own layers.py (unmodified)
    executed on the numpy stand-in for the relay primitives (tests/golden/make_tvm_golden.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tvm_ops.npz"))
    for i in range(3):
        got = M.quantized_softmax(torch.from_numpy(g["sm%d_x" % i]).cuda(), float(g["sm%d_s" % i])).cpu().numpy()
        assert np.array_equal(got, g["sm%d_y" % i]), "softmax %d" % i
    for i in range(4):
        got = M.quantized_gelu(torch.from_numpy(g["ge%d_x" % i]).cuda(), float(g["ge%d_s" % i])).cpu().numpy()
        assert np.array_equal(got, g["ge%d_y" % i]), "gelu %d" % i
    for i in range(4):
        got = M.quantized_layernorm(torch.from_numpy(g["ln%d_x" % i]).cuda(), torch.from_numpy(g["ln%d_b" % i]).cuda()).cpu().numpy()
        assert np.array_equal(got, g["ln%d_y" % i]), "layernorm %d" % i
