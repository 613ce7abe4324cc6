"""Golden vectors of the TVM-semantics row operators, produced by executing the reference's OWN Relay expressions
(/root/reference/TVM_benchmark/models/layers.py:329-404, loaded unmodified) on the numpy stand-in of relay_shim.py:

    python tests/golden/make_tvm_golden.py        ->  tests/golden/tvm_ops.npz

Run in the build container (needs /root/reference); the .npz travels with the repository."""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import relay_shim as R  # noqa: E402

REF = os.environ.get("IVIT_REFERENCE", "/root/reference")


def load_layers():
    R.install()
    spec = importlib.util.spec_from_file_location("ref_tvm_layers", os.path.join(REF, "TVM_benchmark", "models", "layers.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    L = load_layers()
    rng = np.random.default_rng(2024)
    out = {}
    # softmax: int8 scores, three scales (x0 = -21, -4, -257), flat / one-hot rows included
    for i, (shape, s) in enumerate([((2, 3, 17, 50), 0.05), ((1, 2, 9, 197), 0.3), ((2, 1, 5, 64), 0.0039)]):
        x = rng.integers(-128, 128, shape).astype(np.int8)
        x[0, 0, 0, :] = 127
        x[0, 0, 1, :] = -128
        x[0, 0, 1, shape[-1] // 2] = 127
        y = L.quantized_softmax(R.Expr(x.astype(np.int64), "int8"), s)
        assert y.dtype == "int8"
        out["sm%d_x" % i], out["sm%d_s" % i], out["sm%d_y" % i] = x, np.float64(s), y.v.astype(np.int8)
    # gelu: int8 pre-activations, incl. an all-negative row and a scale whose exponentials wrap int32 (x0 = -148, n = 23)
    for i, (shape, s) in enumerate([((2, 7, 96), 0.03), ((1, 5, 200), 0.0734), ((2, 4, 64), 0.004), ((1, 3, 33), 0.12)]):
        x = rng.integers(-128, 128, shape).astype(np.int8)
        x[0, 1, :] = -np.abs(x[0, 1, :].astype(np.int16)).clip(1, 128).astype(np.int8)
        y = L.quantized_gelu(R.Expr(x.astype(np.int64), "int8"), s)
        assert y.dtype == "int32"
        out["ge%d_x" % i], out["ge%d_s" % i], out["ge%d_y" % i] = x, np.float64(s), y.v.astype(np.int32)
    # layernorm: [B, N, C] (the reference reduces over axis 2), magnitudes up to the uint32-variance wrap
    for i, (shape, mag) in enumerate([((2, 5, 768), 3000), ((1, 4, 192), 30000), ((1, 3, 1024), 200000), ((1, 2, 8), 0)]):
        x = rng.integers(-mag, mag + 1, shape).astype(np.int32)
        b = rng.integers(-2 ** 24, 2 ** 24, shape[-1]).astype(np.int32)
        y = L.quantized_layernorm(R.Expr(x.astype(np.int64), "int32"), R.Expr(b.astype(np.int64), "int32"))
        assert y.dtype == "int32"
        out["ln%d_x" % i], out["ln%d_b" % i], out["ln%d_y" % i] = x, b, y.v.astype(np.int32)
    # shift_exp alone (both n the reference uses)
    d = np.concatenate([np.arange(-4000, 1), rng.integers(-2 ** 20, 1, 200)]).astype(np.int32)
    for i, (s, n) in enumerate([(0.05, 16), (0.03 * 1.702, 23), (0.0039, 16)]):
        y = L.shift_exp(R.Expr(d.astype(np.int64), "int32"), s, n)
        out["se%d_d" % i], out["se%d_s" % i], out["se%d_n" % i], out["se%d_y" % i] = d, np.float64(s), np.int64(n), y.v.astype(np.int32)
    path = os.path.join(HERE, "tvm_ops.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%d arrays, %d bytes" % (len(out), os.path.getsize(path)))


if __name__ == "__main__":
    main()
