"""CPU oracle for the I-ViT integer-only operators -- TEST INFRASTRUCTURE ONLY.

ctypes bindings over ``oracle/ivit_oracle.c`` (exact int64/int128 restatement of the
reference's ``models/quantization_utils`` forward semantics, each C function cites the
reference file:line it follows).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this package,
and only as the checker.  The product path (``i-vit_b200/``) never imports it.

Parity status: PINNED -- ``tests/test_oracle_golden.py`` checks every function here
against golden vectors produced by executing the reference's own unmodified modules
(``tests/golden/make_golden.py``, exact-carrier hooks of SURVEY.md section 8c).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libivit_oracle.so")


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, seconds)."""
    src = os.path.join(_HERE, "ivit_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        i64, f32, vp = C.c_int64, C.c_float, C.c_void_p
        L.ivo_sym_scale.restype = f32
        L.ivo_sym_scale.argtypes = [C.c_int, f32, f32]
        L.ivo_quantize_f32.argtypes = [vp, i64, vp, i64, i64, C.c_int, vp]
        L.ivo_dyadic.argtypes = [vp, i64, f32, vp, vp]
        L.ivo_requant.argtypes = [vp, i64, i64, vp, vp, i64, vp, i64, vp, vp, i64, C.c_int, vp, vp]
        L.ivo_gemm_nt.argtypes = [vp, vp, vp, i64, i64, i64, vp]
        L.ivo_gemm_nt_i8.argtypes = [vp, vp, vp, i64, i64, i64, vp]
        L.ivo_x0.restype = i64
        L.ivo_x0.argtypes = [f32]
        L.ivo_gelu_sig_scale.restype = f32
        L.ivo_gelu_sig_scale.argtypes = [f32]
        L.ivo_shiftmax.argtypes = [vp, i64, i64, i64, C.c_int, C.c_int, vp]
        L.ivo_shiftgelu.argtypes = [vp, i64, i64, i64, C.c_int, C.c_int, vp]
        L.ivo_layernorm.argtypes = [vp, i64, i64, vp, vp]
        L.ivo_avgpool_rne.argtypes = [vp, i64, i64, i64, vp]
        L.ivo_set_threads.argtypes = [C.c_int]
        L.ivo_get_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i64(a):
    return np.ascontiguousarray(np.asarray(a), dtype=np.int64)


def _f32(a):
    return np.ascontiguousarray(np.asarray(a), dtype=np.float32)


def set_threads(n: int) -> int:
    """Host threads used by the int8 GEMM (the CPU baseline uses every core)."""
    lib().ivo_set_threads(int(n))
    return lib().ivo_get_threads()


# --------------------------------------------------------------------------- primitives
def sym_scale(bits: int, min_val: float, max_val: float) -> np.float32:
    """quant_utils.py:51-69"""
    return np.float32(lib().ivo_sym_scale(bits, np.float32(min_val), np.float32(max_val)))


def quantize_f32(x, scale, bits: int, per_row: bool = False) -> np.ndarray:
    """quant_utils.py:48,90-92. ``scale`` scalar, or one per leading row when per_row."""
    x = _f32(x)
    s = _f32(scale).reshape(-1)
    out = np.empty(x.shape, np.int64)
    inner = int(np.prod(x.shape[1:])) if per_row else 1
    lib().ivo_quantize_f32(_p(x), x.size, _p(s), s.size, max(inner, 1), bits, _p(out))
    return out


def dyadic(s_in, s_out):
    """batch_frexp of fp64(s_in)/fp64(fp32(s_out)): quant_utils.py:150-175, 221-228.
    Returns (m, e) int64 arrays, un-normalised (|m| in [2^30, 2^31])."""
    s = _f32(s_in).reshape(-1)
    m = np.empty(s.size, np.int64)
    e = np.empty(s.size, np.int64)
    lib().ivo_dyadic(_p(s), s.size, np.float32(s_out), _p(m), _p(e))
    return m, e


def requant(z, m, e, bits: int, w=None, m1=None, e1=None, return_diff: bool = False):
    """fixedpoint_mul.forward, quant_utils.py:192-253. z [..., cols]; m/e len 1 or cols;
    optional residual w (same shape as z, or one row broadcast) with m1/e1."""
    z = _i64(z)
    cols = z.shape[-1]
    rows = z.size // cols
    m, e = _i64(m).reshape(-1), _i64(e).reshape(-1)
    out = np.empty(z.shape, np.int64)
    diff = C.c_int64(0)
    if w is not None:
        w = _i64(w)
        wrows = w.size // cols
        assert rows % wrows == 0
        m1, e1 = _i64(m1).reshape(-1), _i64(e1).reshape(-1)
        lib().ivo_requant(_p(z), rows, cols, _p(m), _p(e), m.size, _p(w), wrows,
                          _p(m1), _p(e1), m1.size, bits, _p(out), C.byref(diff))
    else:
        lib().ivo_requant(_p(z), rows, cols, _p(m), _p(e), m.size, None, 0,
                          None, None, 0, bits, _p(out), C.byref(diff))
    return (out, diff.value) if return_diff else out


def gemm_nt(a, w, bias=None):
    """acc = a @ w.T + bias, exact.  quant_modules.py:93-97, 224-228, 325-330."""
    a = np.asarray(a)
    w = np.asarray(w)
    M, K = a.shape
    N = w.shape[0]
    if a.dtype == np.int8 and w.dtype == np.int8:
        a, w = np.ascontiguousarray(a), np.ascontiguousarray(w)
        b = None if bias is None else np.ascontiguousarray(bias, dtype=np.int32)
        out = np.empty((M, N), np.int32)
        lib().ivo_gemm_nt_i8(_p(a), _p(w), _p(b), M, N, K, _p(out))
        return out.astype(np.int64)
    a32 = np.ascontiguousarray(a, dtype=np.int32)
    w32 = np.ascontiguousarray(w, dtype=np.int32)
    b = None if bias is None else _i64(bias)
    out = np.empty((M, N), np.int64)
    lib().ivo_gemm_nt(_p(a32), _p(w32), _p(b), M, N, K, _p(out))
    return out


def x0_of(scale) -> int:
    """floor(-1/s) in fp32: quant_modules.py:414,473"""
    return int(lib().ivo_x0(np.float32(scale)))


def gelu_sig_scale(scale) -> np.float32:
    """fp32(s*1.702): quant_modules.py:427"""
    return np.float32(lib().ivo_gelu_sig_scale(np.float32(scale)))


def shiftmax(q, x0: int, out_bits: int, n: int = 15):
    """IntSoftmax.forward, quant_modules.py:483-497 (rows over the last dim)."""
    q = _i64(q)
    cols = q.shape[-1]
    out = np.empty(q.shape, np.int64)
    lib().ivo_shiftmax(_p(q), q.size // cols, cols, int(x0), n, out_bits, _p(out))
    return out


def shiftgelu(q, x0: int, out_bits: int = 8, n: int = 23):
    """IntGELU.forward, quant_modules.py:425-445."""
    q = _i64(q)
    cols = q.shape[-1]
    out = np.empty(q.shape, np.int64)
    lib().ivo_shiftgelu(_p(q), q.size // cols, cols, int(x0), n, out_bits, _p(out))
    return out


def layernorm(q, bias_int=None):
    """IntLayerNorm.forward, quant_modules.py:353-386 (integer part)."""
    q = _i64(q)
    cols = q.shape[-1]
    out = np.empty(q.shape, np.int64)
    b = None if bias_int is None else _i64(bias_int)
    lib().ivo_layernorm(_p(q), q.size // cols, cols, _p(b), _p(out))
    return out


def avgpool_rne(q):
    """[B, L, C] -> [B, C], RNE(sum/L): swin_quant.py:554-555 integer reading."""
    q = _i64(q)
    B, L, Cc = q.shape
    out = np.empty((B, Cc), np.int64)
    lib().ivo_avgpool_rne(_p(q), B, L, Cc, _p(out))
    return out
