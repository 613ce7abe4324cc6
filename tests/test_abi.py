"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, and exports every
symbol include/ivit_b200.h declares (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ivit_b200.h")).read()
    return sorted(set(re.findall(r"IVIT_API\s+[\w\s\*]+?\b(ivit_\w+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ["ivit_create", "ivit_gemm_i8", "ivit_requant", "ivit_layernorm", "ivit_shiftmax",
              "ivit_shiftgelu", "ivit_attention_i8", "ivit_quantize_f32", "ivit_dyadic", "ivit_bmm_i32"]:
        assert s in syms


def test_library_loads_and_exports_all_symbols():
    import ivit_b200._lib as L
    dll = L.load_library()
    for s in declared_symbols():
        assert hasattr(dll, s), "libivit_b200.so does not export %s" % s
    assert sorted(L.EXPORTS) == declared_symbols(), "ctypes binding and header disagree"
    assert dll.ivit_version() >= 100


def test_struct_layouts_match_header():
    import ivit_b200._lib as L
    assert ctypes.sizeof(L.Dyadic) == 8
    # field order mirrors the C structs; sizes computed by the same (native) ABI rules
    assert [f[0] for f in L.GemmEpilogue._fields_] == ["mode", "bias", "me", "bits", "residual", "res_dtype",
                                                       "res_ld", "res_me", "two_stage", "me2", "scale",
                                                       "out_dtype", "out_ld", "acc_bits"]
    assert [f[0] for f in L.AttnParams._fields_][:4] == ["n_seq", "n_tok", "n_heads", "head_dim"]


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import ivit_b200._lib as L
    with pytest.raises(L.IvitError):
        L.context()
    h = ctypes.c_void_p()
    rc = L.load_library().ivit_create(0, ctypes.byref(h))
    assert rc != 0 and b"no CUDA device" in L.load_library().ivit_last_error()
