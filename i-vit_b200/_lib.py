"""ctypes binding of ``libivit_b200.so`` (the C ABI declared in ``include/ivit_b200.h``).

This is the only bridge between the Python host code and the sm_100a kernels.  There is no
CPU path: creating a context without a Blackwell GPU raises ``IvitError`` and every operator
built on top of it fails loudly.  torch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# IVIT_B200_SO: a differently built copy of the same library (kernel A/B experiments under tools/)
SO_PATH = os.environ.get("IVIT_B200_SO") or os.path.join(_HERE, "csrc", "libivit_b200.so")

I8, I16, I32, F32, U8, F64 = 0, 1, 2, 3, 4, 5
EPI_RAW_I32, EPI_REQUANT, EPI_CARRIER = 0, 1, 2

TORCH2IVIT = {torch.int8: I8, torch.int16: I16, torch.int32: I32, torch.float32: F32, torch.uint8: U8,
              torch.float64: F64}
IVIT2TORCH = {v: k for k, v in TORCH2IVIT.items()}


class IvitError(RuntimeError):
    pass


class Dyadic(C.Structure):
    _fields_ = [("m", C.c_int32), ("e", C.c_int32)]


class GemmEpilogue(C.Structure):
    _fields_ = [("mode", C.c_int), ("bias", C.c_void_p), ("me", C.c_void_p), ("bits", C.c_int),
                ("residual", C.c_void_p), ("res_dtype", C.c_int), ("res_ld", C.c_int64),
                ("res_me", Dyadic), ("two_stage", C.c_int), ("me2", Dyadic), ("scale", C.c_void_p),
                ("out_dtype", C.c_int), ("out_ld", C.c_int64), ("acc_bits", C.c_int)]


class AttnParams(C.Structure):
    _fields_ = [("n_seq", C.c_int), ("n_tok", C.c_int), ("n_heads", C.c_int), ("head_dim", C.c_int),
                ("me_s", Dyadic), ("x0", C.c_int32), ("n", C.c_int), ("p_bits", C.c_int),
                ("me_o", Dyadic), ("relbias", C.c_void_p), ("me_s2", Dyadic), ("me_b", Dyadic),
                ("mask", C.c_void_p), ("n_win", C.c_int)]


class WinAttnParams(C.Structure):
    _fields_ = [("n_win", C.c_int), ("n_heads", C.c_int), ("n_tok", C.c_int), ("head_dim", C.c_int),
                ("me_s", Dyadic), ("me_s2", Dyadic), ("x0", C.c_int32), ("n", C.c_int), ("p_bits", C.c_int),
                ("me_o", Dyadic), ("bias_rq", C.c_void_p), ("mask_bits", C.c_void_p), ("n_win_img", C.c_int),
                ("mask_add", C.c_int32)]


_vp, _i64, _int = C.c_void_p, C.c_int64, C.c_int
# name -> argtypes (ctx and stream included); all return int except the first three
SIGNATURES = {
    "ivit_dyadic": [_vp, _vp, _int, _vp, _vp, _vp],
    "ivit_quantize_f32": [_vp, _vp, _i64, _vp, _i64, _i64, _int, _int, _vp, _vp],
    "ivit_carrier_to_int": [_vp, _vp, _int, _i64, _int, _vp, _int, _int, _vp, _vp],
    "ivit_int_to_carrier": [_vp, _vp, _int, _i64, _int, _vp, _int, _int, _vp, _vp],
    "ivit_requant": [_vp, _vp, _int, _i64, _int, _vp, _int, _vp, _int, _i64, _vp, _int, _int, _int, _vp, _vp],
    "ivit_gemm_i8": [_vp, _vp, _i64, _vp, _i64, _i64, _i64, C.POINTER(GemmEpilogue), _vp, _vp],
    "ivit_bmm_i32": [_vp, _vp, _int, _i64, _i64, _vp, _i64, _i64, _int, _i64, _int, _int, _int, _vp, _i64, _i64, _vp],
    "ivit_layernorm": [_vp, _vp, _int, _i64, _int, _vp, _vp, _int, _int, _vp, _vp],
    "ivit_shiftmax": [_vp, _vp, _int, _i64, _int, C.c_int32, _int, _int, _int, _vp, _vp],
    "ivit_shiftgelu": [_vp, _vp, _int, _i64, _int, C.c_int32, _int, _vp, _int, _int, _vp, _vp],
    "ivit_attention_i8": [_vp, _vp, C.POINTER(AttnParams), _vp, _vp],
    "ivit_patchify_i8": [_vp, _vp, _int, _int, _int, _int, _int, _vp, _vp],
    "ivit_embed_tokens": [_vp, _vp, _vp, _vp, _int, _int, _int, Dyadic, Dyadic, _int, _vp, _vp],
    "ivit_shiftgelu_build_lut": [_vp, C.c_int32, _int, _vp, _int, _vp, _vp],
    "ivit_shiftgelu_lut": [_vp, _vp, _i64, _int, _vp, _vp, _vp],
    "ivit_layernorm_i16_i8": [_vp, _vp, _i64, _int, _vp, _vp, _vp, _vp],
    "ivit_quantize_patchify": [_vp, _vp, _vp, _int, _int, _int, _int, _int, _vp, _vp],
    "ivit_quantize_patchify_u8": [_vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _int, _vp, _vp],
    "ivit_embed_tokens_fast": [_vp, _vp, _vp, _vp, _int, _int, _int, Dyadic, Dyadic, _vp, _vp],
    "ivit_window_attention_i8": [_vp, _vp, C.POINTER(WinAttnParams), _vp, _vp],
    "ivit_layernorm_gather_i16_i8": [_vp, _vp, _i64, _int, _int, _vp, _int, _int, _vp, _vp, _vp, _vp, _vp],
    "ivit_avgpool_requant_i8": [_vp, _vp, _int, _int, _int, Dyadic, _vp, _vp],
    "ivit_widen_i8_i16": [_vp, _vp, _i64, _vp, _vp],
    "ivit_layernorm_i8_i16x2": [_vp, _vp, _i64, _int, _vp, _vp, Dyadic, _vp, _vp],
    "ivit_tvm_softmax": [_vp, _vp, _i64, _int, C.c_int32, _int, _vp, _vp],
    "ivit_tvm_gelu": [_vp, _vp, _i64, _int, C.c_int32, _int, _vp, _vp],
    "ivit_tvm_layernorm": [_vp, _vp, _i64, _int, _vp, _vp, _vp],
}
EXPORTS = ["ivit_version", "ivit_last_error", "ivit_create", "ivit_destroy", "ivit_num_sms"] + sorted(SIGNATURES)

_dll = None
_lock = threading.Lock()


def load_library() -> C.CDLL:
    """dlopen the in-tree shared library (no compute; works without a GPU)."""
    global _dll
    with _lock:
        if _dll is None:
            if not os.path.exists(SO_PATH):
                raise IvitError("%s is missing: build it with `python i-vit_b200/csrc/build.py` "
                                "(there is no fallback implementation)" % SO_PATH)
            dll = C.CDLL(SO_PATH)
            dll.ivit_version.restype = C.c_int
            dll.ivit_last_error.restype = C.c_char_p
            dll.ivit_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
            dll.ivit_destroy.argtypes = [C.c_void_p]
            dll.ivit_num_sms.argtypes = [C.c_void_p]
            for name, args in SIGNATURES.items():
                fn = getattr(dll, name)
                fn.argtypes = args
                fn.restype = C.c_int
            _dll = dll
    return _dll


def _check(rc: int, what: str):
    if rc != 0:
        raise IvitError("%s failed (%d): %s" % (what, rc, load_library().ivit_last_error().decode()))


class Context:
    """One ``ivit_ctx`` per CUDA device."""

    def __init__(self, device: int):
        dll = load_library()
        h = C.c_void_p()
        _check(dll.ivit_create(int(device), C.byref(h)), "ivit_create")
        self.handle = h
        self.device = int(device)
        self.num_sms = dll.ivit_num_sms(h)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                load_library().ivit_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


_contexts: dict[int, Context] = {}


def context(device=None) -> Context:
    if not torch.cuda.is_available():
        raise IvitError("ivit_b200 needs a CUDA (sm_100a) device: there is no CPU implementation")
    if device is None:
        device = torch.cuda.current_device()
    elif isinstance(device, torch.device):
        device = device.index if device.index is not None else torch.cuda.current_device()
    with _lock:
        ctx = _contexts.get(device)
    if ctx is None:
        ctx = Context(device)
        with _lock:
            _contexts[device] = ctx
    return ctx


def call(name: str, ctx: Context, *args):
    """Invoke ``name(ctx, *args, stream)`` on torch's current stream OF THE CONTEXT'S DEVICE, with that device current
    for the duration of the launch (one process may drive several GPUs; kernel launches go to the current device)."""
    fn = getattr(load_library(), name)
    if torch.cuda.current_device() == ctx.device:
        rc = fn(ctx.handle, *args, torch.cuda.current_stream(ctx.device).cuda_stream)
    else:
        with torch.cuda.device(ctx.device):
            rc = fn(ctx.handle, *args, torch.cuda.current_stream(ctx.device).cuda_stream)
    _check(rc, name)


def ptr(t):
    if t is None:
        return None
    assert t.is_cuda, "device tensor expected"
    return t.data_ptr()


def dy(m: int, e: int) -> Dyadic:
    return Dyadic(int(m), int(e))
