"""Tensor-level wrappers over the C ABI (``_lib.py``): allocate outputs, check shapes, launch on
torch's current stream.  Everything here runs on the GPU through libivit_b200.so; there is no
torch / CPU fallback for any operator."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import F32, I8, I16, I32, TORCH2IVIT, Dyadic, call, context, ptr

_BITS_DTYPE = {8: torch.int8, 16: torch.int16, 32: torch.int32}


def dyadic_host(s_in, s_out):
    """Host version of batch_frexp on fp64(s_in)/fp64(fp32(s_out)) (reference quant_utils.py:150-175,
    221-228) for STATIC tables: returns int32 arrays (m, e), m normalised to fit int32, e clamped
    to [-1, 63] (identical to the device kernel ``ivit_dyadic``)."""
    s = np.asarray(s_in, dtype=np.float32).reshape(-1).astype(np.float64)
    r = s / np.float64(np.float32(s_out))
    mant, ex = np.frexp(r)
    sc = mant * 2147483648.0
    m = np.where(sc >= 0, np.floor(sc + 0.5), -np.floor(-sc + 0.5)).astype(np.int64)   # half away from zero
    e = 31 - ex.astype(np.int64)
    big = np.abs(m) == 2 ** 31
    m = np.where(big, m // 2, m)
    e = np.where(big, e - 1, e)
    e = np.where(m == 0, 63, e)
    e = np.clip(e, -1, 63)
    return m.astype(np.int32), e.astype(np.int32)


def dyadic_table(m, e, device) -> torch.Tensor:
    """Pack (m, e) arrays into the device layout of ``ivit_dyadic_t[]`` (int32 [n, 2])."""
    t = np.stack([np.asarray(m, np.int32).reshape(-1), np.asarray(e, np.int32).reshape(-1)], axis=1)
    return torch.from_numpy(np.ascontiguousarray(t)).to(device)


def dyadic_device(s_in: torch.Tensor, s_out: torch.Tensor) -> torch.Tensor:
    """ivit_dyadic: device-side batch_frexp, no host sync.  Returns int32 [n, 2]."""
    s_in = s_in.reshape(-1).contiguous().float()
    s_out = s_out.reshape(-1)[:1].contiguous().float()
    out = torch.empty((s_in.numel(), 2), dtype=torch.int32, device=s_in.device)
    call("ivit_dyadic", context(s_in.device), ptr(s_in), s_in.numel(), ptr(s_out), ptr(out))
    return out


def quantize_f32(x: torch.Tensor, scale: torch.Tensor, bits: int, per_row: bool = False, out_dtype=None):
    x = x.contiguous().float()
    scale = scale.reshape(-1).contiguous().float()
    out_dtype = out_dtype or _BITS_DTYPE[8 if bits <= 8 else (16 if bits <= 16 else 32)]
    out = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    inner = int(np.prod(x.shape[1:])) if per_row else 1
    call("ivit_quantize_f32", context(x.device), ptr(x), x.numel(), ptr(scale), scale.numel(),
         max(inner, 1), bits, TORCH2IVIT[out_dtype], ptr(out))
    return out


# ---------------------------------------------------------------------------------------------------------------
# Integer shadows of carriers.  The operator API passes fp32/fp64 "integer x scale" carriers between modules
# (quant_modules.py: every forward returns x_int * scaling_factor and the next one divides it out again).  Every
# carrier this package creates from an integer tensor remembers that tensor: a consumer that is handed the carrier --
# or ANY VIEW of it (reshape / permute / transpose / slicing share the storage, so the same sizes, strides and offset
# address the same elements of the integer tensor) -- together with the same scale tensor gets the integers back
# without the divide-and-round pass (29 % of an operator-level DeiT forward).  A shadow is valid only while the carrier
# base tensor object is alive and unmodified (weak reference + version counter) and for the scale it was made with
# (the scale tensor is kept alive, so its address cannot be reused); copies (cat, contiguous() of a permuted view,
# arithmetic on the carrier) have no shadow and take the conversion kernel.  IVIT_SHADOW=0 disables the mechanism.
# ---------------------------------------------------------------------------------------------------------------
import os as _os
import weakref as _weakref

_SHADOW = {}
_SHADOW_ON = _os.environ.get("IVIT_SHADOW", "1") != "0"
_WIDER = {torch.int8: (torch.int8, torch.int16, torch.int32), torch.int16: (torch.int16, torch.int32),
          torch.int32: (torch.int32,)}


def _shadow_register(carrier: torch.Tensor, q: torch.Tensor, s: torch.Tensor):
    if not _SHADOW_ON or carrier._base is not None or q.shape != carrier.shape or not q.is_contiguous():
        return
    if q.storage_offset() != 0 or carrier.storage_offset() != 0:
        return                   # views are re-derived with the CARRIER's absolute storage offsets (as_strided below)
    key = id(carrier)
    _SHADOW[key] = (_weakref.ref(carrier, lambda _r, k=key: _SHADOW.pop(k, None)), carrier._version, q, s, s._version)


def _shadow_lookup(x: torch.Tensor, s: torch.Tensor):
    """The integer tensor behind carrier `x` (a view with x's sizes / strides) if x is a registered carrier, or a view
    of one, and `s` is the scale it was made with; else None."""
    if not _SHADOW_ON:
        return None
    base = x._base if x._base is not None else x
    e = _SHADOW.get(id(base))
    if e is None:
        return None
    ref, ver, q, s0, sver = e
    if ref() is not base or base._version != ver or s0._version != sver or x.dtype != base.dtype:
        return None
    if s.data_ptr() != s0.data_ptr() or s.numel() != s0.numel() or s.dtype != s0.dtype:
        return None
    if x is base:
        return q
    return q.as_strided(x.size(), x.stride(), x.storage_offset())


def carrier_to_int_any(x: torch.Tensor, s: torch.Tensor):
    """Integers of carrier x in whatever integer type they were produced in (shadow hit: no conversion pass), else int32."""
    qv = _shadow_lookup(x, s)
    if qv is not None:
        return qv.contiguous()
    return carrier_to_int(x, s, torch.int32)


def carrier_to_int(x: torch.Tensor, s: torch.Tensor, out_dtype=torch.int32):
    """z = RNE(x / s[c]); x is an fp32 (or fp64, see int_to_carrier) carrier."""
    qv = _shadow_lookup(x, s)
    if qv is not None and out_dtype in _WIDER.get(qv.dtype, ()):
        return qv.contiguous() if qv.dtype == out_dtype else qv.to(out_dtype)
    x = x.contiguous()
    if x.dtype not in (torch.float32, torch.float64):
        x = x.float()
    s = s.reshape(-1).contiguous().float()
    cols = x.shape[-1]
    out = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    call("ivit_carrier_to_int", context(x.device), ptr(x), TORCH2IVIT[x.dtype], x.numel() // cols, cols, ptr(s),
         s.numel(), TORCH2IVIT[out_dtype], ptr(out))
    return out


def int_to_carrier(q: torch.Tensor, s: torch.Tensor, out_dtype=torch.float32):
    """x = q * s[c].  out_dtype float64 keeps integers beyond 2^24 exact (IntLayerNorm outputs)."""
    q = q.contiguous()
    s_arg = s                                              # the caller's scale tensor: identity of the shadow
    s = s.reshape(-1).contiguous().float()
    cols = q.shape[-1]
    out = torch.empty(q.shape, dtype=out_dtype, device=q.device)
    call("ivit_int_to_carrier", context(q.device), ptr(q), TORCH2IVIT[q.dtype], q.numel() // cols, cols,
         ptr(s), s.numel(), TORCH2IVIT[out_dtype], ptr(out))
    _shadow_register(out, q, s_arg)
    return out


def requant(z: torch.Tensor, me: torch.Tensor, bits: int, w: torch.Tensor = None, me1: torch.Tensor = None,
            out_dtype=None):
    z = z.contiguous()
    cols = z.shape[-1]
    rows = z.numel() // cols
    out_dtype = out_dtype or _BITS_DTYPE[8 if bits <= 8 else (16 if bits <= 16 else 32)]
    out = torch.empty(z.shape, dtype=out_dtype, device=z.device)
    if w is not None:
        w = w.contiguous()
        w_rows = w.numel() // cols
        call("ivit_requant", context(z.device), ptr(z), TORCH2IVIT[z.dtype], rows, cols, ptr(me), me.shape[0],
             ptr(w), TORCH2IVIT[w.dtype], w_rows, ptr(me1), me1.shape[0], bits, TORCH2IVIT[out_dtype], ptr(out))
    else:
        call("ivit_requant", context(z.device), ptr(z), TORCH2IVIT[z.dtype], rows, cols, ptr(me), me.shape[0],
             None, 0, 0, None, 0, bits, TORCH2IVIT[out_dtype], ptr(out))
    return out


def gemm_i8(a: torch.Tensor, w: torch.Tensor, *, bias=None, mode="raw", me=None, bits=8, residual=None,
            res_me=(0, 63), two_stage=False, me2=(0, 63), scale=None, out=None, acc_bits=0):
    """acc = a @ w.T (+bias) on the tcgen05 int8 path, with the fused epilogue selected by ``mode``:
    'raw' -> int32, 'carrier' -> fp32 acc*scale[n], 'requant' -> int8/int16 (per-channel ``me``)."""
    assert a.dtype == torch.int8 and w.dtype == torch.int8 and a.dim() == 2 and w.dim() == 2
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and a.stride(1) == 1 and w.is_contiguous()
    epi = _lib.GemmEpilogue()
    epi.bias = ptr(bias)
    if mode == "raw":
        epi.mode, odt = _lib.EPI_RAW_I32, torch.int32
    elif mode == "carrier":
        epi.mode, odt = _lib.EPI_CARRIER, torch.float32
        epi.scale = ptr(scale)
    elif mode == "requant":
        epi.mode, odt = _lib.EPI_REQUANT, _BITS_DTYPE[bits]
        epi.me = ptr(me)
        epi.bits = bits
        if residual is not None:
            assert residual.dtype == torch.int16 and residual.stride(-1) == 1
            epi.residual = ptr(residual)
            epi.res_dtype = I16
            epi.res_ld = residual.stride(0)
            epi.res_me = Dyadic(int(res_me[0]), int(res_me[1]))
        epi.two_stage = 1 if two_stage else 0
        epi.me2 = Dyadic(int(me2[0]), int(me2[1]))
    else:
        raise ValueError(mode)
    if out is None:
        out = torch.empty((M, N), dtype=odt, device=a.device)
    assert out.dtype == odt and out.stride(-1) == 1
    epi.out_dtype = TORCH2IVIT[odt]
    epi.out_ld = out.stride(0)
    epi.acc_bits = int(acc_bits)
    call("ivit_gemm_i8", context(a.device), ptr(a), a.stride(0), ptr(w), M, N, K, C.byref(epi), ptr(out))
    return out


def bmm_i32(a: torch.Tensor, b: torch.Tensor, trans_b: bool):
    """Batched raw int32 matmul on strided [batch, M, K] views (last dim contiguous).
    trans_b: b is [batch, N, K]; else [batch, K, N]."""
    assert a.dim() == 3 and b.dim() == 3 and a.stride(2) == 1 and b.stride(2) == 1
    batch, M, K = a.shape
    N = b.shape[1] if trans_b else b.shape[2]
    c = torch.empty((batch, M, N), dtype=torch.int32, device=a.device)
    call("ivit_bmm_i32", context(a.device), ptr(a), TORCH2IVIT[a.dtype], a.stride(1), a.stride(0),
         ptr(b), b.stride(1), b.stride(0), 1 if trans_b else 0, batch, M, N, K, ptr(c), N, M * N)
    return c


def layernorm(x: torch.Tensor, bias_int: torch.Tensor, me: torch.Tensor = None, bits: int = 8, out_dtype=None):
    x = x.contiguous()
    Cc = x.shape[-1]
    out_dtype = out_dtype or (torch.int32 if me is None else _BITS_DTYPE[bits])
    out = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    call("ivit_layernorm", context(x.device), ptr(x), TORCH2IVIT[x.dtype], x.numel() // Cc, Cc, ptr(bias_int),
         ptr(me), bits, TORCH2IVIT[out_dtype], ptr(out))
    return out


def shiftmax(q: torch.Tensor, x0: int, out_bits: int, n: int = 15, out_dtype=None):
    q = q.contiguous()
    cols = q.shape[-1]
    out_dtype = out_dtype or (torch.int16 if out_bits == 16 else torch.int8)
    out = torch.empty(q.shape, dtype=out_dtype, device=q.device)
    call("ivit_shiftmax", context(q.device), ptr(q), TORCH2IVIT[q.dtype], q.numel() // cols, cols, int(x0), n,
         out_bits, TORCH2IVIT[out_dtype], ptr(out))
    return out


def shiftgelu(q: torch.Tensor, x0: int, me: torch.Tensor = None, bits: int = 8, n: int = 23, out_dtype=None):
    q = q.contiguous()
    cols = q.shape[-1]
    out_dtype = out_dtype or (torch.int16 if me is None else _BITS_DTYPE[bits])
    out = torch.empty(q.shape, dtype=out_dtype, device=q.device)
    call("ivit_shiftgelu", context(q.device), ptr(q), TORCH2IVIT[q.dtype], q.numel() // cols, cols, int(x0), n,
         ptr(me), bits, TORCH2IVIT[out_dtype], ptr(out))
    return out


def attention_i8(qkv: torch.Tensor, n_seq: int, n_tok: int, n_heads: int, head_dim: int, me_s, x0: int,
                 me_o, p_bits: int = 16, n: int = 15, relbias=None, me_s2=(0, 63), me_b=(0, 63), mask=None,
                 n_win: int = 0, out=None):
    assert qkv.dtype == torch.int8 and qkv.is_contiguous()
    assert qkv.shape == (n_seq * n_tok, 3 * n_heads * head_dim)
    if out is None:
        out = torch.empty((n_seq * n_tok, n_heads * head_dim), dtype=torch.int8, device=qkv.device)
    p = _lib.AttnParams()
    p.n_seq, p.n_tok, p.n_heads, p.head_dim = n_seq, n_tok, n_heads, head_dim
    p.me_s = Dyadic(int(me_s[0]), int(me_s[1]))
    p.x0, p.n, p.p_bits = int(x0), n, p_bits
    p.me_o = Dyadic(int(me_o[0]), int(me_o[1]))
    p.relbias = ptr(relbias)
    p.me_s2 = Dyadic(int(me_s2[0]), int(me_s2[1]))
    p.me_b = Dyadic(int(me_b[0]), int(me_b[1]))
    p.mask = ptr(mask)
    p.n_win = n_win
    call("ivit_attention_i8", context(qkv.device), ptr(qkv), C.byref(p), ptr(out))
    return out


def patchify_i8(x: torch.Tensor, patch: int):
    assert x.dtype == torch.int8 and x.dim() == 4 and x.is_contiguous()
    B, Cin, H, W = x.shape
    out = torch.empty((B * (H // patch) * (W // patch), Cin * patch * patch), dtype=torch.int8, device=x.device)
    call("ivit_patchify_i8", context(x.device), ptr(x), B, Cin, H, W, patch, ptr(out))
    return out


def embed_tokens(pe16: torch.Tensor, cls32: torch.Tensor, pos16: torch.Tensor, B: int, n_tok: int, C: int,
                 me, me_res, bits: int = 16, out=None):
    """cls-token cat + position-embedding residual QuantAct (vit_quant.py:259-265)."""
    assert pe16.dtype == torch.int16 and cls32.dtype == torch.int32 and pos16.dtype == torch.int16
    if out is None:
        out = torch.empty((B * n_tok, C), dtype=torch.int16, device=pe16.device)
    call("ivit_embed_tokens", context(pe16.device), ptr(pe16), ptr(cls32), ptr(pos16), B, n_tok, C,
         Dyadic(int(me[0]), int(me[1])), Dyadic(int(me_res[0]), int(me_res[1])), bits, ptr(out))
    return out


def shiftgelu_build_lut(x0: int, me: torch.Tensor, n: int = 23, bits: int = 8) -> torch.Tensor:
    """64 KiB composite table of IntGELU + scalar 8-bit QuantAct (built once per layer, on the device)."""
    lut = torch.empty(65536, dtype=torch.int8, device=me.device)
    call("ivit_shiftgelu_build_lut", context(me.device), int(x0), n, ptr(me), bits, ptr(lut))
    return lut


def shiftgelu_lut(q: torch.Tensor, lut: torch.Tensor, out: torch.Tensor = None):
    assert q.dtype == torch.int8 and q.is_contiguous()
    cols = q.shape[-1]
    if out is None:
        out = torch.empty_like(q)
    call("ivit_shiftgelu_lut", context(q.device), ptr(q), q.numel() // cols, cols, ptr(lut), ptr(out))
    return out


def layernorm_i16_i8(x: torch.Tensor, bias_int: torch.Tensor, me: torch.Tensor, out: torch.Tensor = None):
    assert x.dtype == torch.int16 and x.is_contiguous()
    Cc = x.shape[-1]
    if out is None:
        out = torch.empty(x.shape, dtype=torch.int8, device=x.device)
    call("ivit_layernorm_i16_i8", context(x.device), ptr(x), x.numel() // Cc, Cc, ptr(bias_int), ptr(me), ptr(out))
    return out


def quantize_patchify(img: torch.Tensor, scale: torch.Tensor, patch: int, out: torch.Tensor = None):
    """fp32 NCHW image -> int8 patch rows (input QuantAct + unfold in one pass)."""
    assert img.dtype == torch.float32 and img.dim() == 4 and img.is_contiguous()
    B, Cin, H, W = img.shape
    if out is None:
        out = torch.empty((B * (H // patch) * (W // patch), Cin * patch * patch), dtype=torch.int8, device=img.device)
    call("ivit_quantize_patchify", context(img.device), ptr(img), ptr(scale), B, Cin, H, W, patch, ptr(out))
    return out


def quantize_patchify_u8(img: torch.Tensor, mean: torch.Tensor, std: torch.Tensor, scale: torch.Tensor, patch: int,
                         out: torch.Tensor = None):
    """uint8 NCHW image -> int8 patch rows: ToTensor + Normalize (utils/data_utils.py:90-91) + input QuantAct + unfold."""
    assert img.dtype == torch.uint8 and img.dim() == 4 and img.is_contiguous()
    B, Cin, H, W = img.shape
    if out is None:
        out = torch.empty((B * (H // patch) * (W // patch), Cin * patch * patch), dtype=torch.int8, device=img.device)
    call("ivit_quantize_patchify_u8", context(img.device), ptr(img), ptr(mean), ptr(std), ptr(scale), B, Cin, H, W, patch, ptr(out))
    return out


def embed_tokens_fast(pe16, cls32, pos16, B: int, n_tok: int, C: int, me, me_res, out=None):
    if out is None:
        out = torch.empty((B * n_tok, C), dtype=torch.int16, device=pe16.device)
    call("ivit_embed_tokens_fast", context(pe16.device), ptr(pe16), ptr(cls32), ptr(pos16), B, n_tok, C,
         Dyadic(int(me[0]), int(me[1])), Dyadic(int(me_res[0]), int(me_res[1])), ptr(out))
    return out


# ---------------------------------------------------------------------------------------------------------------
# Swin hot path
# ---------------------------------------------------------------------------------------------------------------
def window_attention_i8(qkv: torch.Tensor, n_win: int, n_heads: int, me_s, me_s2, x0: int, me_o, bias_rq: torch.Tensor,
                        mask_bits: torch.Tensor = None, n_win_img: int = 0, mask_add: int = 0, out=None):
    """tcgen05 window attention (49 tokens, head_dim 32, 8-bit Shiftmax); raises IvitError(ENOTSUP) outside its domain."""
    assert qkv.dtype == torch.int8 and qkv.is_contiguous() and qkv.shape == (n_win * 49, 96 * n_heads)
    assert bias_rq.dtype == torch.int16 and bias_rq.is_contiguous() and bias_rq.shape == (n_heads, 49, 49)
    if out is None:
        out = torch.empty((n_win * 49, 32 * n_heads), dtype=torch.int8, device=qkv.device)
    p = _lib.WinAttnParams()
    p.n_win, p.n_heads, p.n_tok, p.head_dim = n_win, n_heads, 49, 32
    p.me_s = Dyadic(int(me_s[0]), int(me_s[1]))
    p.me_s2 = Dyadic(int(me_s2[0]), int(me_s2[1]))
    p.x0, p.n, p.p_bits = int(x0), 15, 8
    p.me_o = Dyadic(int(me_o[0]), int(me_o[1]))
    p.bias_rq = ptr(bias_rq)
    p.mask_bits = ptr(mask_bits)
    p.n_win_img, p.mask_add = int(n_win_img), int(mask_add)
    call("ivit_window_attention_i8", context(qkv.device), ptr(qkv), C.byref(p), ptr(out))
    return out


def layernorm_gather(x: torch.Tensor, rows_out: int, C_out: int, G: int, rowmap, L_out: int, L_in: int,
                     bias_int: torch.Tensor, me: torch.Tensor, out=None, xcopy=None):
    """IntLayerNorm + QuantAct over gathered rows (window permutation, G = 1; 2x2 patch merging, G = 4)."""
    assert x.dtype == torch.int16 and x.is_contiguous()
    if out is None:
        out = torch.empty((rows_out, C_out), dtype=torch.int8, device=x.device)
    call("ivit_layernorm_gather_i16_i8", context(x.device), ptr(x), rows_out, C_out, G, ptr(rowmap), L_out, L_in,
         ptr(bias_int), ptr(me), ptr(out), ptr(xcopy))
    return out


def avgpool_requant_i8(x: torch.Tensor, B: int, L: int, Cc: int, me, out=None):
    assert x.dtype == torch.int8 and x.is_contiguous()
    if out is None:
        out = torch.empty((B, Cc), dtype=torch.int8, device=x.device)
    call("ivit_avgpool_requant_i8", context(x.device), ptr(x), B, L, Cc, Dyadic(int(me[0]), int(me[1])), ptr(out))
    return out


def layernorm_i8_i16x2(x: torch.Tensor, bias_int: torch.Tensor, me: torch.Tensor, me2, out=None):
    """IntLayerNorm over int8 rows + two 16-bit QuantActs in a row (Swin patch embedding tail)."""
    assert x.dtype == torch.int8 and x.is_contiguous() and x.dim() == 2
    if out is None:
        out = torch.empty(x.shape, dtype=torch.int16, device=x.device)
    call("ivit_layernorm_i8_i16x2", context(x.device), ptr(x), x.shape[0], x.shape[1], ptr(bias_int), ptr(me),
         Dyadic(int(me2[0]), int(me2[1])), ptr(out))
    return out


def widen_i8_i16(x: torch.Tensor, out=None):
    """int8 -> int16 storage, values unchanged (PatchMerging output entering the int16 residual stream)."""
    assert x.dtype == torch.int8 and x.is_contiguous() and x.numel() % 16 == 0
    if out is None:
        out = torch.empty(x.shape, dtype=torch.int16, device=x.device)
    call("ivit_widen_i8_i16", context(x.device), ptr(x), x.numel(), ptr(out))
    return out
