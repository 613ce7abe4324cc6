"""CPU oracle of the CALIBRATION forward (one unfrozen pass, ``running_stat = True``): which activation ranges does the
exact-integer evaluation of the reference's formulas record on a batch?  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates ``QuantAct.forward`` in calibration mode (quant_modules.py:170-192: per-call min / max of the incoming carrier,
taken as the range on the first batch; the output scale derives from it, :191-192) around the integer operators of
oracle/__init__.py, in the call order of vit_quant.py:254-282, 130-143, 59-88 and layers_quant.py:144-153, 184-196.  Every
carrier is formed with the same IEEE operation the reference module uses (fp32 ``integer * scale``; fp64 behind
IntLayerNorm, whose integers exceed 24 bits -- the ranges are rounded to fp32 when stored, as ``QuantAct`` does), so the
ranges are a deterministic function of (weights, images) and a correct GPU implementation must reproduce them BIT FOR BIT
(tests/test_zz_reference_graphs_gpu.py).  Against the reference's own (literal fp32-carrier) calibration run the ranges
agree to within its carrier noise; tests/test_calib_oracle.py states the bound.
"""
from __future__ import annotations

import numpy as np

import oracle as O

F32 = np.float32
_EPS = np.finfo(np.float32).eps


def _sym_scale(bits, mn, mx):
    """quant_utils.py:51-69 (fp32, true division, eps clamp); mn / mx arrays or scalars."""
    n = F32(2 ** (bits - 1) - 1)
    s = (np.maximum(-np.asarray(mn, F32), np.asarray(mx, F32)) / n).astype(F32)
    return np.maximum(s, F32(_EPS)).astype(F32)


def _car(z, s):
    """fp32 carrier integer * scale (quant_modules.py:97, 206, 228, 445, 497); s scalar or per last-dim channel."""
    return (np.asarray(z).astype(F32) * np.asarray(s, F32)).astype(F32)


class _Calib:
    def __init__(self):
        self.ranges = {}

    def qact_input(self, name, bits, x):
        """QuantAct without an incoming scale (quant_modules.py:194-196): range of the fp32 tensor, quantise."""
        mn, mx = F32(x.min()), F32(x.max())
        self.ranges[name] = (float(mn), float(mx))
        s = _sym_scale(bits, mn, mx).reshape(1)
        return O.quantize_f32(x, s[0], bits), s

    def qact(self, name, bits, z, s_in, carrier, ident_z=None, ident_s=None, ident_carrier=None):
        """Requantising QuantAct (quant_modules.py:170-206).  carrier: what the module receives as x (fp32 or fp64);
        range = min / max of x (or identity + x, :171), stored as fp32; then fixedpoint_mul with the new scale."""
        x_act = carrier if ident_carrier is None else (ident_carrier + carrier)
        mn, mx = F32(x_act.min()), F32(x_act.max())
        self.ranges[name] = (float(mn), float(mx))
        s_out = _sym_scale(bits, mn, mx).reshape(1)
        m, e = O.dyadic(np.asarray(s_in, F32).reshape(-1), s_out[0])
        if ident_z is None:
            return O.requant(z, m, e, bits), s_out
        m1, e1 = O.dyadic(np.asarray(ident_s, F32).reshape(-1), s_out[0])
        return O.requant(z, m, e, bits, ident_z, m1, e1), s_out


def _linear(params, name, x_int, s_in):
    """QuantLinear / QuantConv2d (quant_modules.py:68-97, 305-330): per-row weight scale, int8 weights, int32 bias."""
    w = np.asarray(params[name + ".weight"], F32)
    v = w.reshape(w.shape[0], -1)
    s_w = _sym_scale(8, v.min(axis=1), v.max(axis=1))
    w_q = O.quantize_f32(v, s_w, 8, per_row=True)
    out_scale = (s_w * F32(np.asarray(s_in, F32).reshape(-1)[0])).astype(F32)
    b_q = None
    if name + ".bias" in params:
        b_q = O.quantize_f32(np.asarray(params[name + ".bias"], F32), out_scale, 32, per_row=True)
    return O.gemm_nt(x_int.astype(np.int8), w_q.astype(np.int8), b_q), out_scale


def _layernorm(params, name, x_int, C):
    """IntLayerNorm (quant_modules.py:353-386): integers + per-channel output scale; the carrier is fp64."""
    g = np.asarray(params[name + ".weight"], F32)
    b = np.asarray(params[name + ".bias"], F32)
    sf0 = (np.sqrt(F32(C)).astype(F32) / F32(2 ** 30)).astype(F32)
    bias_int = np.floor(((b / g).astype(F32) / sf0).astype(F32)).astype(np.int64)
    y = O.layernorm(x_int, bias_int)
    out_sf = (sf0 * g).astype(F32)
    return y, out_sf, y.astype(np.float64) * out_sf.astype(np.float64)


def _bmm_nt(a, b):
    Bb, Hh, M, _ = a.shape
    out = np.empty((Bb, Hh, M, b.shape[2]), np.int64)
    for i in range(Bb):
        for h in range(Hh):
            out[i, h] = O.gemm_nt(np.ascontiguousarray(a[i, h]), np.ascontiguousarray(b[i, h]))
    return out


def deit_calibrate(params: dict, meta: dict, images: np.ndarray) -> dict:
    """params: the float parameters by the reference's names (``model.state_dict()`` as numpy); meta: embed_dim, depth,
    num_heads, patch, n_tok; images: fp32 [B, 3, H, W].  Returns {QuantAct name: (min, max)} as fp32 values."""
    C, H, N, P, depth = meta["embed_dim"], meta["num_heads"], meta["n_tok"], meta["patch"], meta["depth"]
    D = C // H
    B = images.shape[0]
    cal = _Calib()
    q, s_img = cal.qact_input("qact_input", 8, np.asarray(images, F32))                         # vit_quant.py:257
    Cin, Hh, Ww = q.shape[1:]
    patches = q.reshape(B, Cin, Hh // P, P, Ww // P, P).transpose(0, 2, 4, 1, 3, 5).reshape(-1, Cin * P * P)
    acc, s_conv = _linear(params, "patch_embed.proj", patches, s_img)                            # layers_quant.py:190
    x16, s_pe = cal.qact("patch_embed.qact", 16, acc, s_conv, _car(acc, s_conv))                 # :195
    cls_f = np.asarray(params["cls_token"], F32).reshape(1, 1, C)
    x_car = np.concatenate([np.broadcast_to(cls_f, (B, 1, C)), _car(x16, s_pe).reshape(B, N - 1, C)], axis=1)   # vit_quant.py:259-262
    cls_z = np.rint((cls_f.reshape(-1) / s_pe[0]).astype(F32)).astype(np.int64)
    z_cat = np.concatenate([np.broadcast_to(cls_z, (B, 1, C)), x16.reshape(B, N - 1, C)], axis=1).reshape(B * N, C)
    pos_q, s_pos = cal.qact_input("qact_pos", 16, np.asarray(params["pos_embed"], F32).reshape(1, N, C))   # :264
    x, s_x = cal.qact("qact1", 16, z_cat, s_pe, x_car.reshape(B * N, C), pos_q.reshape(N, C), s_pos,
                      np.broadcast_to(_car(pos_q, s_pos), (B, N, C)).reshape(B * N, C))          # :265

    for i in range(depth):
        p = "blocks.%d." % i
        x1, s_x1 = x, s_x
        y, sf, car64 = _layernorm(params, p + "norm1", x1, C)                                    # :131
        t, s_t = cal.qact(p + "qact1", 8, y, sf, car64)                                          # :132
        acc, s_acc = _linear(params, p + "attn.qkv", t, s_t)                                     # :61
        qkv, s_qkv = cal.qact(p + "attn.qact1", 8, acc, s_acc, _car(acc, s_acc))                 # :62
        qkv5 = qkv.reshape(B, N, 3, H, D).transpose(2, 0, 3, 1, 4)
        qh, kh, vh = qkv5[0], qkv5[1], qkv5[2]
        S = _bmm_nt(qh.astype(np.int8), kh.astype(np.int8))                                      # :70-71
        s_mm = (s_qkv * s_qkv).astype(F32)                                                       # quant_modules.py:226
        scale = F32(D ** -0.5)
        attn_car = (_car(S, s_mm) * scale).astype(F32)                                           # vit_quant.py:72
        s_sc = (s_mm * scale).astype(F32)                                                        # :73
        z = np.rint((attn_car / s_sc).astype(F32)).astype(np.int64)                              # what fixedpoint_mul recovers
        sc8, s_attn = cal.qact(p + "attn.qact_attn1", 8, z, s_sc, attn_car)                      # :74
        pr = O.shiftmax(sc8, O.x0_of(s_attn[0]), 16)                                             # :76
        s_p = F32(1.0 / 2 ** 15)
        o = _bmm_nt(pr, vh.transpose(0, 1, 3, 2))                                                # :79-80
        s_pv = (s_p * s_qkv).astype(F32)
        o = o.transpose(0, 2, 1, 3).reshape(B * N, C)                                            # :81
        o8, s_o = cal.qact(p + "attn.qact2", 8, o, s_pv, _car(o, s_pv))                          # :83
        acc, s_acc = _linear(params, p + "attn.proj", o8, s_o)                                   # :84
        a3, s_a3 = cal.qact(p + "attn.qact3", 16, acc, s_acc, _car(acc, s_acc))                  # :85
        x2, s_x2 = cal.qact(p + "qact2", 16, a3, s_a3, _car(a3, s_a3), x1, s_x1, _car(x1, s_x1))   # :135
        y, sf, car64 = _layernorm(params, p + "norm2", x2, C)                                    # :137
        t, s_t = cal.qact(p + "qact3", 8, y, sf, car64)                                          # :138
        acc, s_acc = _linear(params, p + "mlp.fc1", t, s_t)                                      # layers_quant.py:145
        g, s_g = cal.qact(p + "mlp.qact_gelu", 8, acc, s_acc, _car(acc, s_acc))                  # :146
        gy = O.shiftgelu(g, O.x0_of(O.gelu_sig_scale(s_g[0])))                                   # :147
        s_go = (s_g * F32(1.0 / 2 ** 7)).astype(F32)                                             # quant_modules.py:440-443
        g8, s_g8 = cal.qact(p + "mlp.qact1", 8, gy, s_go, _car(gy, s_go))                        # :148
        acc, s_acc = _linear(params, p + "mlp.fc2", g8, s_g8)                                    # :150
        m2, s_m2 = cal.qact(p + "mlp.qact2", 16, acc, s_acc, _car(acc, s_acc))                   # :151
        x, s_x = cal.qact(p + "qact4", 16, m2, s_m2, _car(m2, s_m2), x2, s_x2, _car(x2, s_x2))   # vit_quant.py:141

    y, sf, car64 = _layernorm(params, "norm", x, C)                                              # :271
    y0 = y.reshape(B, N, C)[:, 0]                                                                # :272
    cal.qact("qact2", 8, y0, sf, car64.reshape(B, N, C)[:, 0])                                   # :273
    return cal.ranges
