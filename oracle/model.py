"""Whole-model CPU oracle: the DeiT integer forward restated on the oracle primitives
(oracle/__init__.py) over a frozen parameter pack.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Call order follows the reference: vit_quant.py:254-282 (forward_features / forward),
:130-143 (Block), :59-88 (Attention); layers_quant.py:144-153 (Mlp), :184-196 (PatchEmbed).
``capture`` receives the integer tensor at every operator boundary under the reference's module
name, so that it can be compared with the reference-generated digests (tests/golden/).
"""
from __future__ import annotations

import concurrent.futures as cf
import os

import numpy as np

import oracle as O


def _me(pack, key):
    t = pack[key].astype(np.int64)
    return t[:, 0], t[:, 1]


def _bmm_nt(a, b):
    """[B,H,M,K] x [B,H,N,K]^T -> [B,H,M,N] through the C contraction (numpy has no fast int64 matmul)."""
    Bb, Hh, M, _ = a.shape
    N = b.shape[2]
    out = np.empty((Bb, Hh, M, N), np.int64)
    i8 = a.dtype == np.int8 and b.dtype == np.int8
    for i in range(Bb):
        for h in range(Hh):
            out[i, h] = O.gemm_nt(np.ascontiguousarray(a[i, h]), np.ascontiguousarray(b[i, h]))
    return out


def _linear(pack, name, x):
    return O.gemm_nt(x.astype(np.int8), pack[name + ".weight_integer"], pack[name + ".bias_integer"])


def deit_forward(pack, images: np.ndarray, capture: dict = None):
    """images: fp32 [B,3,H,W].  Returns fp32 logits [B, classes]."""
    mt = pack.meta
    C, H, D, N, P = mt["embed_dim"], mt["num_heads"], mt["head_dim"], mt["n_tok"], mt["patch"]
    B = images.shape[0]
    cap = capture if capture is not None else {}

    def rec(name, arr, shape=None):
        if capture is not None:
            cap[name] = arr.reshape(shape) if shape is not None else arr
        return arr

    q = O.quantize_f32(images, pack["qact_input.scale"][0], 8)                        # vit_quant.py:257
    rec("qact_input", q)
    Cin, Hh, Ww = q.shape[1:]
    patches = q.reshape(B, Cin, Hh // P, P, Ww // P, P).transpose(0, 2, 4, 1, 3, 5).reshape(-1, Cin * P * P)
    acc = _linear(pack, "patch_embed.proj", patches)                                  # layers_quant.py:190
    rec("patch_embed.proj", acc.reshape(B, Hh // P, Ww // P, C).transpose(0, 3, 1, 2))
    x = O.requant(acc, *_me(pack, "patch_embed.qact.me"), 16).reshape(B, N - 1, C)    # :195
    rec("patch_embed.qact", x)
    cls = np.broadcast_to(pack["cls_token_integer"].astype(np.int64), (B, 1, C))      # vit_quant.py:259-262
    x = np.concatenate([cls, x], axis=1).reshape(B * N, C)
    rec("qact_pos", pack["pos_embed_integer"].astype(np.int64), (1, N, C))            # :264
    x = O.requant(x, *_me(pack, "qact1.me"), 16, pack["pos_embed_integer"], *_me(pack, "qact1.me_res"))   # :265
    rec("qact1", x, (B, N, C))

    for i in range(mt["depth"]):
        p = "blocks.%d." % i
        x1 = x
        t = O.layernorm(x1, pack[p + "norm1.bias_integer"])                           # :131
        rec(p + "norm1", t, (B, N, C))
        t = O.requant(t, *_me(pack, p + "qact1.me"), 8)                               # :132
        rec(p + "qact1", t, (B, N, C))
        acc = _linear(pack, p + "attn.qkv", t)                                        # :61
        rec(p + "attn.qkv", acc, (B, N, 3 * C))
        qkv = O.requant(acc, *_me(pack, p + "attn.qact1.me"), 8)                      # :62
        rec(p + "attn.qact1", qkv, (B, N, 3 * C))
        qkv5 = qkv.reshape(B, N, 3, H, D).transpose(2, 0, 3, 1, 4)                    # :63-64
        qh, kh, vh = qkv5[0], qkv5[1], qkv5[2]                                        # [B,H,N,D]
        s = _bmm_nt(qh.astype(np.int8), kh.astype(np.int8))                           # :70-71
        rec(p + "attn.matmul_1", s)
        s = O.requant(s, *_me(pack, p + "attn.qact_attn1.me"), 8)                     # :74
        rec(p + "attn.qact_attn1", s)
        pr = O.shiftmax(s, int(pack[p + "attn.int_softmax.x0"][0]), mt["softmax_bits"])   # :76
        rec(p + "attn.int_softmax", pr)
        o = _bmm_nt(pr, vh.transpose(0, 1, 3, 2))                                     # :79-80
        rec(p + "attn.matmul_2", o)
        o = o.transpose(0, 2, 1, 3).reshape(B * N, C)                                 # :81
        o = O.requant(o, *_me(pack, p + "attn.qact2.me"), 8)                          # :83
        rec(p + "attn.qact2", o, (B, N, C))
        acc = _linear(pack, p + "attn.proj", o)                                       # :84
        rec(p + "attn.proj", acc, (B, N, C))
        a3 = O.requant(acc, *_me(pack, p + "attn.qact3.me"), 16)                      # :85
        rec(p + "attn.qact3", a3, (B, N, C))
        x2 = O.requant(a3, *_me(pack, p + "qact2.me"), 16, x1, *_me(pack, p + "qact2.me_res"))   # :135
        rec(p + "qact2", x2, (B, N, C))
        t = O.layernorm(x2, pack[p + "norm2.bias_integer"])                           # :137
        rec(p + "norm2", t, (B, N, C))
        t = O.requant(t, *_me(pack, p + "qact3.me"), 8)                               # :138
        rec(p + "qact3", t, (B, N, C))
        acc = _linear(pack, p + "mlp.fc1", t)                                         # layers_quant.py:145
        rec(p + "mlp.fc1", acc, (B, N, -1))
        g = O.requant(acc, *_me(pack, p + "mlp.qact_gelu.me"), 8)                     # :146
        rec(p + "mlp.qact_gelu", g, (B, N, -1))
        g = O.shiftgelu(g, int(pack[p + "mlp.act.x0"][0]))                            # :147
        rec(p + "mlp.act", g, (B, N, -1))
        g = O.requant(g, *_me(pack, p + "mlp.qact1.me"), 8)                           # :148
        rec(p + "mlp.qact1", g, (B, N, -1))
        acc = _linear(pack, p + "mlp.fc2", g)                                         # :150
        rec(p + "mlp.fc2", acc, (B, N, C))
        m2 = O.requant(acc, *_me(pack, p + "mlp.qact2.me"), 16)                       # :151
        rec(p + "mlp.qact2", m2, (B, N, C))
        x = O.requant(m2, *_me(pack, p + "qact4.me"), 16, x2, *_me(pack, p + "qact4.me_res"))   # vit_quant.py:141
        rec(p + "qact4", x, (B, N, C))

    t = O.layernorm(x, pack["norm.bias_integer"])                                     # :271
    rec("norm", t, (B, N, C))
    t = t.reshape(B, N, C)[:, 0]                                                      # :272
    t = O.requant(t, *_me(pack, "qact2.me"), 8)                                       # :273
    rec("qact2", t)
    acc = _linear(pack, "head", t)                                                    # :280
    rec("head", acc)
    assert np.abs(acc).max() < 2 ** 24, "head accumulator outside the exact fp32 range"
    return (acc.astype(np.float32) * pack["head.out_scale"][None, :]).astype(np.float32)


# --------------------------------------------------------------------------------------------
# Swin (swin_quant.py): windows, cyclic shift + mask, relative-position bias as a QuantAct identity, patch merging
# --------------------------------------------------------------------------------------------
def _win_part(x, ws):
    """[B,H,W,C] -> [B*nW, ws*ws, C]   swin_quant.py:18-32"""
    B, H, W, C = x.shape
    x = x.reshape(B, H // ws, ws, W // ws, ws, C).transpose(0, 1, 3, 2, 4, 5)
    return np.ascontiguousarray(x).reshape(-1, ws * ws, C)


def _win_rev(w, ws, H, W):
    """[B*nW, ws*ws, C] -> [B,H,W,C]   swin_quant.py:35-50"""
    C = w.shape[-1]
    B = w.shape[0] // ((H // ws) * (W // ws))
    x = w.reshape(B, H // ws, W // ws, ws, ws, C).transpose(0, 1, 3, 2, 4, 5)
    return np.ascontiguousarray(x).reshape(B, H, W, C)


def _masked_scores(q, scale_f32, mask01):
    """Integer input of IntSoftmax behind the shifted-window mask (swin_quant.py:151-156): the reference adds -100.0
    to the fp32 carrier q * s and the softmax divides the scale out again (quant_modules.py:484; under the exact-carrier
    hook: round(fp64(x) / fp64(s))).  Restated with the same fp32 / fp64 operations."""
    s32 = np.float32(scale_f32)
    x = (q.astype(np.float32) * s32).astype(np.float32)
    x = (x + (mask01.astype(np.float32) * np.float32(-100.0))).astype(np.float32)
    return np.rint(x.astype(np.float64) / np.float64(s32)).astype(np.int64)


def swin_forward(pack, images: np.ndarray, capture: dict = None):
    """images: fp32 [B,3,H,W].  Returns fp32 logits [B, classes]."""
    mt = pack.meta
    P, G = mt["patch"], mt["grid"]
    B = images.shape[0]
    cap = capture if capture is not None else {}

    def rec(name, arr, shape=None):
        if capture is not None:
            cap[name] = arr.reshape(shape) if shape is not None else arr
        return arr

    q = O.quantize_f32(images, pack["qact_input.scale"][0], 8)                        # swin_quant.py:540
    rec("qact_input", q)
    Cin = q.shape[1]
    C = mt["embed_dim"]
    patches = q.reshape(B, Cin, G, P, G, P).transpose(0, 2, 4, 1, 3, 5).reshape(-1, Cin * P * P)
    acc = _linear(pack, "patch_embed.proj", patches)                                  # layers_quant.py:190
    rec("patch_embed.proj", acc.reshape(B, G, G, C).transpose(0, 3, 1, 2))
    t = O.requant(acc, *_me(pack, "patch_embed.qact_before_norm.me"), 8)              # :193
    rec("patch_embed.qact_before_norm", t, (B, G * G, C))
    t = O.layernorm(t, pack["patch_embed.norm.bias_integer"])                         # :194
    rec("patch_embed.norm", t, (B, G * G, C))
    x = O.requant(t, *_me(pack, "patch_embed.qact.me"), 16)                           # :195
    rec("patch_embed.qact", x, (B, G * G, C))
    x = O.requant(x, *_me(pack, "qact1.me"), 16)                                      # swin_quant.py:546
    rec("qact1", x, (B, G * G, C))

    R = G
    for li, depth in enumerate(mt["depths"]):
        nH, ws = mt["num_heads"][li], mt["window"][li]
        N, D = ws * ws, C // nH
        L = R * R
        for bi in range(depth):
            p = "layers.%d.blocks.%d." % (li, bi)
            shift = mt["shift"][li][bi]
            x1 = x                                                                    # [B*L, C]
            t = O.layernorm(x1, pack[p + "norm1.bias_integer"])                       # :256
            rec(p + "norm1", t, (B, L, C))
            t = O.requant(t, *_me(pack, p + "qact1.me"), 8)                           # :257
            rec(p + "qact1", t, (B, L, C))
            t = t.reshape(B, R, R, C)
            if shift > 0:
                t = np.roll(t, (-shift, -shift), axis=(1, 2))                         # :261-265
            xw = _win_part(t, ws)                                                     # :269-271  [B_, N, C]
            B_ = xw.shape[0]
            acc = _linear(pack, p + "attn.qkv", xw.reshape(-1, C))                    # :128
            rec(p + "attn.qkv", acc, (B_, N, 3 * C))
            qkv = O.requant(acc, *_me(pack, p + "attn.qact1.me"), 8)                  # :129
            rec(p + "attn.qact1", qkv, (B_, N, 3 * C))
            qkv5 = qkv.reshape(B_, N, 3, nH, D).transpose(2, 0, 3, 1, 4)
            qh, kh, vh = qkv5[0], qkv5[1], qkv5[2]
            s = _bmm_nt(qh.astype(np.int8), kh.astype(np.int8))                       # :135-136
            rec(p + "attn.matmul_1", s)
            s = O.requant(s, *_me(pack, p + "attn.qact_attn1.me"), 8)                 # :140
            rec(p + "attn.qact_attn1", s)
            rec(p + "attn.qact_table", pack[p + "attn.qact_table.table_integer"].astype(np.int64))   # :142-143
            bias = pack[p + "attn.bias_integer"].astype(np.int64)                     # [nH, N, N]  :144-147
            s = O.requant(s.reshape(B_, nH * N, N), *_me(pack, p + "attn.qact2.me"), 8,
                          bias.reshape(nH * N, N), *_me(pack, p + "attn.qact2.me_res")).reshape(B_, nH, N, N)   # :149
            rec(p + "attn.qact2", s)
            if shift > 0:                                                             # :151-155
                m01 = pack[p + "attn_mask"]                                           # [nW, N, N]
                nW = m01.shape[0]
                s = _masked_scores(s.reshape(B_ // nW, nW, nH, N, N), pack[p + "attn.qact2.scale"][0],
                                   m01[None, :, None, :, :]).reshape(B_, nH, N, N)
            pr = O.shiftmax(s, int(pack[p + "attn.log_int_softmax.x0"][0]), mt["softmax_bits"])   # :156
            rec(p + "attn.log_int_softmax", pr)
            o = _bmm_nt(pr, vh.transpose(0, 1, 3, 2))                                 # :161-162
            rec(p + "attn.matmul_2", o)
            o = o.transpose(0, 2, 1, 3).reshape(B_ * N, C)                            # :163
            o = O.requant(o, *_me(pack, p + "attn.qact3.me"), 8)                      # :164
            rec(p + "attn.qact3", o, (B_, N, C))
            acc = _linear(pack, p + "attn.proj", o)                                   # :166
            rec(p + "attn.proj", acc, (B_, N, C))
            a4 = O.requant(acc, *_me(pack, p + "attn.qact4.me"), 16)                  # :167
            rec(p + "attn.qact4", a4, (B_, N, C))
            t = _win_rev(a4.reshape(B_, N, C), ws, R, R)                              # :278-281
            if shift > 0:
                t = np.roll(t, (shift, shift), axis=(1, 2))                           # :284-288
            t = np.ascontiguousarray(t).reshape(B * L, C)
            x2 = O.requant(t, *_me(pack, p + "qact2.me"), 16, x1, *_me(pack, p + "qact2.me_res"))   # :293
            rec(p + "qact2", x2, (B, L, C))
            t = O.layernorm(x2, pack[p + "norm2.bias_integer"])                       # :295
            rec(p + "norm2", t, (B, L, C))
            t = O.requant(t, *_me(pack, p + "qact3.me"), 8)                           # :296
            rec(p + "qact3", t, (B, L, C))
            acc = _linear(pack, p + "mlp.fc1", t)                                     # layers_quant.py:145
            rec(p + "mlp.fc1", acc, (B, L, -1))
            g = O.requant(acc, *_me(pack, p + "mlp.qact_gelu.me"), 8)                 # :146
            rec(p + "mlp.qact_gelu", g, (B, L, -1))
            g = O.shiftgelu(g, int(pack[p + "mlp.act.x0"][0]))                        # :147
            rec(p + "mlp.act", g, (B, L, -1))
            g = O.requant(g, *_me(pack, p + "mlp.qact1.me"), 8)                       # :148
            rec(p + "mlp.qact1", g, (B, L, -1))
            acc = _linear(pack, p + "mlp.fc2", g)                                     # :150
            rec(p + "mlp.fc2", acc, (B, L, C))
            m2 = O.requant(acc, *_me(pack, p + "mlp.qact2.me"), 16)                   # :151
            rec(p + "mlp.qact2", m2, (B, L, C))
            x = O.requant(m2, *_me(pack, p + "qact4.me"), 16, x2, *_me(pack, p + "qact4.me_res"))   # swin_quant.py:299
            rec(p + "qact4", x, (B, L, C))
        if li + 1 < len(mt["depths"]):                                                # PatchMerging :328-349
            d = "layers.%d.downsample." % li
            t = x.reshape(B, R, R, C)
            t = np.concatenate([t[:, 0::2, 0::2], t[:, 1::2, 0::2], t[:, 0::2, 1::2], t[:, 1::2, 1::2]], axis=-1)   # :337-341
            R //= 2
            t = np.ascontiguousarray(t).reshape(B * R * R, 4 * C)
            t = O.layernorm(t, pack[d + "norm.bias_integer"])                         # :344
            rec(d + "norm", t, (B, R * R, 4 * C))
            t = O.requant(t, *_me(pack, d + "qact1.me"), 8)                           # :345
            rec(d + "qact1", t, (B, R * R, 4 * C))
            acc = _linear(pack, d + "reduction", t)                                   # :346
            C *= 2
            rec(d + "reduction", acc, (B, R * R, C))
            x = O.requant(acc, *_me(pack, d + "qact2.me"), 8)                         # :347
            rec(d + "qact2", x, (B, R * R, C))

    L = R * R
    t = O.layernorm(x, pack["norm.bias_integer"])                                     # :552
    rec("norm", t, (B, L, C))
    t = O.requant(t, *_me(pack, "qact2.me"), 8)                                       # :553
    rec("qact2", t, (B, L, C))
    t = O.avgpool_rne(t.reshape(B, L, C))                                             # :554 (token average; see avgpool_rne)
    t = O.requant(t, *_me(pack, "qact3.me"), 8)                                       # :555
    rec("qact3", t, (B, C, 1))
    acc = _linear(pack, "head", t)                                                    # :562
    rec("head", acc)
    assert np.abs(acc).max() < 2 ** 24, "head accumulator outside the exact fp32 range"
    return (acc.astype(np.float32) * pack["head.out_scale"][None, :]).astype(np.float32)


def deit_forward_parallel(pack, images: np.ndarray, threads: int = None):
    """Batch-sharded over host threads (ctypes releases the GIL): the CPU baseline uses every core."""
    threads = threads or os.cpu_count() or 1
    B = images.shape[0]
    threads = max(1, min(threads, B))
    chunks = np.array_split(np.arange(B), threads)
    with cf.ThreadPoolExecutor(threads) as ex:
        outs = list(ex.map(lambda idx: deit_forward(pack, images[idx]), [c for c in chunks if len(c)]))
    return np.concatenate(outs, axis=0)
