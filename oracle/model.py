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


def deit_forward_parallel(pack, images: np.ndarray, threads: int = None):
    """Batch-sharded over host threads (ctypes releases the GIL): the CPU baseline uses every core."""
    threads = threads or os.cpu_count() or 1
    B = images.shape[0]
    threads = max(1, min(threads, B))
    chunks = np.array_split(np.arange(B), threads)
    with cf.ThreadPoolExecutor(threads) as ex:
        outs = list(ex.map(lambda idx: deit_forward(pack, images[idx]), [c for c in chunks if len(c)]))
    return np.concatenate(outs, axis=0)
