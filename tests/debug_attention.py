"""Debug helper (not collected by pytest): localise mismatches of the fused attention kernel against the oracle (which
rows / heads / channel quarters), and dump per-row (max, sum) through IVIT_ATTN_DBG_PTR.  Lives under tests/ because it
uses the oracle:  python tests/debug_attention.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ivit_b200.kernels as K  # noqa: E402
import oracle as O  # noqa: E402
from test_kernels_gpu import oracle_attention  # noqa: E402


def run(n_seq, n_tok, H, D=64, p_bits=16, s=0.031):
    rng = np.random.default_rng(n_tok * 31 + H)
    qkv = rng.integers(-128, 128, (n_seq * n_tok, 3 * H * D)).astype(np.int8)
    qkv[::7, :H * D] = np.clip(qkv[::7, :H * D].astype(np.int32) * 3, -128, 127).astype(np.int8)
    s_attn = np.float32(s)
    acc_scale = np.float32(127 * s_attn / (D * 127 * 40))
    m_s, e_s = K.dyadic_host(np.array([acc_scale], np.float32), s_attn)
    x0 = O.x0_of(s_attn)
    m_o, e_o = K.dyadic_host(np.array([2.0 ** -(p_bits - 1) * 0.02], np.float32), np.float32(0.02 * 1.3))
    me_s, me_o = (int(m_s[0]), int(e_s[0])), (int(m_o[0]), int(e_o[0]))
    want = oracle_attention(qkv, n_seq, n_tok, H, D, me_s, x0, me_o, p_bits)
    got = K.attention_i8(torch.from_numpy(qkv).cuda(), n_seq, n_tok, H, D, me_s, x0, me_o, p_bits=p_bits).cpu().numpy().astype(np.int64)
    bad = got != want
    print("n_seq=%d n_tok=%d H=%d x0=%d: %d / %d mismatches, max |diff| %d" % (n_seq, n_tok, H, x0, bad.sum(), bad.size, np.abs(got - want).max()))
    if bad.any():
        b4 = bad.reshape(n_seq, n_tok, H, 4, D // 4)
        print("  per image      ", b4.sum(axis=(1, 2, 3, 4)).tolist())
        print("  per head       ", b4.sum(axis=(0, 1, 3, 4)).tolist())
        print("  per ch quarter ", b4.sum(axis=(0, 1, 2, 4)).tolist())
        rows = b4.sum(axis=(0, 2, 3, 4))
        print("  per 32-row group", [int(rows[i:i + 32].sum()) for i in range(0, n_tok, 32)])
        print("  rows with mismatches: %d of %d; first rows %s" % ((rows > 0).sum(), n_tok, np.nonzero(rows)[0][:20].tolist()))
        d = (got - want)[bad]
        print("  diff histogram ", {int(k): int(v) for k, v in zip(*np.unique(d, return_counts=True))})


def rowstats(n_tok, H=1, D=64, s=0.031):
    """Per-row (max, sum) of the kernel (IVIT_ATTN_DBG_PTR) against numpy."""
    rng = np.random.default_rng(n_tok * 31 + H)
    qkv = rng.integers(-128, 128, (n_tok, 3 * H * D)).astype(np.int8)
    s_attn = np.float32(s)
    acc_scale = np.float32(127 * s_attn / (D * 127 * 40))
    m_s, e_s = K.dyadic_host(np.array([acc_scale], np.float32), s_attn)
    x0 = O.x0_of(s_attn)
    m_o, e_o = K.dyadic_host(np.array([2.0 ** -15 * 0.02], np.float32), np.float32(0.02 * 1.3))
    me_s, me_o = (int(m_s[0]), int(e_s[0])), (int(m_o[0]), int(e_o[0]))
    n_mt = (n_tok + 127) // 128
    dbg = torch.zeros(H * n_mt * 128, dtype=torch.int64, device="cuda")
    os.environ["IVIT_ATTN_DBG_PTR"] = "%x" % dbg.data_ptr()
    K.attention_i8(torch.from_numpy(qkv).cuda(), 1, n_tok, H, D, me_s, x0, me_o, p_bits=16)
    torch.cuda.synchronize()
    del os.environ["IVIT_ATTN_DBG_PTR"]
    d = dbg.cpu().numpy().astype(np.uint64)
    blk = qkv.astype(np.int64)
    q, k = blk[:, :D], blk[:, H * D:H * D + D]
    sc = O.requant(q @ k.T, [me_s[0]], [me_s[1]], 8)
    mx = sc.max(axis=1)
    # E table through the oracle's shiftmax is not exposed; recompute int_exp_shift here (quant_modules.py:469-481)
    def E(dv):
        t = dv + (dv >> 1) - (dv >> 4)
        t = np.maximum(t, 15 * x0)
        kk = t // x0
        r = t - x0 * kk
        return ((r - 2 * x0) << np.maximum(15 - kk - 1, 0)) >> np.where(15 - kk - 1 < 0, 1, 0)
    S = E(sc - mx[:, None]).sum(axis=1)
    bad = 0
    for r in range(n_tok):
        v = int(d[(r // 128) * 128 + (r % 128)])
        gm, gs = (v >> 48) - 128, v & ((1 << 48) - 1)
        if gm != mx[r] or gs != S[r]:
            bad += 1
            if bad <= 6:
                print("   row %d: kernel max %d sum %d | numpy max %d sum %d | diff %d  E(pad)=%d" % (r, gm, gs, mx[r], S[r], gs - S[r], int(E(np.array(-128 - mx[r])))))
    print("n_tok=%d: %d / %d rows with a wrong (max, sum)" % (n_tok, bad, n_tok))


if __name__ == "__main__":
    for n in (128, 120, 127, 197):
        rowstats(n)
    for cfg in [(1, 120, 1), (1, 127, 1), (2, 197, 3), (1, 129, 2), (40, 197, 8), (3, 224, 2)]:
        run(*cfg)
