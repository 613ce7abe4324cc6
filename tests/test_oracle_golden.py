"""The CPU oracle (oracle/ivit_oracle.c) must reproduce, bit for bit, the vectors obtained by
executing the reference's own modules (tests/golden/make_golden.py).  This is what PINS it."""
import numpy as np
import pytest

import oracle as O


def norm_me(m, e):
    """(2^31, e) == (2^30, e-1)"""
    m, e = m.copy(), e.copy()
    big = np.abs(m) == 2 ** 31
    m[big] //= 2
    e[big] -= 1
    return m, e


def test_batch_frexp(kat):
    m, e = O.dyadic(kat["frexp_s_in"], kat["frexp_s_out"])
    assert np.array_equal(m, kat["frexp_m"]) and np.array_equal(e, kat["frexp_e"])
    # |m| == 2^31 needs a mantissa >= 1 - 2^-33; a ratio of two fp32 scales tops out at
    # 1 - 2^-24, so it is unreachable from the model (norm_me() handles it anyway).
    assert (m < 0).any()


@pytest.mark.parametrize("bits", [8, 16, 32])
def test_sym_scale(kat, bits):
    got = np.array([O.sym_scale(bits, a, b) for a, b in zip(kat["symscale_min"], kat["symscale_max"])], np.float32)
    assert np.array_equal(got, kat["symscale_b%d" % bits])


@pytest.mark.parametrize("bits", [8, 16])
def test_quantize_input(kat, bits):
    lo, hi = kat["qin%d_range" % bits]
    s = O.sym_scale(bits, lo, hi)
    assert s == kat["qin%d_sf" % bits].reshape(-1)[0]
    q = O.quantize_f32(kat["qin%d_x" % bits], s, bits)
    assert np.array_equal(q, kat["qin%d_q" % bits])
    n = 2 ** (bits - 1) - 1
    assert q.max() == n or q.min() == -n - 1, "KAT should exercise the clamp"


def test_requant_cases(kat):
    for ci, bits, perch, resid in kat["rq_cases"]:
        z, s_in, sf = kat["rq%d_z" % ci], kat["rq%d_s_in" % ci], kat["rq%d_sf" % ci].reshape(-1)[0]
        m, e = O.dyadic(s_in, sf)
        if resid:
            m1, e1 = O.dyadic(kat["rq%d_s_id" % ci], sf)
            got, diff = O.requant(z, m, e, int(bits), kat["rq%d_w" % ci], m1, e1, return_diff=True)
        else:
            got, diff = O.requant(z, m, e, int(bits), return_diff=True)
        assert np.array_equal(got, kat["rq%d_q" % ci]), "requant case %d" % ci
        assert diff == 0, "fp64 product rounding changed a result (case %d)" % ci
        # normalised (m fits int32) form is equivalent
        mn, en = norm_me(m, e)
        assert np.abs(mn).max() < 2 ** 31 or (mn == -2 ** 31).any()
        if not resid:
            assert np.array_equal(O.requant(z, mn, en, int(bits)), kat["rq%d_q" % ci])


def test_requant_ties_exercised(kat):
    ci = [c for c in kat["rq_cases"] if c[0] == 8][0][0]
    z, s_in, sf = kat["rq%d_z" % ci], kat["rq%d_s_in" % ci], kat["rq%d_sf" % ci].reshape(-1)[0]
    m, e = O.dyadic(s_in, sf)
    frac = (z * m) % (2 ** e)
    assert (frac * 2 == 2 ** e).sum() > 10, "tie case should contain exact .5 ties"


def test_quant_linear(kat):
    w, b, a, s_a = kat["lin_w"], kat["lin_b"], kat["lin_a"], kat["lin_s_a"]
    s_w = np.array([O.sym_scale(8, r.min(), r.max()) for r in w], np.float32)
    wq = O.quantize_f32(w, s_w, 8, per_row=True)
    assert np.array_equal(wq, kat["lin_wq"])
    s_b = (s_w * np.float32(s_a)).astype(np.float32)
    assert np.array_equal(s_b, kat["lin_sf"].reshape(-1))
    bq = O.quantize_f32(b, s_b, 32, per_row=True)
    assert np.array_equal(bq, kat["lin_bq"])
    acc = O.gemm_nt(a, wq, bq)
    assert np.array_equal(acc, kat["lin_acc"])
    acc8 = O.gemm_nt(a.astype(np.int8), wq.astype(np.int8), bq)
    assert np.array_equal(acc8, kat["lin_acc"])


def test_quant_conv(kat):
    w, b, x, s = kat["conv_w"], kat["conv_b"], kat["conv_x"], kat["conv_s"]
    Cout = w.shape[0]
    wf = w.reshape(Cout, -1)
    s_w = np.array([O.sym_scale(8, r.min(), r.max()) for r in wf], np.float32)
    wq = O.quantize_f32(wf, s_w, 8, per_row=True)
    assert np.array_equal(wq.reshape(w.shape), kat["conv_wq"])
    s_b = (s_w * np.float32(s)).astype(np.float32)
    assert np.array_equal(s_b, kat["conv_sf"])
    bq = O.quantize_f32(b, s_b, 32, per_row=True)
    assert np.array_equal(bq, kat["conv_bq"])
    B, Cin, H, W = x.shape
    k = w.shape[2]
    # unfold non-overlapping k x k patches -> [B*Hp*Wp, Cin*k*k] (conv == GEMM, quant_modules.py:329)
    p = x.reshape(B, Cin, H // k, k, W // k, k).transpose(0, 2, 4, 1, 3, 5).reshape(-1, Cin * k * k)
    acc = O.gemm_nt(p, wq, bq).reshape(B, H // k, W // k, Cout).transpose(0, 3, 1, 2)
    assert np.array_equal(acc, kat["conv_acc"])


def test_quant_matmul(kat):
    A, B = kat["mm_A"], kat["mm_B"]
    for i in range(A.shape[0]):
        for h in range(A.shape[1]):
            acc = O.gemm_nt(A[i, h], B[i, h].T.copy())
            assert np.array_equal(acc, kat["mm_acc"][i, h])
            acc2 = O.gemm_nt(kat["mm2_P"][i, h], kat["mm2_V"][i, h].T.copy())
            assert np.array_equal(acc2, kat["mm2_acc"][i, h])
    assert np.float32(kat["mm_sA"]) * np.float32(kat["mm_sB"]) == kat["mm_sf"].reshape(-1)[0]


@pytest.mark.parametrize("tag,bits", [("sm16", 16), ("sm8", 8), ("sm16b", 16), ("sm8b", 8)])
def test_shiftmax(kat, tag, bits):
    x0 = O.x0_of(kat[tag + "_s"])
    p = O.shiftmax(kat[tag + "_q"], x0, bits)
    assert np.array_equal(p, kat[tag + "_p"])
    assert p.min() >= 0 and p.max() <= 2 ** (bits - 1)


@pytest.mark.parametrize("tag,bits", [("sm16_x1", 16), ("sm8_x2", 8), ("sm16_x5", 16), ("sm8_fine", 8), ("sm16_fine", 16)])
def test_shiftmax_extreme_scales(kat, tag, bits):
    """Coarse (x0 = -2 ... -6) and very fine (x0 ~ -1e6: the 2^31-1 clamp of the sum bites and the reference's own
    result leaves [0, 2^(bits-1)]) input scales -- vectors generated by the reference's IntSoftmax."""
    x0 = O.x0_of(kat[tag + "_s"])
    assert np.array_equal(O.shiftmax(kat[tag + "_q"], x0, bits), kat[tag + "_p"])


@pytest.mark.parametrize("tag", ["gelu_x7", "gelu_x3", "gelu_x2", "gelu_x1", "gelu_fine"])
def test_shiftgelu_extreme_scales(kat, tag):
    """Pre-GELU scales the round-1 kernels refused (x0 in [-7, -1]: e^(-x_max) of an all-negative row is 2^(23+184))
    and a very fine one -- vectors generated by the reference's IntGELU."""
    x0 = O.x0_of(O.gelu_sig_scale(kat[tag + "_s"]))
    assert np.array_equal(O.shiftgelu(kat[tag + "_q"], x0), kat[tag + "_o"])


@pytest.mark.parametrize("tag", ["gelu_a", "gelu_b", "gelu_c"])
def test_shiftgelu(kat, tag):
    s = kat[tag + "_s"]
    x0 = O.x0_of(O.gelu_sig_scale(s))
    o = O.shiftgelu(kat[tag + "_q"], x0)
    assert np.array_equal(o, kat[tag + "_o"])
    assert np.float32(s) * np.float32(1 / 128) == kat[tag + "_sf"].reshape(-1)[0]


@pytest.mark.parametrize("tag", ["ln_a", "ln_b", "ln_c", "ln_d"])
def test_layernorm(kat, tag):
    q, g, beta = kat[tag + "_q"], kat[tag + "_g"], kat[tag + "_beta"]
    C = q.shape[1]
    assert kat[tag + "_rowsum_mod"][0] == C // 2, "KAT row 3 should be an exact .5 mean tie"
    sf0 = np.float32(np.sqrt(np.float32(C))) / np.float32(2 ** 30)
    bq = np.floor((beta / g).astype(np.float32) / sf0).astype(np.int64)
    assert np.array_equal(bq, kat[tag + "_bq"])
    o = O.layernorm(q, bq)
    assert np.array_equal(o, kat[tag + "_o"])
    assert np.array_equal((sf0 * g).astype(np.float32), kat[tag + "_sf"].reshape(-1))
