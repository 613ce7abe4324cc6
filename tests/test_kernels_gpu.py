"""GPU parity tests: every entry point of the C ABI (include/ivit_b200.h) against the CPU oracle
(oracle/, pinned to the reference by tests/test_oracle_golden.py) on identical seeded inputs.
Bit-exact: all of this is integer arithmetic (the two fp32 outputs, carrier and logits, are a
single IEEE fp32 multiply and must match bitwise too)."""
import numpy as np
import pytest
import torch

import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    import ivit_b200.kernels as k
    return k


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def rand_me(rng, n, e_lo=32, e_hi=44, neg_every=0):
    m = rng.integers(2 ** 30, 2 ** 31, n).astype(np.int64)
    if neg_every:
        m[::neg_every] *= -1
    e = rng.integers(e_lo, e_hi + 1, n).astype(np.int64)
    return m, e


def me_dev(K, m, e):
    return K.dyadic_table(m, e, "cuda")


def assert_equal(got, want, what):
    got = got.cpu().numpy().astype(np.int64) if isinstance(got, torch.Tensor) else np.asarray(got)
    want = np.asarray(want).astype(np.int64)
    if got.shape != want.shape:
        got = got.reshape(want.shape)
    bad = np.argwhere(got != want)
    if len(bad):
        i = tuple(bad[0])
        pytest.fail("%s: %d / %d mismatches, first at %s: got %d want %d" %
                    (what, len(bad), want.size, i, got[i], want[i]))


# ------------------------------------------------------------------------------- primitives
def test_dyadic_device_and_host(K, kat):
    m_ref, e_ref = O.dyadic(kat["frexp_s_in"], kat["frexp_s_out"])
    big = np.abs(m_ref) == 2 ** 31
    m_n = np.where(big, m_ref // 2, m_ref)
    e_n = np.clip(np.where(big, e_ref - 1, e_ref), -1, 63)
    mh, eh = K.dyadic_host(kat["frexp_s_in"], kat["frexp_s_out"])
    assert np.array_equal(mh, m_n) and np.array_equal(eh, e_n)
    t = K.dyadic_device(dev(kat["frexp_s_in"]), dev(np.array([kat["frexp_s_out"]], np.float32))).cpu().numpy()
    assert np.array_equal(t[:, 0], m_n) and np.array_equal(t[:, 1], e_n)


@pytest.mark.parametrize("bits", [8, 16])
def test_quantize_f32(K, kat, bits):
    lo, hi = kat["qin%d_range" % bits]
    s = O.sym_scale(bits, lo, hi)
    q = K.quantize_f32(dev(kat["qin%d_x" % bits]), dev(np.array([s], np.float32)), bits)
    assert_equal(q, kat["qin%d_q" % bits], "quantize_f32 b%d" % bits)
    # large vectorised case
    rng = np.random.default_rng(bits)
    x = (rng.standard_normal((3, 3, 32, 32)) * 2).astype(np.float32)
    s2 = np.float32(0.0173)
    assert_equal(K.quantize_f32(dev(x), dev(np.array([s2])), 8), O.quantize_f32(x, s2, 8), "quantize_f32 vec")


def test_quantize_weights_per_row(K, kat):
    w = kat["lin_w"]
    s_w = np.array([O.sym_scale(8, r.min(), r.max()) for r in w], np.float32)
    q = K.quantize_f32(dev(w), dev(s_w), 8, per_row=True)
    assert_equal(q, kat["lin_wq"], "weight quantisation")
    s_b = (s_w * np.float32(kat["lin_s_a"])).astype(np.float32)
    qb = K.quantize_f32(dev(kat["lin_b"]), dev(s_b), 32, per_row=True)
    assert_equal(qb, kat["lin_bq"], "bias quantisation")


def test_carrier_roundtrip(K):
    rng = np.random.default_rng(3)
    q = rng.integers(-30000, 30000, (37, 48)).astype(np.int32)
    s = rng.uniform(1e-4, 1e-2, 48).astype(np.float32)
    x = K.int_to_carrier(dev(q), dev(s))
    assert np.array_equal(x.cpu().numpy(), (q.astype(np.float32) * s[None, :]).astype(np.float32))
    assert_equal(K.carrier_to_int(x, dev(s), torch.int32), q, "carrier roundtrip per-channel")
    x1 = K.int_to_carrier(dev(q.astype(np.int16)), dev(s[:1]))
    assert_equal(K.carrier_to_int(x1, dev(s[:1]), torch.int16), q, "carrier roundtrip scalar")


def test_requant_kat(K, kat):
    for ci, bits, perch, resid in kat["rq_cases"]:
        z, s_in, sf = kat["rq%d_z" % ci], kat["rq%d_s_in" % ci], kat["rq%d_sf" % ci].reshape(-1)[0]
        m, e = K.dyadic_host(s_in, sf)
        zt = dev(z.astype(np.int32))
        if resid:
            m1, e1 = K.dyadic_host(kat["rq%d_s_id" % ci], sf)
            got = K.requant(zt, me_dev(K, m, e), int(bits), dev(kat["rq%d_w" % ci].astype(np.int16)), me_dev(K, m1, e1))
        else:
            got = K.requant(zt, me_dev(K, m, e), int(bits))
        assert_equal(got, kat["rq%d_q" % ci], "requant KAT case %d" % ci)


def test_requant_random_edges(K):
    rng = np.random.default_rng(11)
    rows, cols = 64, 96
    z = rng.integers(-2 ** 31, 2 ** 31, (rows, cols)).astype(np.int64)
    z[0, :8] = [0, 1, -1, 2 ** 31 - 1, -2 ** 31, 2 ** 30, -2 ** 30, 5]
    m, e = rand_me(rng, cols, e_lo=-1, e_hi=63, neg_every=3)
    m[:4] = [2 ** 30, -2 ** 30, 2 ** 31 - 1, -2 ** 31]
    for bits, dt in [(8, np.int8), (16, np.int16), (32, np.int32)]:
        want = O.requant(z, m, e, bits)
        got = K.requant(dev(z.astype(np.int32)), me_dev(K, m, e), bits)
        assert_equal(got, want, "requant edges b%d" % bits)
    # broadcast residual row (pos_embed, vit_quant.py:265)
    w = rng.integers(-32768, 32768, (1, cols)).astype(np.int64)
    m1, e1 = rand_me(rng, 1, 28, 36)
    z16 = rng.integers(-32768, 32768, (rows, cols)).astype(np.int64)
    ms, es = rand_me(rng, 1, 28, 36)
    want = O.requant(z16, ms, es, 16, w, m1, e1)
    got = K.requant(dev(z16.astype(np.int16)), me_dev(K, ms, es), 16, dev(w.astype(np.int16)), me_dev(K, m1, e1))
    assert_equal(got, want, "requant broadcast residual")


# ------------------------------------------------------------------------------- row operators
@pytest.mark.parametrize("tag", ["ln_a", "ln_b", "ln_c", "ln_d"])
def test_layernorm_kat(K, kat, tag):
    q, bq = kat[tag + "_q"], kat[tag + "_bq"]
    dt = np.int16 if np.abs(q).max() < 32768 else np.int32     # the tie row of ln_d leaves the int16 range
    got = K.layernorm(dev(q.astype(dt)), dev(bq.astype(np.int32)))
    assert_equal(got, kat[tag + "_o"], "layernorm " + tag)
    # fused per-channel QuantAct (8 bit), negative multipliers where gamma < 0
    sf = kat[tag + "_sf"].reshape(-1)
    s_out = np.float32(np.abs(kat[tag + "_o"].astype(np.float64) * sf.astype(np.float64)).max() / 127.0)
    m, e = K.dyadic_host(sf, s_out)
    want = O.requant(kat[tag + "_o"], m, e, 8)
    got = K.layernorm(dev(q.astype(dt)), dev(bq.astype(np.int32)), me_dev(K, m, e), 8)
    assert_equal(got, want, "layernorm+qact " + tag)


def test_layernorm_random_widths(K):
    rng = np.random.default_rng(5)
    for C, dt, mag in [(96, np.int8, 127), (192, np.int16, 20000), (384, np.int16, 32767), (768, np.int16, 9000),
                       (1536, np.int8, 127), (100, np.int16, 500)]:
        q = rng.integers(-mag, mag + 1, (67, C)).astype(np.int64)
        bq = rng.integers(-2 ** 24, 2 ** 24, C).astype(np.int64)
        want = O.layernorm(q, bq)
        got = K.layernorm(dev(q.astype(dt)), dev(bq.astype(np.int32)))
        assert_equal(got, want, "layernorm C=%d" % C)


@pytest.mark.parametrize("tag,bits", [("sm16", 16), ("sm8", 8), ("sm16b", 16), ("sm8b", 8)])
def test_shiftmax_kat(K, kat, tag, bits):
    x0 = O.x0_of(kat[tag + "_s"])
    got = K.shiftmax(dev(kat[tag + "_q"].astype(np.int8)), x0, bits)
    assert_equal(got, kat[tag + "_p"], "shiftmax " + tag)


def test_shiftmax_random(K):
    rng = np.random.default_rng(6)
    for cols, s, bits in [(197, 0.02, 16), (49, 0.004, 8), (197, 0.11, 16), (64, 0.0007, 16), (300, 0.05, 8)]:
        q = rng.integers(-128, 128, (50, cols)).astype(np.int64)
        x0 = O.x0_of(np.float32(s))
        assert_equal(K.shiftmax(dev(q.astype(np.int8)), x0, bits), O.shiftmax(q, x0, bits), "shiftmax cols=%d s=%g" % (cols, s))


@pytest.mark.parametrize("tag", ["gelu_a", "gelu_b", "gelu_c"])
def test_shiftgelu_kat(K, kat, tag):
    s = kat[tag + "_s"]
    x0 = O.x0_of(O.gelu_sig_scale(s))
    got = K.shiftgelu(dev(kat[tag + "_q"].astype(np.int8)), x0)
    assert_equal(got, kat[tag + "_o"], "shiftgelu " + tag)
    # fused scalar QuantAct (mlp.qact1, layers_quant.py:148)
    s_in = np.float32(s) * np.float32(1 / 128)
    s_out = np.float32(np.abs(kat[tag + "_o"]).max() * float(s_in) / 127.0)
    m, e = K.dyadic_host(np.array([s_in], np.float32), s_out)
    want = O.requant(kat[tag + "_o"], m, e, 8)
    got = K.shiftgelu(dev(kat[tag + "_q"].astype(np.int8)), x0, me_dev(K, m, e), 8)
    assert_equal(got, want, "shiftgelu+qact " + tag)


def test_shiftgelu_random(K):
    rng = np.random.default_rng(8)
    for cols, s in [(768, 0.03), (3072, 0.045), (384, 0.012), (1536, 0.07)]:
        q = rng.integers(-128, 128, (33, cols)).astype(np.int64)
        q[0] = rng.integers(-128, -1, cols)
        x0 = O.x0_of(O.gelu_sig_scale(np.float32(s)))
        assert_equal(K.shiftgelu(dev(q.astype(np.int8)), x0, out_dtype=torch.int32), O.shiftgelu(q, x0),
                     "shiftgelu cols=%d" % cols)


@pytest.mark.parametrize("tag,bits", [("sm16_x1", 16), ("sm8_x2", 8), ("sm16_x5", 16), ("sm8_fine", 8), ("sm16_fine", 16)])
def test_shiftmax_kat_extreme_scales(K, kat, tag, bits):
    x0 = O.x0_of(kat[tag + "_s"])
    got = K.shiftmax(dev(kat[tag + "_q"].astype(np.int8)), x0, bits, out_dtype=torch.int32)
    assert_equal(got, kat[tag + "_p"], "shiftmax " + tag)


@pytest.mark.parametrize("tag", ["gelu_x7", "gelu_x3", "gelu_x2", "gelu_x1", "gelu_fine"])
def test_shiftgelu_kat_extreme_scales(K, kat, tag):
    """Scales the round-1 kernels refused (x0 > -8), vectors from the reference's IntGELU; the fused QuantAct
    form and the 64 KiB table form must agree with the oracle as well."""
    s = kat[tag + "_s"]
    x0 = O.x0_of(O.gelu_sig_scale(s))
    q = kat[tag + "_q"]
    got = K.shiftgelu(dev(q.astype(np.int8)), x0, out_dtype=torch.int32)
    assert_equal(got, kat[tag + "_o"], "shiftgelu " + tag)
    s_in = np.float32(s) * np.float32(1 / 128)
    s_out = np.float32(np.abs(kat[tag + "_o"]).max() * float(s_in) / 127.0)
    m, e = K.dyadic_host(np.array([s_in], np.float32), s_out)
    want = O.requant(kat[tag + "_o"], m, e, 8)
    me = me_dev(K, m, e)
    assert_equal(K.shiftgelu(dev(q.astype(np.int8)), x0, me, 8), want, "shiftgelu+qact " + tag)
    lut = K.shiftgelu_build_lut(x0, me)
    assert_equal(K.shiftgelu_lut(dev(np.ascontiguousarray(np.tile(q, (1, 2))).astype(np.int8)), lut),
                 np.tile(want, (1, 2)), "shiftgelu LUT " + tag)          # row max unchanged by tiling; cols % 16 == 0


@pytest.mark.parametrize("s", [1e-4, 6e-4, 2.3e-3, 0.009, 0.0734, 0.12, 0.3, 0.59, 1.0])
def test_shiftgelu_and_shiftmax_over_scales(K, s):
    """Property sweep over input scales 1e-4 ... 1: general kernels == oracle, no refusal."""
    rng = np.random.default_rng(int(s * 1e6))
    q = rng.integers(-128, 128, (40, 208)).astype(np.int64)
    q[0] = rng.integers(-128, -1, 208)
    q[1] = -128
    q[2] = 127
    x0g = O.x0_of(O.gelu_sig_scale(np.float32(s)))
    assert_equal(K.shiftgelu(dev(q.astype(np.int8)), x0g, out_dtype=torch.int32), O.shiftgelu(q, x0g), "shiftgelu s=%g" % s)
    x0s = O.x0_of(np.float32(s))
    for bits in (8, 16):
        assert_equal(K.shiftmax(dev(q.astype(np.int8)), x0s, bits, out_dtype=torch.int32), O.shiftmax(q, x0s, bits),
                     "shiftmax s=%g bits=%d" % (s, bits))


def test_unsupported_arguments_fail_loudly(K):
    from ivit_b200._lib import IvitError
    q = torch.zeros((4, 64), dtype=torch.int8, device="cuda")
    with pytest.raises(IvitError, match="x0"):
        K.shiftgelu(q, 0)
    with pytest.raises(IvitError, match="int32"):
        K.shiftgelu(q, -4000)                              # un-fused int16 output cannot hold the result below 2^-8 scales
    with pytest.raises(IvitError, match="int32"):
        K.shiftmax(q, -(1 << 20), 8)
    with pytest.raises(IvitError, match="bits"):
        K.requant(q.int(), me_dev(K, [2 ** 30], [31]), 7, out_dtype=torch.int32)


def test_patchify(K):
    rng = np.random.default_rng(9)
    x = rng.integers(-128, 128, (3, 3, 32, 48)).astype(np.int8)
    p = 16
    got = K.patchify_i8(dev(x), p).cpu().numpy()
    B, Cin, H, W = x.shape
    want = x.reshape(B, Cin, H // p, p, W // p, p).transpose(0, 2, 4, 1, 3, 5).reshape(-1, Cin * p * p)
    assert np.array_equal(got, want)


# ------------------------------------------------------------------------------- GEMM (tcgen05)
GEMM_SHAPES = [(128, 128, 128), (128, 256, 128), (256, 128, 256), (128, 256, 768), (197, 192, 192),
               (394, 576, 192), (300, 768, 768), (130, 1000, 768), (2, 1000, 192), (640, 3072, 768),
               (513, 768, 3072), (100, 96, 48), (1300, 2304, 768)]


def gemm_inputs(rng, M, N, K_):
    a = rng.integers(-128, 128, (M, K_)).astype(np.int8)
    w = rng.integers(-128, 128, (N, K_)).astype(np.int8)
    b = rng.integers(-2 ** 20, 2 ** 20, N).astype(np.int32)
    return a, w, b


def ref_acc(a, w, b):
    return a.astype(np.int32) @ w.astype(np.int32).T + b.astype(np.int32)[None, :]


@pytest.mark.parametrize("M,N,K_", GEMM_SHAPES)
def test_gemm_raw_i32(K, M, N, K_):
    rng = np.random.default_rng(M * 7 + N * 3 + K_)
    a, w, b = gemm_inputs(rng, M, N, K_)
    got = K.gemm_i8(dev(a), dev(w), bias=dev(b), mode="raw")
    assert_equal(got, ref_acc(a, w, b), "gemm raw %dx%dx%d" % (M, N, K_))


def test_gemm_matches_oracle_kat(K, kat):
    # reference QuantLinear known answer (K=48 is not a multiple of 128: TMA zero fill)
    a, wq, bq = kat["lin_a"], kat["lin_wq"], kat["lin_bq"]
    got = K.gemm_i8(dev(a.astype(np.int8)), dev(wq.astype(np.int8)), bias=dev(bq.astype(np.int32)), mode="raw")
    assert_equal(got, kat["lin_acc"], "QuantLinear KAT")
    sf = kat["lin_sf"].reshape(-1)
    got = K.gemm_i8(dev(a.astype(np.int8)), dev(wq.astype(np.int8)), bias=dev(bq.astype(np.int32)),
                    mode="carrier", scale=dev(sf))
    want = (kat["lin_acc"].astype(np.float32) * sf[None, :]).astype(np.float32)
    assert np.array_equal(got.cpu().numpy(), want)


@pytest.mark.parametrize("M,N,K_", [(197, 192, 192), (300, 768, 768), (256, 3072, 768), (130, 1000, 768),
                                    (1300, 2304, 768), (2048, 768, 256), (640, 1024, 128)])     # M >= 512: 2-CTA clusters (W multicast)
@pytest.mark.parametrize("e_lo,e_hi", [(33, 44), (20, 50)])
def test_gemm_requant_i8(K, M, N, K_, e_lo, e_hi):
    rng = np.random.default_rng(M + N + K_ + e_lo)
    a, w, b = gemm_inputs(rng, M, N, K_)
    m, e = rand_me(rng, N, e_lo, e_hi, neg_every=5)
    m[:3] = [2 ** 30, -2 ** 30, 2 ** 31 - 1]           # power-of-two multipliers produce exact ties
    acc = ref_acc(a, w, b)
    want = O.requant(acc, m, e, 8)
    got = K.gemm_i8(dev(a), dev(w), bias=dev(b), mode="requant", me=me_dev(K, m, e), bits=8)
    assert_equal(got, want, "gemm rq8 %dx%dx%d e[%d,%d]" % (M, N, K_, e_lo, e_hi))
    # with the tightest valid accumulator bound most channels take the tie-free form, the 2^30 multipliers stay flagged
    bits = int(np.abs(acc).max()).bit_length()
    got = K.gemm_i8(dev(a), dev(w), bias=dev(b), mode="requant", me=me_dev(K, m, e), bits=8, acc_bits=bits)
    assert_equal(got, want, "gemm rq8 acc_bits=%d %dx%dx%d" % (bits, M, N, K_))


@pytest.mark.parametrize("M,N,K_", [(197, 192, 768), (300, 768, 3072), (128, 256, 128), (1411, 768, 768), (1024, 1280, 384),
                                    # 128-wide un-paired tiles (N = 384: DeiT-small, Swin stage 3): several tiles per CTA,
                                    # an M tail, one n-tile, a long K, a short K
                                    (20000, 384, 384), (1411, 384, 1536), (130, 128, 256), (6272, 384, 96)])
@pytest.mark.parametrize("variant", ["plain", "residual", "two_stage"])
def test_gemm_requant_i16(K, M, N, K_, variant):
    rng = np.random.default_rng(M + N + K_ + len(variant))
    a, w, b = gemm_inputs(rng, M, N, K_)
    m, e = rand_me(rng, N, 33, 40, neg_every=4)
    acc = ref_acc(a, w, b)
    kw = {}
    if variant == "plain":
        want = O.requant(acc, m, e, 16)
    else:
        res = rng.integers(-32768, 32768, (M, N)).astype(np.int16)
        m1, e1 = rand_me(rng, 1, 30, 33)
        kw = dict(residual=dev(res), res_me=(m1[0], e1[0]))
        if variant == "residual":
            want = O.requant(acc, m, e, 16, res, m1, e1)
        else:
            m2, e2 = rand_me(rng, 1, 30, 33)
            if M == 128:
                m2[0], e2[0], m1[0], e1[0] = 2 ** 30, 31, -2 ** 30, 32      # power-of-two ratios: exact ties in both scalar stages
                kw["res_me"] = (m1[0], e1[0])
            q1 = O.requant(acc, m, e, 16)
            want = O.requant(q1, m2, e2, 16, res, m1, e1)
            kw.update(two_stage=True, me2=(m2[0], e2[0]))
    got = K.gemm_i8(dev(a), dev(w), bias=dev(b), mode="requant", me=me_dev(K, m, e), bits=16, **kw)
    assert_equal(got, want, "gemm rq16 %s %dx%dx%d" % (variant, M, N, K_))
    got = K.gemm_i8(dev(a), dev(w), bias=dev(b), mode="requant", me=me_dev(K, m, e), bits=16,
                    acc_bits=int(np.abs(acc).max()).bit_length(), **kw)
    assert_equal(got, want, "gemm rq16 %s acc_bits %dx%dx%d" % (variant, M, N, K_))


@pytest.mark.parametrize("e2,e1", [(18, 47), (24, 31), (31, 32), (32, 18), (40, 40), (47, 24), (48, 33), (33, 55)])
def test_gemm_two_stage_exponent_sweep(K, e2, e1):
    """Residual-block epilogue (attn.proj / mlp.fc2: two-stage requant + int16 residual) over the exponent range of
    the scalar dyadics: e <= 47 takes the fixed-shift form on packed int16 pairs, larger e the unified form; the first
    stage saturates at +-32768 on part of the tile and the residual includes the int16 extremes."""
    M, N, K_ = 700, 512, 256
    rng = np.random.default_rng(1000 * e2 + e1)
    a, w, b = gemm_inputs(rng, M, N, K_)
    m, e = rand_me(rng, N, 33, 38, neg_every=7)
    e[: N // 4] = 32                                      # large first-stage results: the 16-bit clamp is active
    acc = ref_acc(a, w, b)
    res = rng.integers(-32768, 32768, (M, N)).astype(np.int16)
    res[0, :8] = [-32768, 32767, -32768, 32767, 0, 1, -1, 12345]
    m2 = np.array([rng.integers(2 ** 30, 2 ** 31) | 1], np.int64)       # odd multipliers: no reachable ties
    m1 = np.array([-(rng.integers(2 ** 30, 2 ** 31) | 1)], np.int64)
    q1 = O.requant(acc, m, e, 16)
    assert (np.abs(q1) >= 32767).any()
    want = O.requant(q1, m2, np.array([e2]), 16, res, m1, np.array([e1]))
    got = K.gemm_i8(dev(a), dev(w), bias=dev(b), mode="requant", me=me_dev(K, m, e), bits=16, residual=dev(res),
                    res_me=(int(m1[0]), e1), two_stage=True, me2=(int(m2[0]), e2))
    assert_equal(got, want, "gemm two-stage e2=%d e1=%d" % (e2, e1))


def test_gemm_strided_a_and_out(K):
    # A as a column slice of a wider buffer (lda > K), out into a wider buffer
    rng = np.random.default_rng(77)
    M, N, K_ = 200, 256, 256
    big = rng.integers(-128, 128, (M, 512)).astype(np.int8)
    w = rng.integers(-128, 128, (N, K_)).astype(np.int8)
    A = dev(big)[:, 256:512]
    out = torch.zeros((M, 512), dtype=torch.int32, device="cuda")
    K.gemm_i8(A, dev(w), mode="raw", out=out[:, 128:128 + N])
    want = big[:, 256:].astype(np.int32) @ w.astype(np.int32).T
    assert_equal(out[:, 128:128 + N], want, "strided gemm")
    assert int(out[:, :128].abs().sum()) == 0 and int(out[:, 128 + N:].abs().sum()) == 0


# ------------------------------------------------------------------------------- batched matmul
def test_bmm_kat(K, kat):
    A, B = kat["mm_A"], kat["mm_B"]                    # [2,3,11,16] x [2,3,16,11]
    a = dev(A.astype(np.int8)).reshape(6, 11, 16)
    b = dev(B.astype(np.int8)).reshape(6, 16, 11)
    assert_equal(K.bmm_i32(a, b, trans_b=False), kat["mm_acc"].reshape(6, 11, 11), "bmm QK KAT")
    bt = b.transpose(1, 2).contiguous()
    assert_equal(K.bmm_i32(a, bt, trans_b=True), kat["mm_acc"].reshape(6, 11, 11), "bmm QK^T KAT")
    p = dev(kat["mm2_P"].astype(np.int16)).reshape(6, 11, 11)
    v = dev(kat["mm2_V"].astype(np.int8)).reshape(6, 11, 16)
    assert_equal(K.bmm_i32(p, v, trans_b=False), kat["mm2_acc"].reshape(6, 11, 16), "bmm PV KAT (int16 P)")


# ------------------------------------------------------------------------------- fused attention
def oracle_attention(qkv, n_seq, n_tok, H, D, me_s, x0, me_o, p_bits):
    Cc = H * D
    out = np.zeros((n_seq * n_tok, Cc), np.int64)
    for b in range(n_seq):
        blk = qkv[b * n_tok:(b + 1) * n_tok].astype(np.int64)
        for h in range(H):
            q = blk[:, h * D:(h + 1) * D]
            k = blk[:, Cc + h * D:Cc + (h + 1) * D]
            v = blk[:, 2 * Cc + h * D:2 * Cc + (h + 1) * D]
            s = O.requant(q @ k.T, [me_s[0]], [me_s[1]], 8)
            p = O.shiftmax(s, x0, p_bits)
            o = O.requant(p @ v, [me_o[0]], [me_o[1]], 8)
            out[b * n_tok:(b + 1) * n_tok, h * D:(h + 1) * D] = o
    return out


@pytest.mark.parametrize("n_seq,n_tok,H,D,p_bits", [(2, 197, 3, 64, 16), (3, 49, 3, 32, 8), (1, 197, 12, 64, 16),
                                                    (2, 64, 2, 64, 16), (2, 50, 4, 32, 16), (1, 17, 1, 64, 8),
                                                    (1, 129, 2, 64, 16), (1, 224, 1, 64, 16), (2, 5, 2, 64, 16),
                                                    (1, 200, 2, 64, 16), (3, 128, 1, 64, 16),
                                                    # more (image, head) items than resident CTAs (2 x 148): the persistent
                                                    # loop with next-item operand prefetch, one and two m-tiles per item
                                                    (110, 64, 3, 64, 16), (42, 197, 8, 64, 16)])
def test_attention(K, n_seq, n_tok, H, D, p_bits):
    rng = np.random.default_rng(n_tok * 31 + H)
    qkv = rng.integers(-128, 128, (n_seq * n_tok, 3 * H * D)).astype(np.int8)
    # make some rows strongly peaked so the softmax is not flat
    qkv[::7, :H * D] = np.clip(qkv[::7, :H * D].astype(np.int32) * 3, -128, 127).astype(np.int8)
    s_attn = np.float32(0.031)
    acc_scale = np.float32(127 * s_attn / (D * 127 * 40))       # scores use most of the int8 range
    m_s, e_s = K.dyadic_host(np.array([acc_scale], np.float32), s_attn)
    x0 = O.x0_of(s_attn)
    m_o, e_o = K.dyadic_host(np.array([2.0 ** -(p_bits - 1) * 0.02], np.float32), np.float32(0.02 * 1.3))
    me_s, me_o = (int(m_s[0]), int(e_s[0])), (int(m_o[0]), int(e_o[0]))
    want = oracle_attention(qkv, n_seq, n_tok, H, D, me_s, x0, me_o, p_bits)
    got = K.attention_i8(dev(qkv), n_seq, n_tok, H, D, me_s, x0, me_o, p_bits=p_bits)
    assert np.abs(want).max() > 12, "test should produce non-trivial outputs"
    assert_equal(got, want, "attention n_tok=%d H=%d D=%d P%d" % (n_tok, H, D, p_bits))


@pytest.mark.parametrize("s_attn", [0.3, 0.11, 0.0122, 0.0039, 0.001, 0.00031, 2.1e-5])
def test_attention_over_softmax_scales(K, s_attn):
    """DeiT attention over Shiftmax input scales 0.3 ... 2e-5 (x0 = -4 ... -47619): the pipelined tcgen05 kernel covers all of
    them (round 1's needed a scale above 1/256), with 64-bit row sums below ~5e-4.  Bit-exact against the oracle."""
    n_seq, n_tok, H, D = 3, 197, 2, 64
    rng = np.random.default_rng(int(s_attn * 1e7))
    qkv = rng.integers(-128, 128, (n_seq * n_tok, 3 * H * D)).astype(np.int8)
    qkv[::7, :H * D] = np.clip(qkv[::7, :H * D].astype(np.int32) * 3, -128, 127).astype(np.int8)
    s_attn = np.float32(s_attn)
    acc_scale = np.float32(127 * s_attn / (D * 127 * 40))
    m_s, e_s = K.dyadic_host(np.array([acc_scale], np.float32), s_attn)
    x0 = O.x0_of(s_attn)
    m_o, e_o = K.dyadic_host(np.array([2.0 ** -15 * 0.02], np.float32), np.float32(0.02 * 1.3))
    me_s, me_o = (int(m_s[0]), int(e_s[0])), (int(m_o[0]), int(e_o[0]))
    want = oracle_attention(qkv, n_seq, n_tok, H, D, me_s, x0, me_o, 16)
    got = K.attention_i8(dev(qkv), n_seq, n_tok, H, D, me_s, x0, me_o, p_bits=16)
    assert_equal(got, want, "attention s=%g x0=%d" % (float(s_attn), x0))


@pytest.mark.parametrize("n_win,H,with_mask", [(4, 3, True), (1, 6, False), (16, 2, True)])
def test_attention_swin_bias_and_mask(K, n_win, H, with_mask):
    """Window attention of Swin in the fused kernel (swin_quant.py:135-164): scores -> qact_attn1 -> qact2 with the
    relative-position bias as identity -> shifted-window mask -> 8-bit Shiftmax -> P V -> qact3, 49 tokens, head_dim 32."""
    n_tok, D = 49, 32
    n_seq = 2 * n_win
    rng = np.random.default_rng(100 * n_win + H)
    qkv = rng.integers(-128, 128, (n_seq * n_tok, 3 * H * D)).astype(np.int8)
    bias = rng.integers(-128, 128, (H, n_tok, n_tok)).astype(np.int8)
    s_a, s_b, s_2 = np.float32(0.05), np.float32(0.004), np.float32(0.043)
    acc_scale = np.float32(127 * s_a / (D * 127 * 30))
    m_s, e_s = K.dyadic_host(np.array([acc_scale], np.float32), s_a)
    m_2, e_2 = K.dyadic_host(np.array([s_a], np.float32), s_2)
    m_b, e_b = K.dyadic_host(np.array([s_b], np.float32), s_2)
    x0 = O.x0_of(s_2)
    m_o, e_o = K.dyadic_host(np.array([2.0 ** -7 * 0.02], np.float32), np.float32(0.02 * 0.9))
    mask01 = np.zeros((n_win, n_tok, n_tok), np.int64)
    if with_mask:
        grp = rng.integers(0, 3, (n_win, n_tok))
        mask01 = (grp[:, :, None] != grp[:, None, :]).astype(np.int64)      # block structure as in swin_quant.py:223-247
        mask01[0] = 0                                                       # the first window is never masked
    add = int(np.rint(np.float64(-100.0) / np.float64(s_2)))               # integer addend of a masked entry (App. A.5)
    Cc = H * D
    want = np.zeros((n_seq * n_tok, Cc), np.int64)
    for b in range(n_seq):
        blk = qkv[b * n_tok:(b + 1) * n_tok].astype(np.int64)
        for h in range(H):
            q, k, v = blk[:, h * D:(h + 1) * D], blk[:, Cc + h * D:Cc + (h + 1) * D], blk[:, 2 * Cc + h * D:2 * Cc + (h + 1) * D]
            s = O.requant(q @ k.T, [m_s[0]], [e_s[0]], 8)
            s = O.requant(s, [m_2[0]], [e_2[0]], 8, bias[h].astype(np.int64), [m_b[0]], [e_b[0]])
            s = s + add * mask01[b % n_win]
            pr = O.shiftmax(s, x0, 8)
            want[b * n_tok:(b + 1) * n_tok, h * D:(h + 1) * D] = O.requant(pr @ v, [m_o[0]], [e_o[0]], 8)
    mask_dev = dev((add * mask01).astype(np.int32)) if with_mask else None
    got = K.attention_i8(dev(qkv), n_seq, n_tok, H, D, (int(m_s[0]), int(e_s[0])), x0, (int(m_o[0]), int(e_o[0])), p_bits=8,
                         relbias=dev(bias), me_s2=(int(m_2[0]), int(e_2[0])), me_b=(int(m_b[0]), int(e_b[0])),
                         mask=mask_dev, n_win=n_win if with_mask else 0)
    assert np.abs(want).max() > 8, "test should produce non-trivial outputs"
    assert_equal(got, want, "swin attention n_win=%d H=%d mask=%s" % (n_win, H, with_mask))


# ------------------------------------------------------------------------------- hot-path specialisations
@pytest.mark.parametrize("cols,s", [(768, 0.03), (3072, 0.045), (1536, 0.0734), (384, 0.012), (4096, 0.02), (48, 0.05),
                                    (768, 0.15), (3072, 0.6), (384, 0.0011)])
def test_shiftgelu_lut_matches_general_and_oracle(K, cols, s):
    rng = np.random.default_rng(cols)
    q = rng.integers(-128, 128, (77, cols)).astype(np.int64)
    q[0] = rng.integers(-128, -1, cols)               # all-negative row
    q[1] = 127
    q[2] = -128
    x0 = O.x0_of(O.gelu_sig_scale(np.float32(s)))
    s_in = np.float32(s) * np.float32(1 / 128)
    m, e = K.dyadic_host(np.array([s_in], np.float32), np.float32(0.9 * 127 * 127 * float(s_in) / 127.0))
    want = O.requant(O.shiftgelu(q, x0), m, e, 8)
    me = me_dev(K, m, e)
    lut = K.shiftgelu_build_lut(x0, me)
    got = K.shiftgelu_lut(dev(q.astype(np.int8)), lut)
    assert_equal(got, want, "shiftgelu LUT cols=%d" % cols)
    gen = K.shiftgelu(dev(q.astype(np.int8)), x0, me, 8)
    assert torch.equal(got, gen)


@pytest.mark.parametrize("C,mag", [(768, 9000), (192, 20000), (384, 32767), (96, 300), (1024, 5000), (8, 100)])
def test_layernorm_i16_i8_fast(K, C, mag):
    rng = np.random.default_rng(C)
    q = rng.integers(-mag, mag + 1, (203, C)).astype(np.int64)
    q[1] = 7                                          # zero variance
    q[2] = 0
    q[2, 0] = mag                                     # one-hot
    half = C // 2
    q[3, 0] += half - (q[3].sum() % C) if abs(q[3, 0]) < 30000 - C else 0      # exact .5 mean tie when it fits
    bq = rng.integers(-2 ** 24, 2 ** 24, C).astype(np.int64)
    z = O.layernorm(q, bq)
    m, e = rand_me(rng, C, 40, 52, neg_every=5)
    m[:2] = [2 ** 30, -2 ** 30]
    e[3] = 20                                         # general (e < 32) requant inside the fast kernel
    want = O.requant(z, m, e, 8)
    got = K.layernorm_i16_i8(dev(q.astype(np.int16)), dev(bq.astype(np.int32)), me_dev(K, m, e))
    assert_equal(got, want, "layernorm_i16_i8 C=%d" % C)


@pytest.mark.parametrize("C,rows,mag", [(256, 1, 3000), (256, 7, 32767), (512, 3, 12000), (512, 4, 500), (768, 5, 32767),
                                        (768, 9, 40), (1024, 2, 32767), (768, 1031, 9000),
                                        # 16 and 8 lanes per row group (two rows / one row per lane)
                                        (384, 1, 32767), (384, 6, 9000), (384, 1027, 20000), (192, 3, 32767), (192, 1030, 700),
                                        (128, 5, 32767), (64, 9, 1000), (64, 2, 32767)])
def test_layernorm_i16_i8_four_rows_per_warp(K, C, rows, mag):
    """The four-rows-per-warp kernel (32 / 16 / 8 lanes per row group): row counts around the group of four, extreme magnitudes
    (|x - mu| close to the integer square root of the variance: the 32-bit y * F product is at its bound), constant rows."""
    rng = np.random.default_rng(C + rows)
    q = rng.integers(-mag, mag + 1, (rows, C)).astype(np.int64)
    q[0, :] = -mag
    q[0, C // 3] = mag                                 # one element carries almost all of the variance
    if rows > 1:
        q[1] = mag                                     # zero variance at the top of the range
    if rows > 2:
        q[2, ::2] = mag
        q[2, 1::2] = -mag                              # maximal variance
    bq = rng.integers(-2 ** 24, 2 ** 24, C).astype(np.int64)
    z = O.layernorm(q, bq)
    m, e = rand_me(rng, C, 40, 52, neg_every=3)
    want = O.requant(z, m, e, 8)
    got = K.layernorm_i16_i8(dev(q.astype(np.int16)), dev(bq.astype(np.int32)), me_dev(K, m, e))
    assert_equal(got, want, "layernorm_i16_i8 (four rows per warp) C=%d rows=%d" % (C, rows))


def test_quantize_patchify_fused(K):
    rng = np.random.default_rng(21)
    for (B, Cin, H, W, p) in [(2, 3, 32, 48, 16), (3, 3, 16, 16, 4), (1, 3, 224, 224, 16)]:
        x = (rng.standard_normal((B, Cin, H, W)) * 2).astype(np.float32)
        s = np.float32(0.0191)
        q = O.quantize_f32(x, s, 8).astype(np.int8)
        want = q.reshape(B, Cin, H // p, p, W // p, p).transpose(0, 2, 4, 1, 3, 5).reshape(-1, Cin * p * p)
        got = K.quantize_patchify(dev(x), dev(np.array([s])), p).cpu().numpy()
        assert np.array_equal(got, want), (B, Cin, H, W, p)


def test_quantize_patchify_u8_matches_reference_transform(K):
    """uint8 pixels -> ToTensor -> Normalize (utils/data_utils.py:90-91) -> input QuantAct -> unfold, against the same
    chain evaluated by torch in fp32 on the CPU (every step is a single correctly-rounded fp32 operation)."""
    import torch
    rng = np.random.default_rng(31)
    mean = torch.tensor([0.485, 0.456, 0.406]); std = torch.tensor([0.229, 0.224, 0.225])
    for (B, H, W) in [(2, 32, 48), (1, 224, 224)]:
        u = rng.integers(0, 256, (B, 3, H, W)).astype(np.uint8)
        u[0, :, 0, :16] = np.arange(16, dtype=np.uint8) * 17           # includes 0 and 255
        s = np.float32(0.02071)
        t = torch.from_numpy(u).to(torch.float32).div(255)             # ToTensor
        x = t.sub(mean[None, :, None, None]).div(std[None, :, None, None])   # Normalize
        q = O.quantize_f32(x.numpy(), s, 8).astype(np.int8)
        want = q.reshape(B, 3, H // 16, 16, W // 16, 16).transpose(0, 2, 4, 1, 3, 5).reshape(-1, 3 * 256)
        got = K.quantize_patchify_u8(dev(u), mean.cuda(), std.cuda(), dev(np.array([s])), 16).cpu().numpy()
        assert np.array_equal(got, want), (B, H, W)


@pytest.mark.parametrize("B,N,C", [(3, 17, 64), (2, 197, 768), (5, 50, 192)])
def test_embed_tokens_fast_matches_general(K, B, N, C):
    rng = np.random.default_rng(22 + C)
    pe = rng.integers(-32768, 32768, (B * (N - 1), C)).astype(np.int16)
    cls = rng.integers(-200000, 200000, C).astype(np.int32)        # the cls token is not clamped to 16 bits
    pos = rng.integers(-32768, 32768, (N, C)).astype(np.int16)
    for (m, e), (m1, e1) in [((1518500250, 31), (1234567891, 33)), ((2 ** 30, 30), (-2 ** 30, 32)), ((1900000001, 40), (1100000000, 17))]:
        x = np.concatenate([np.broadcast_to(cls.astype(np.int64), (B, 1, C)), pe.reshape(B, N - 1, C).astype(np.int64)], axis=1)
        want = O.requant(x.reshape(B * N, C), [m], [e], 16, pos.astype(np.int64), [m1], [e1])
        a = K.embed_tokens(dev(pe), dev(cls), dev(pos), B, N, C, (m, e), (m1, e1), 16)
        b = K.embed_tokens_fast(dev(pe), dev(cls), dev(pos), B, N, C, (m, e), (m1, e1))
        assert_equal(a, want, "embed_tokens")
        assert_equal(b, want, "embed_tokens_fast")


def _four_squares(n, rng):
    """n = a^2 + b^2 + c^2 + d^2 by randomised greedy search (n < 2^40)."""
    import math
    for c in range(6):                              # n - c^2 - d^2 = x^2 or 2 x^2 (powers of two, squares, their neighbours)
        for d in range(6):
            r = n - c * c - d * d
            if r < 0:
                continue
            x = math.isqrt(r)
            if x * x == r:
                return x, 0, c, d
            if r % 2 == 0 and math.isqrt(r // 2) ** 2 == r // 2:
                return math.isqrt(r // 2), math.isqrt(r // 2), c, d
    for _ in range(20000):
        a = math.isqrt(n) - int(rng.integers(0, 6))
        r1 = n - a * a
        if r1 < 0:
            continue
        b = math.isqrt(r1) - int(rng.integers(0, 12))
        r2 = r1 - b * b
        if b < 0 or r2 < 0:
            continue
        c = math.isqrt(r2) - int(rng.integers(0, 12))
        r3 = r2 - c * c
        if c < 0 or r3 < 0:
            continue
        d = math.isqrt(r3)
        if d * d == r3:
            return a, b, c, d
    return None


def test_layernorm_isqrt_edge_variances(K):
    """The 10-step integer sqrt (quant_modules.py:366-370) has a closed form for 2^22 <= V < 2^40 except when
    V = (s+1)^2 - 1 (the iteration then alternates s, s+1); rows [+-a, +-b, +-c, +-d] give V = 2(a^2+b^2+c^2+d^2)
    exactly (mean 0), so every even V is reachable: alternating cases, range boundaries, tiny and huge variances."""
    import math
    rng = np.random.default_rng(77)
    targets = []
    for s in [2048, 2050, 2052, 4096, 4100, 65530, 65534, 65536, 65538, 65540, 100000, 300000, 724000, 1048570, 1048574]:     # even s: V = s(s+2) is even
        targets += [s * (s + 2), s * s, s * s + 2, s * (s + 2) - 2]
    targets += [2 ** 22, 2 ** 22 - 2, 2 ** 22 + 2, 2 ** 32, 2 ** 32 - 2, 2 ** 32 + 2, 2 ** 40 - 2, 2 ** 40, 2 ** 40 + 2, 2, 8, 200, 2 ** 41 + 6]
    targets += [int(v) * 2 for v in rng.integers(1, 2 ** 39, 60)]
    rows = []
    for V in targets:
        fs = _four_squares(V // 2, rng)
        if fs is None:
            continue
        a, b, c, d = fs
        rows.append([a, -a, b, -b, c, -c, d, -d])
    assert len(rows) > 80
    q = np.array(rows, dtype=np.int64)
    assert (np.abs(q) < 2 ** 31).all()
    bq = rng.integers(-1000, 1000, 8).astype(np.int64)
    want = O.layernorm(q, bq)
    got = K.layernorm(dev(q.astype(np.int32)), dev(bq.astype(np.int32)))
    assert_equal(got, want, "layernorm isqrt edge cases")


# ------------------------------------------------------------------------------- Swin hot path (round 2)
def pack_mask_bits(mask01):
    """[n_win, N, N] 0/1 -> uint64 [n_win, N]: bit j of (w, i) = key j masked for query i (host side of SwinEngine)."""
    n_win, N, _ = mask01.shape
    w = (mask01.astype(np.uint64) << np.arange(N, dtype=np.uint64)[None, None, :]).sum(axis=2, dtype=np.uint64)
    return np.ascontiguousarray(w)


@pytest.mark.parametrize("n_seq,n_win,H,with_mask,s_a,s_2", [
    (8, 4, 3, True, 0.05, 0.043), (2, 1, 6, False, 0.05, 0.043), (32, 16, 2, True, 0.05, 0.043),
    (5, 1, 1, False, 0.021, 0.017),            # odd number of windows (last pair half empty), one head (pair half empty)
    (6, 2, 4, True, 0.011, 0.0041),            # e_2 < 32 (s_a / s_2 > 1), fine softmax scale
    (1300, 4, 3, True, 0.03, 0.2),             # more work units than resident CTAs; coarse softmax scale (x0 = -5)
    (64, 64, 24, True, 0.05, 0.043)])          # 24 heads (Swin stage 4), 64 windows per image (stage 1)
def test_window_attention_tc(K, n_seq, n_win, H, with_mask, s_a, s_2):
    """tcgen05 window attention (two 49-token windows per MMA tile, head pairs per TMA box) against the oracle and,
    bit for bit, against the general mma.sync kernel.  swin_quant.py:121-164."""
    n_tok, D = 49, 32
    rng = np.random.default_rng(100 * n_win + H + n_seq)
    qkv = rng.integers(-128, 128, (n_seq * n_tok, 3 * H * D)).astype(np.int8)
    qkv[::5, :H * D] = np.clip(qkv[::5, :H * D].astype(np.int32) * 3, -128, 127).astype(np.int8)
    bias = rng.integers(-128, 128, (H, n_tok, n_tok)).astype(np.int8)
    s_a, s_b, s_2 = np.float32(s_a), np.float32(0.004), np.float32(s_2)
    acc_scale = np.float32(127 * s_a / (D * 127 * 30))
    m_s, e_s = K.dyadic_host(np.array([acc_scale], np.float32), s_a)
    m_2, e_2 = K.dyadic_host(np.array([s_a], np.float32), s_2)
    m_b, e_b = K.dyadic_host(np.array([s_b], np.float32), s_2)
    x0 = O.x0_of(s_2)
    m_o, e_o = K.dyadic_host(np.array([2.0 ** -7 * 0.02], np.float32), np.float32(0.02 * 0.9))
    mask01 = np.zeros((n_win, n_tok, n_tok), np.int64)
    if with_mask:
        grp = rng.integers(0, 3, (n_win, n_tok))
        mask01 = (grp[:, :, None] != grp[:, None, :]).astype(np.int64)
        mask01[0] = 0
    add = int(np.rint(np.float64(-100.0) / np.float64(s_2)))
    me_s, me_2, me_b, me_o = [(int(a[0]), int(b[0])) for a, b in ((m_s, e_s), (m_2, e_2), (m_b, e_b), (m_o, e_o))]
    bias_rq = K.requant(dev(bias.reshape(-1, 1).astype(np.int32)), me_dev(K, m_b, e_b), 16).reshape(H, n_tok, n_tok).contiguous()
    assert_equal(bias_rq, O.requant(bias.astype(np.int64), [m_b[0]], [e_b[0]], 16), "bias requant")
    got = K.window_attention_i8(dev(qkv), n_seq, H, me_s, me_2, x0, me_o, bias_rq,
                                mask_bits=dev(pack_mask_bits(mask01).view(np.int64)) if with_mask else None,
                                n_win_img=n_win if with_mask else 0, mask_add=add if with_mask else 0)
    gen = K.attention_i8(dev(qkv), n_seq, n_tok, H, D, me_s, x0, me_o, p_bits=8, relbias=dev(bias), me_s2=me_2, me_b=me_b,
                         mask=dev((add * mask01).astype(np.int32)) if with_mask else None, n_win=n_win if with_mask else 0)
    assert torch.equal(got, gen), "tcgen05 window attention differs from the general kernel: %d elements" % int((got != gen).sum())
    if n_seq <= 64:                                   # the numpy oracle loop is slow
        Cc = H * D
        nb = min(n_seq, 8)
        want = np.zeros((nb * n_tok, Cc), np.int64)
        for b in range(nb):
            blk = qkv[b * n_tok:(b + 1) * n_tok].astype(np.int64)
            for h in range(H):
                q, k, v = blk[:, h * D:(h + 1) * D], blk[:, Cc + h * D:Cc + (h + 1) * D], blk[:, 2 * Cc + h * D:2 * Cc + (h + 1) * D]
                s = O.requant(q @ k.T, [m_s[0]], [e_s[0]], 8)
                s = O.requant(s, [m_2[0]], [e_2[0]], 8, bias[h].astype(np.int64), [m_b[0]], [e_b[0]])
                s = s + add * mask01[b % n_win]
                want[b * n_tok:(b + 1) * n_tok, h * D:(h + 1) * D] = O.requant(O.shiftmax(s, x0, 8) @ v, [m_o[0]], [e_o[0]], 8)
        assert np.abs(want).max() > 8
        assert_equal(got[:nb * n_tok], want, "window attention n_win=%d H=%d mask=%s" % (n_win, H, with_mask))


def test_window_attention_refuses_scales_outside_its_domain(K):
    from ivit_b200._lib import IvitError
    qkv = torch.zeros((2 * 49, 96), dtype=torch.int8, device="cuda")
    b = torch.zeros((1, 49, 49), dtype=torch.int16, device="cuda")
    with pytest.raises(IvitError, match="fast-form"):
        K.window_attention_i8(qkv, 2, 1, (2 ** 30, 20), (2 ** 30, 31), -20, (2 ** 30, 40), b)      # e_s < 32
    with pytest.raises(IvitError, match="fast-form"):
        K.window_attention_i8(qkv, 2, 1, (2 ** 30 + 1, 40), (2 ** 30, 31), -20, (2 ** 30 + 1, 40), b)   # qact2: reachable tie


@pytest.mark.parametrize("C,G,L_out,mag", [(96, 1, 56, 9000), (192, 1, 28, 20000), (384, 1, 49, 32767), (768, 1, 49, 5000),
                                           (384, 4, 16, 9000), (768, 4, 49, 30000), (1536, 4, 9, 32767), (48, 1, 5, 300),
                                           (1024, 4, 4, 700), (2048, 4, 7, 12000), (1792, 1, 3, 900)])
def test_layernorm_gather(K, C, G, L_out, mag):
    """IntLayerNorm + QuantAct over rows gathered through a per-image map: window permutation (G = 1, with the permuted
    int16 copy) and 2 x 2 patch merging (G = 4)."""
    rng = np.random.default_rng(C + G)
    B = 5
    L_in = L_out * G
    Cs = C // G
    x = rng.integers(-mag, mag + 1, (B * L_in, Cs)).astype(np.int64)
    x[1] = 7
    x[2] = 0
    x[2, 0] = mag
    if G == 1:
        rowmap = rng.permutation(L_out).astype(np.int32)
        gathered = x.reshape(B, L_in, Cs)[:, rowmap].reshape(B * L_out, C)
    else:
        rowmap = rng.permutation(L_in).astype(np.int32).reshape(L_out, 4)
        gathered = x.reshape(B, L_in, Cs)[:, rowmap.reshape(-1)].reshape(B * L_out, C)
    bq = rng.integers(-2 ** 24, 2 ** 24, C).astype(np.int64)
    m, e = rand_me(rng, C, 40, 52, neg_every=5)
    want = O.requant(O.layernorm(gathered, bq), m, e, 8)
    xcopy = torch.zeros((B * L_out, C), dtype=torch.int16, device="cuda") if G == 1 else None
    got = K.layernorm_gather(dev(x.astype(np.int16)), B * L_out, C, G, dev(rowmap.reshape(-1)), L_out, L_in,
                             dev(bq.astype(np.int32)), me_dev(K, m, e), xcopy=xcopy)
    assert_equal(got, want, "layernorm_gather C=%d G=%d" % (C, G))
    if G == 1:
        assert_equal(xcopy, gathered, "gathered residual copy")
        ident = K.layernorm_gather(dev(x.astype(np.int16)), B * L_out, C, 1, None, L_out, L_in, dev(bq.astype(np.int32)), me_dev(K, m, e))
        assert_equal(ident, O.requant(O.layernorm(x, bq), m, e, 8), "identity map")


def test_avgpool_requant(K):
    rng = np.random.default_rng(77)
    for B, L, C in [(3, 49, 768), (2, 49, 1024), (1, 4, 8), (5, 196, 96)]:
        x = rng.integers(-128, 128, (B, L, C)).astype(np.int64)
        x[0, :, 0] = 1                                   # mean exactly 1
        x[0, :, 1] = np.arange(L) % 2                    # a .5-ish mean
        m, e = K.dyadic_host(np.array([0.031], np.float32), np.float32(0.027))
        want = O.requant(O.avgpool_rne(x), [m[0]], [e[0]], 8)
        got = K.avgpool_requant_i8(dev(x.astype(np.int8)), B, L, C, (int(m[0]), int(e[0])))
        assert_equal(got, want, "avgpool B=%d L=%d C=%d" % (B, L, C))


@pytest.mark.parametrize("C,rows", [(96, 1000), (128, 77), (192, 64), (40, 33)])
def test_layernorm_i8_i16x2(K, C, rows):
    """Swin patch embedding tail: IntLayerNorm over int8 rows + patch_embed.qact (per channel, 16 bit) + qact1 (scalar, 16 bit)."""
    rng = np.random.default_rng(C)
    q = rng.integers(-128, 128, (rows, C)).astype(np.int64)
    q[1] = 7
    q[2] = 0
    q[2, 0] = 127
    bq = rng.integers(-2 ** 24, 2 ** 24, C).astype(np.int64)
    m, e = rand_me(rng, C, 44, 50, neg_every=5)
    m2, e2 = K.dyadic_host(np.array([0.00031], np.float32), np.float32(0.00047))
    if C == 128:                                      # equal ranges of the two QuantActs (the synthetic Swin): identity dyadic
        m2, e2 = np.array([1 << 30]), np.array([30])
    want = O.requant(O.requant(O.layernorm(q, bq), m, e, 16), [m2[0]], [e2[0]], 16)
    got = K.layernorm_i8_i16x2(dev(q.astype(np.int8)), dev(bq.astype(np.int32)), me_dev(K, m, e), (int(m2[0]), int(e2[0])))
    assert np.abs(want).max() > 1000
    assert_equal(got, want, "layernorm_i8_i16x2 C=%d" % C)


def test_widen_i8_i16(K):
    rng = np.random.default_rng(3)
    x = rng.integers(-128, 128, (37, 48)).astype(np.int8)
    assert_equal(K.widen_i8_i16(dev(x)), x.astype(np.int64), "widen")
