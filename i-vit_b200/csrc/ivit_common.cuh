// Shared device helpers for the I-ViT sm_100a kernels: exact dyadic requantisation,
// shift-exponential, saturating packs, warp reductions, vector load/store.
//
// Integer semantics follow SURVEY.md Appendix A (restating the reference's
// models/quantization_utils/quant_utils.py and quant_modules.py, cited per function).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ivit_b200.h"

#define IVIT_DEVINL __device__ __forceinline__

namespace ivit {

// ------------------------------------------------------------------------------------
// Dyadic requant   out = RNE(z * m / 2^e)      fixedpoint_mul.forward, quant_utils.py:229-230
//   m : int32 (|m| in [2^30, 2^31), sign = sign of the scale ratio; 2^31 normalised away)
//   e : clamped to [-1, 63] by the host/dyadic kernel (e >= 63 => 0, e <= -1 saturates)
// Exact 64-bit product; result saturated to int32 (callers clamp further to 8/16 bits).
// ------------------------------------------------------------------------------------
IVIT_DEVINL int32_t sat_i64_to_i32(long long v) {
    v = v > 2147483647LL ? 2147483647LL : v;
    v = v < -2147483648LL ? -2147483648LL : v;
    return (int32_t)v;
}

IVIT_DEVINL long long requant64(long long z, int32_t m, int32_t e) {
    // general form used by the row kernels (z may exceed 32 bits after LayerNorm)
    if (e >= 63) return 0;
    // |z| < 2^32 guaranteed by callers => |z*m| < 2^63
    long long p = z * (long long)m;
    if (e <= 0) {
        // e == 0 : p ; e == -1 : 2p  (saturating)
        if (e < 0) {
            if (p > (1LL << 61)) return (1LL << 62);
            if (p < -(1LL << 61)) return -(1LL << 62);
            p *= 2;
        }
        return p;
    }
    const long long half = 1LL << (e - 1);
    const long long t = p + half;                    // |p| <= 2^62, half <= 2^61: no overflow
    long long q = t >> e;                            // floor((p + half) / 2^e)  == round half up
    const bool tie = (t & ((1ULL << e) - 1ULL)) == 0ULL;
    q -= (long long)(tie & (q & 1LL));               // exact .5 -> even
    return q;
}

IVIT_DEVINL int32_t requant32(int32_t z, int32_t m, int32_t e) {
    return sat_i64_to_i32(requant64((long long)z, m, e));
}

// Fast path for 32 <= e <= 62 when the caller knows ties are impossible or handled:
// p = z*m + half, q = hi32(p) >> (e-32).  Used by the GEMM epilogue (see ivit_gemm.cu).
IVIT_DEVINL int32_t requant32_e32(int32_t z, int32_t m, int32_t e) {
    const long long half = 1LL << (e - 1);
    const long long t = (long long)z * (long long)m + half;
    const int32_t hi = (int32_t)(t >> 32);
    const uint32_t lo = (uint32_t)t;
    const int sh = e - 32;
    int32_t q = hi >> sh;
    const bool tie = (lo == 0u) && ((hi & ((1 << sh) - 1)) == 0);
    q -= (int32_t)(tie & (q & 1));
    return q;
}

// A scalar (kernel-uniform) dyadic applied to operands with |z| < 2^zbits.
struct UniRq {
    int32_t m, e;
    long long half;
    int fast;                         // 16 <= e <= 62 and exact ties unreachable -> branch-free form
};
IVIT_DEVINL UniRq make_unirq(ivit_dyadic_t d, int zbits) {
    UniRq u;
    u.m = d.m; u.e = d.e;
    u.half = (d.e >= 1 && d.e <= 62) ? (1LL << (d.e - 1)) : 0;
    const int tz = __ffs(d.m) - 1;
    // no reachable tie: either the tie bit lies above every operand bit (e-1-tz > zbits), or the product has no
    // fractional bits at all (tz >= e: e.g. the identity ratio m = 2^30, e = 30 of two QuantActs with equal ranges)
    u.fast = (d.e >= 16 && d.e <= 62 && d.m != 0 && ((d.e - 1 - tz > zbits) || tz >= d.e)) ? 1 : 0;
    return u;
}
// out-of-line general form: keeps the (rarely taken) slow path from being inlined at every call site
// (code size matters: the attention kernel was instruction-fetch bound)
static __device__ __noinline__ int32_t requant32_general(int32_t z, int32_t m, int32_t e) {
    return sat_i64_to_i32(requant64((long long)z, m, e));
}
IVIT_DEVINL int32_t unirq_apply(const UniRq& u, int32_t z) {
    if (u.fast) {                                                  // uniform branch
        const long long t = (long long)z * (long long)u.m + u.half;
        return (u.e >= 32) ? ((int32_t)(t >> 32) >> (u.e - 32)) : (int32_t)(t >> u.e);
    }
    return requant32_general(z, u.m, u.e);
}

template <int BITS>
IVIT_DEVINL int32_t clamp_bits(int32_t v) {
    constexpr int32_t hi = (BITS >= 32) ? 2147483647 : ((1 << (BITS - 1)) - 1);
    constexpr int32_t lo = -hi - 1;
    return v < lo ? lo : (v > hi ? hi : v);
}
IVIT_DEVINL int32_t clamp_bits_rt(int32_t v, int bits) {
    if (bits >= 32) return v;
    const int32_t hi = (1 << (bits - 1)) - 1;
    const int32_t lo = -hi - 1;
    return v < lo ? lo : (v > hi ? hi : v);
}
IVIT_DEVINL int32_t clamp_i64_bits(long long v, int bits) {
    const long long hi = (bits >= 32) ? 2147483647LL : ((1LL << (bits - 1)) - 1);
    const long long lo = -hi - 1;
    return (int32_t)(v < lo ? lo : (v > hi ? hi : v));
}

// ------------------------------------------------------------------------------------
// int_exp_shift   quant_modules.py:410-423 (IntGELU) / :469-481 (IntSoftmax)
//   t = d + floor(d/2) - floor(d/16); t = max(t, n*x0); k = floor(t/x0); r = t - x0*k
//   E = max(floor((r/2 - x0) * 2^(n-k)), 0) = ((r - 2*x0) << (n-k)) >> 1
// x0 < 0.  inv_x0 = 1.0f / x0 (host supplied) seeds the floor division; two exact corrections make it an exact
// floor: |k| <= 184 (t in [n*x0, 184]), so the fp32 quotient is within 2e-5 of the true one and its floor is off by
// at most one.
// Domain (checked by the callers): 1 <= -x0 <= 2^24, -256 <= d <= 256.  The result is exact below 2^62 and SATURATES at
// 2^62 above (k < 0 with a tiny |x0|: ShiftGELU's e^(-x_max) of an all-negative row at a coarse input scale reaches
// 2^(n+184)).  Saturation is invisible to both users: Shiftmax only has d <= 0 (k >= 0, E < 2^40), and ShiftGELU feeds
// that term into a sum that is clamped at 2^31 - 1 (quant_modules.py:435-437) -- any value >= 2^31 acts the same.
// ------------------------------------------------------------------------------------
IVIT_DEVINL long long shiftexp(int32_t d, int32_t x0, float inv_x0, int n) {
    int32_t t = d + (d >> 1) - (d >> 4);
    const int32_t lim = n * x0;
    t = t < lim ? lim : t;
    int32_t k = __float2int_rd(__int2float_rn(t) * inv_x0);
    int32_t r = t - x0 * k;                          // want x0 < r <= 0
    if (r > 0) { k -= 1; r += x0; }
    if (r <= x0) { k += 1; r -= x0; }
    const int32_t base = r - 2 * x0;                 // in (|x0|, 2|x0|] : 2 <= base <= 2^25
    const int sh = n - k - 1;
    if (sh > 36) return 1LL << 62;                   // >= 2^38: saturated (see above)
    return (sh >= 0) ? ((long long)base << sh) : (long long)(base >> 1);
}

// IntLayerNorm's integer square root: k = 2^16; 10x k = floor((k + floor(V/k)) / 2)   quant_modules.py:366-370
// (exactly 10 steps, no early exit).  For V < 2^52 the 64-bit division is done in fp64: V and k are exact
// doubles, IEEE division is correctly rounded, and a non-integer quotient V/k is at least 1/k >= 2^-26 away
// from the next integer while the rounding error is below 2^-52 * 2^36, so truncation gives the exact floor.
//
// Early exit (exact): with s = floor(sqrt(V)), the iterate strictly decreases while k > s and never drops below s,
// so after the first step "k_next >= k" happens exactly when k == s; from there the sequence is either constant
// (k_next == s) or alternates s+1, s, s+1, ... (only when V == (s+1)^2 - 1).  The value after the remaining steps
// follows from their parity.
//
// Closed form (exact) for 2^22 <= V < 2^40: with x = k / sqrt(V), one step gives x' <= (x + 1/x) / 2 (the floors only
// lower k, never below s).  From k0 = 2^16: x1 <= 16.1 in both directions (sqrt(V) >= 2^11 resp. <= 2^20), then
// 8.04, 4.08, 2.16, 1.31, 1.037, 1.00066, 1 + 2.2e-7 -- i.e. k - sqrt(V) < 2^20 * 2.2e-7 < 1 after at most 8 steps, so
// k has reached s by step 8 <= 10 and the result is s, unless V == (s+1)^2 - 1 (alternating case: simulated).
IVIT_DEVINL unsigned long long ln_isqrt10_loop(unsigned long long V) {
    unsigned long long k = 65536ULL;
    const bool small = V < (1ULL << 52);
    const double Vd = (double)V;
#pragma unroll 1
    for (int it = 0; it < 10; ++it) {
        const unsigned long long q = small ? (unsigned long long)(Vd / (double)k) : (V / k);
        const unsigned long long kn = (k + q) >> 1;
        if (it > 0 && kn >= k) {                     // k == floor(sqrt(V)) reached: kn is k or k + 1
            const int rem = 9 - it;                  // steps still to take after this one
            return (kn == k || (rem & 1)) ? k : kn;
        }
        k = kn;
    }
    return k;
}
IVIT_DEVINL unsigned long long ln_isqrt10(unsigned long long V) {
    if (V >= (1ULL << 22) && V < (1ULL << 40)) {
        // s = floor(sqrt(V)): fp32 estimate (relative error < 2^-22, s < 2^20 -> off by at most 1) + exact fix-up
        uint32_t s = (uint32_t)__fsqrt_rn(__ull2float_rn(V));
        while ((unsigned long long)s * s > V) --s;
        while ((unsigned long long)(s + 1) * (s + 1) <= V) ++s;
        if ((unsigned long long)(s + 1) * (s + 1) - 1 != V) return (unsigned long long)s;
    }
    return ln_isqrt10_loop(V);
}

// floor((2^31-1) / S) for 1 <= S <= 2^31-1     quant_modules.py:438, 492
IVIT_DEVINL uint32_t recip_factor(uint32_t S) { return 2147483647u / S; }

// ------------------------------------------------------------------------------------
// Warp reductions
// ------------------------------------------------------------------------------------
IVIT_DEVINL int32_t warp_max_i32(int32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { int32_t t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
    return v;
}
IVIT_DEVINL long long warp_sum_i64(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
IVIT_DEVINL int32_t warp_sum_i32(int32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------
// Typed element access (op kernels are templated on the integer storage type)
// ------------------------------------------------------------------------------------
template <typename T> struct DType;
template <> struct DType<int8_t>  { static constexpr int id = IVIT_I8;  static constexpr int bits = 8; };
template <> struct DType<int16_t> { static constexpr int id = IVIT_I16; static constexpr int bits = 16; };
template <> struct DType<int32_t> { static constexpr int id = IVIT_I32; static constexpr int bits = 32; };

IVIT_DEVINL int32_t load_int(const void* p, int dtype, long long i) {
    switch (dtype) {
        case IVIT_I8:  return (int32_t)((const int8_t*)p)[i];
        case IVIT_U8:  return (int32_t)((const uint8_t*)p)[i];
        case IVIT_I16: return (int32_t)((const int16_t*)p)[i];
        default:       return ((const int32_t*)p)[i];
    }
}
IVIT_DEVINL void store_int(void* p, int dtype, long long i, int32_t v) {
    switch (dtype) {
        case IVIT_I8:  ((int8_t*)p)[i] = (int8_t)v; break;
        case IVIT_U8:  ((uint8_t*)p)[i] = (uint8_t)v; break;
        case IVIT_I16: ((int16_t*)p)[i] = (int16_t)v; break;
        default:       ((int32_t*)p)[i] = v; break;
    }
}

}  // namespace ivit
