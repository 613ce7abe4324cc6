// Bandwidth-oriented specialisations of the row operators for the fused engine (sm_100a).
//
// The general kernels of ivit_ops.cu evaluate the reference formulas element by element
// (~100 integer instructions per element for ShiftGELU).  On the hot path the operands are
// int8, so the expensive functions have tiny domains and become exact lookup tables:
//
//   ShiftGELU + mlp.qact1 (quant_modules.py:410-445, layers_quant.py:147-148):
//       out = clamp8(RNE(q * sigma(q, rowmax) * m / 2^e))   depends only on (q, rowmax) in int8 x int8
//       -> 64 KB table per layer built ONCE on the device with the general formula
//          (ivit_shiftgelu_build_lut), applied with one shared-memory byte lookup per element
//          (ivit_shiftgelu_lut): the kernel is HBM-bound (1 B read + 1 B written per element).
//
//   IntLayerNorm + QuantAct (quant_modules.py:353-386): int16 rows -> int8, 16-byte vector
//       loads, per-channel constants (bias, m, e) held in registers across the rows a warp
//       processes (a lane always owns the same channels).
#include <stdlib.h>

#include "ivit_common.cuh"
#include "ivit_internal.h"
#include "ivit_ptx.cuh"

namespace ivit {

// ------------------------------------------------------------------------------------
// LUT construction: lut[(mx+128)*256 + (q+128)] for -128 <= q <= mx <= 127 (q > mx unused -> 0)
// ------------------------------------------------------------------------------------
__global__ void gelu_lut_build_kernel(int32_t x0, float inv_x0, int n, const ivit_dyadic_t* __restrict__ me,
                                      int bits, int8_t* __restrict__ lut) {
    const int mx = (int)blockIdx.x - 128;
    const int q = (int)threadIdx.x - 128;
    const ivit_dyadic_t d = me[0];
    int32_t o = 0;
    if (q <= mx) {
        const long long Em = shiftexp(-mx, x0, inv_x0, n);
        const long long E = shiftexp(q - mx, x0, inv_x0, n);
        long long S = E + Em;
        S = S > 2147483647LL ? 2147483647LL : S;
        const long long F = 2147483647LL / S;
        const long long sig = (E * F) >> (31 - 8 + 1);
        o = clamp_i64_bits(requant64((long long)q * sig, d.m, d.e), bits);
    }
    lut[blockIdx.x * 256 + threadIdx.x] = (int8_t)o;
}

// ------------------------------------------------------------------------------------
// LUT application.  One warp per row; the row lives in registers as 16-byte vectors.
// Three rows per warp are in flight, one in each stage of
//     (a) global loads of the row            (b) row max -> load of its 256-byte table line
//     (c) table line -> shared memory, one byte lookup per element, 16-byte stores
// so that neither the row's memory round trip nor the dependent table-line load is exposed (the single-row form
// spent 2/3 of its warp time waiting on them: long-scoreboard stalls, profiles/ncu_full_r1h).
// ------------------------------------------------------------------------------------
template <int MAXV>
__global__ void __launch_bounds__(256)
gelu_lut_apply_kernel(const int8_t* __restrict__ q, int64_t rows, int cols, const int8_t* __restrict__ lut,
                      int8_t* __restrict__ out) {
    __shared__ __align__(256) uint8_t s_lut[8][256];
    ptx::grid_dep_wait();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nvec = cols >> 4;
    const int64_t warp0 = (int64_t)blockIdx.x * 8 + w;
    const int64_t nwarps = (int64_t)gridDim.x * 8;
    uint4 v[3][MAXV];
    uint2 line[3];
    auto stage_a = [&](int64_t row, uint4 (&vv)[MAXV]) {                 // loads of one row
        if (row >= rows) return;
        const uint4* src = reinterpret_cast<const uint4*>(q + row * (int64_t)cols);
#pragma unroll
        for (int j = 0; j < MAXV; ++j) {
            const int vi = lane + 32 * j;
            vv[j] = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);   // four int8 -128: neutral for the max
            if (vi < nvec) vv[j] = __ldg(src + vi);
        }
    };
    auto stage_b = [&](int64_t row, const uint4 (&vv)[MAXV], uint2& ln) {   // row max -> its table line (8 bytes per lane)
        if (row >= rows) return;
        // signed byte max with the native 16x2 SIMD max: a signed 16-bit compare is decided by its high byte, so
        // max.s16x2 over the words gives the max of bytes 3 and 1, over the words shifted left by 8 that of bytes 2 and 0
        uint32_t mo = 0x80008000u, me_ = 0x80008000u;
#pragma unroll
        for (int j = 0; j < MAXV; ++j) {
            const uint32_t in[4] = {vv[j].x, vv[j].y, vv[j].z, vv[j].w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                mo = __vmaxs2(mo, in[u]);
                me_ = __vmaxs2(me_, in[u] << 8);
            }
        }
        int32_t mx = max(max((int32_t)mo >> 24, (int32_t)(mo << 16) >> 24), max((int32_t)me_ >> 24, (int32_t)(me_ << 16) >> 24));
        mx = warp_max_i32(mx);
        ln = __ldg(reinterpret_cast<const uint2*>(lut + (mx + 128) * 256) + lane);
    };
    // The warp's table line sits at a 256-byte aligned shared address and is stored permuted by 0x80, so that the RAW
    // input byte is the index: the lookup address is one PRMT (byte k of the word | upper bytes of the base), the
    // lookup one LDS.U8, and four results are merged by three PRMTs -- 3.25 instructions per element.
    const uint32_t tl_base = (uint32_t)__cvta_generic_to_shared(&s_lut[w][0]);
    auto stage_c = [&](int64_t row, const uint4 (&vv)[MAXV], const uint2& ln) {
        if (row >= rows) return;
        __syncwarp();                                            // the previous row's lookups are done
        reinterpret_cast<uint2*>(s_lut[w])[lane ^ 16] = ln;      // entry (q + 128) -> index (q & 0xff)
        __syncwarp();
        uint4* dst = reinterpret_cast<uint4*>(out + row * (int64_t)cols);
#pragma unroll
        for (int j = 0; j < MAXV; ++j) {
            const int vi = lane + 32 * j;
            if (vi < nvec) {
                const uint32_t in[4] = {vv[j].x, vv[j].y, vv[j].z, vv[j].w};
                uint32_t o[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    uint32_t t[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        asm("ld.shared.u8 %0, [%1];" : "=r"(t[k]) : "r"(__byte_perm(in[u], tl_base, 0x7650 + k)));
                    o[u] = __byte_perm(__byte_perm(t[0], t[1], 0x0040), __byte_perm(t[2], t[3], 0x0040), 0x5410);
                }
                dst[vi] = make_uint4(o[0], o[1], o[2], o[3]);
            }
        }
    };
    // rows of this warp: r_i = warp0 + i * nwarps.  Iteration i runs (a) on r_{i+2}, (b) on r_{i+1}, (c) on r_i;
    // the loop is unrolled by three so that the register slots rotate statically.
    stage_a(warp0, v[0]);
    stage_a(warp0 + nwarps, v[1]);
    stage_b(warp0, v[0], line[0]);
    for (int64_t row = warp0; row < rows; row += 3 * nwarps) {
        stage_a(row + 2 * nwarps, v[2]);
        stage_b(row + nwarps, v[1], line[1]);
        stage_c(row, v[0], line[0]);
        stage_a(row + 3 * nwarps, v[0]);
        stage_b(row + 2 * nwarps, v[2], line[2]);
        stage_c(row + nwarps, v[1], line[1]);
        stage_a(row + 4 * nwarps, v[1]);
        stage_b(row + 3 * nwarps, v[0], line[0]);
        stage_c(row + 2 * nwarps, v[2], line[2]);
    }
}

// ------------------------------------------------------------------------------------
// LayerNorm int16 -> int8 with the per-channel QuantAct fused.  C = 8 * NV * 32 at most.
// ------------------------------------------------------------------------------------
struct alignas(16) LnCol { int32_t m; int32_t sh; long long c; };   // fast requant constants: hi32(z0*m + c) >> sh

__device__ __forceinline__ long long mul_wide_s32(int32_t a, int32_t b) {
    long long r;
    asm("mul.wide.s32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ long long mad_wide_s32(int32_t a, int32_t b, long long c) {
    long long r;
    asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c));
    return r;
}

// dp2a with signed 16-bit halves of a and unsigned / signed bytes 0, 1 of b:  a.lo * b.b0 + a.hi * b.b1 + c
__device__ __forceinline__ int32_t dp2a_lo_su(uint32_t a, uint32_t b, int32_t c) {
    int32_t d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// dp2a with signed 16-bit halves of a and signed bytes 2, 3 of b:  a.lo * b.b2 + a.hi * b.b3 + c
__device__ __forceinline__ int32_t dp2a_hi_ss(uint32_t a, uint32_t b, int32_t c) {
    int32_t d;
    asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int32_t dp2a_lo_ss(uint32_t a, uint32_t b, int32_t c) {
    int32_t d;
    asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// LPR lanes per row (32 / LPR rows per warp), NV 16-byte vectors (8 channels) per lane, FULL: C == 8 * NV * LPR.
// Everything that is per row (the two reductions, mean, integer sqrt, reciprocal factor) is computed redundantly by
// the row's LPR lanes, so narrow rows-per-warp splits (LPR = 16 for C = 768) halve that overhead per row.
//
// The row stays PACKED in registers (NV x 4 words of two int16) and is read in one statistics pass built on the
// 16 x 8-bit dot product (IDP.2A), two channels per instruction and 32-bit accumulators only:
//     sum = SUM x                       dp2a(w, {1, 1})
//     ssq = SUM x^2 = 256 * SUM x*xh + SUM x*xl     x = 256*xh + xl, xh signed high byte, xl unsigned low byte:
//                                       dp2a.s32.s32(w, {xh0, xh1}) and dp2a.s32.u32(w, {xl0, xl1}); |terms| < 2^23
//     V   = SUM (x - mu)^2 = ssq - mu * (2 * sum - C * mu)          exact in 64 bits (|.| < 2^42)
// and the second pass unpacks and centres in one instruction, y = dp2a(w, {1, 0} or {0, 1}, -mu).
// PRE: the next pair of rows is loaded into a second register set before this pair's arithmetic (2 blocks per SM);
// otherwise the kernel is compiled for MINB blocks per SM and relies on occupancy to cover the loads.
template <int NV, int LPR, bool FULL, int MINB, bool PRE>
__global__ void __launch_bounds__(256, MINB)
layernorm_i16_i8_kernel(const int16_t* __restrict__ x, int64_t rows, int C, const int32_t* __restrict__ bias_int,
                        const ivit_dyadic_t* __restrict__ me, int8_t* __restrict__ out) {
    constexpr int RPW = 32 / LPR;                                // rows per warp
    ptx::grid_dep_wait();
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPR, rsel = lane / LPR;
    const int nvec = C >> 3;                                     // vectors of 8 int16
    const int64_t warp0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * 8;
    // per-channel constants staged once per block.  Fast form (all channels): 32 <= e <= 62, no reachable
    // exact tie (|z| < 2^31: v2(z*m) <= 30 + ctz(m) < e-1) and |bias| < 2^30, so that
    //   RNE((floor(y*F/2) + b) * m / 2^e) == hi32(floor(y*F/2)*m + (b*m + 2^(e-1))) >> (e-32)
    __shared__ LnCol s_c[1024];
    __shared__ int32_t s_b[1024];
    auto load_row = [&](int64_t rbase, uint4 (&w)[NV]) {
        const int64_t row = rbase + rsel;
        const uint4* src = reinterpret_cast<const uint4*>(x + (row < rows ? row : rows - 1) * (int64_t)C);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int vi = sub + LPR * j;
            w[j] = make_uint4(0, 0, 0, 0);
            if (FULL || vi < nvec) w[j] = __ldg(src + vi);
        }
    };
    uint4 w[NV], wn[PRE ? NV : 1];
    int64_t rbase = warp0 * RPW;
    if (rbase < rows) load_row(rbase, w);                        // first rows in flight while the constants are staged
    int ok = 1;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int32_t b = bias_int[c];
        const ivit_dyadic_t d = me[c];
        const int tz = __ffs(d.m) - 1;
        const bool f = (d.e >= 32 && d.e <= 62) && (d.e - 1 - tz > 31) && (b > -(1 << 30)) && (b < (1 << 30));
        ok &= f ? 1 : 0;
        LnCol p;
        p.m = d.m; p.sh = d.e - 32;
        p.c = (d.e >= 1 && d.e <= 62) ? ((long long)b * (long long)d.m + (1LL << (d.e - 1))) : 0;
        // channel c = 8*vi + u is stored at [u][vi]: consecutive lanes (vi) read consecutive 16-byte entries (no bank conflicts)
        s_c[(c & 7) * nvec + (c >> 3)] = p;
        s_b[(c & 7) * nvec + (c >> 3)] = b;
    }
    const bool fast = __syncthreads_and(ok) != 0;
    const float inv_c = 1.0f / (float)C;
    for (; rbase < rows; rbase += nwarps * RPW) {
        const int64_t row = rbase + rsel;
        const bool row_ok = row < rows;
        const int64_t rnext = rbase + nwarps * RPW;
        if constexpr (PRE) { if (rnext < rows) load_row(rnext, wn); }   // in flight during this pair's arithmetic
        // ---- statistics over the packed row ----
        int32_t sum = 0;                                         // |sum| <= 1024 * 32768 < 2^31
        int32_t sh = 0, sl = 0;                                  // SUM x*xh (|.| <= 128 ch * 2^22), SUM x*xl (<= 128 ch * 2^23)
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const uint32_t tw[4] = {w[j].x, w[j].y, w[j].z, w[j].w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                sum = dp2a_lo_ss(tw[u], 0x0101u, sum);
                const uint32_t hl = __byte_perm(tw[u], 0u, 0x3120);            // bytes {xl0, xl1, xh0, xh1}
                sh = dp2a_hi_ss(tw[u], hl, sh);
                sl = dp2a_lo_su(tw[u], hl, sl);
            }
        }
        long long ssq = (long long)sh * 256 + (long long)sl;
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, o);
            ssq += __shfl_xor_sync(0xffffffffu, ssq, o);
        }
        // mu = RNE(sum / C): float estimate of the floor quotient + exact integer fix-up (|sum| < 2^26: estimate within 2)
        int32_t qd = (int32_t)floorf((float)sum * inv_c), rem = sum - qd * C;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            if (rem < 0) { qd -= 1; rem += C; }
            if (rem >= C) { qd += 1; rem -= C; }
        }
        if (2 * rem > C || (2 * rem == C && (qd & 1))) qd += 1;
        const int32_t mu = qd;
        const long long Vs = ssq - (long long)mu * (2LL * (long long)sum - (long long)C * (long long)mu);
        const unsigned long long k = ln_isqrt10((unsigned long long)Vs);
        const int32_t F = (int32_t)(k <= 0xffffffffULL ? (2147483647u / (uint32_t)k) : 0u);   // floor((2^31-1)/k), <= 2^31/64
        uint2* dst = reinterpret_cast<uint2*>(out + (row_ok ? row : rbase) * (int64_t)C);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int vi = sub + LPR * j;
            if ((FULL || vi < nvec) && row_ok) {
                const uint32_t tw[4] = {w[j].x, w[j].y, w[j].z, w[j].w};
                int32_t r[8], z[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int32_t y = dp2a_lo_ss(tw[u >> 1], (u & 1) ? 0x0100u : 0x0001u, -mu);   // x - mu, |.| <= 65535
                    z[u] = ((y * F) >> 1);                             // floor(y * F / 2), |.| <= 2^30
                    asm("" : "+r"(z[u]));                                                  // a plain 32-bit value from here on
                }
                if (fast) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int4 pw = *reinterpret_cast<const int4*>(&s_c[u * nvec + vi]);   // one 16-byte load: {m, sh, c}
                        const long long pc = (long long)(((unsigned long long)(uint32_t)pw.w << 32) | (uint32_t)pw.z);
                        r[u] = (int32_t)(((long long)z[u] * (long long)pw.x + pc) >> 32) >> pw.y;   // IMAD.HI with 64-bit addend
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const LnCol p = s_c[u * nvec + vi];
                        long long o = (long long)z[u] + (long long)s_b[u * nvec + vi];
                        o = o > 2147483647LL ? 2147483647LL : (o < -2147483648LL ? -2147483648LL : o);
                        r[u] = requant32_general((int32_t)o, p.m, p.sh + 32);
                    }
                }
                uint32_t lo, hi, w0, w1;
                asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(r[3]), "r"(r[2]), "r"(0));
                asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(w0) : "r"(r[1]), "r"(r[0]), "r"(hi));
                asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(lo) : "r"(r[7]), "r"(r[6]), "r"(0));
                asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(w1) : "r"(r[5]), "r"(r[4]), "r"(lo));
                dst[vi] = make_uint2(w0, w1);
            }
        }
        if (rnext < rows) {
            if constexpr (PRE) {
#pragma unroll
                for (int j = 0; j < NV; ++j) w[j] = wn[j];
            } else {
                load_row(rnext, w);
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// Four rows per warp.  The warp is split into groups of LPG lanes (32, 16 or 8); a group holds RPL = LPG / 8 of the four
// rows and every lane of it a 1/LPG slice of EACH of them (C = 8 * NVL * LPG: 768 = 3 x 32, 384 = 3 x 16, 192 = 3 x 8, ...),
// so one 16-byte shared-memory load of the per-channel constants serves RPL rows (the 16-lane kernel above spends 47 % of
// the shared-memory data pipe on them); a transposing butterfly inside the group leaves each run of 8 lanes with the sums
// of ONE of the four rows, whose scalars (mean, integer square root, reciprocal) it computes and hands to the lanes that
// need them -- half the scalar work per row of the 16-lane kernel.  The fast / general requant decision is hoisted out of
// the row loop, whole groups of rows use one base pointer and immediate offsets, and there is no second register set:
// a vector of the next rows is requested the moment the current one has been consumed (rolling prefetch).
// Same integers as layernorm_i16_i8_kernel.
// ------------------------------------------------------------------------------------
template <int NVL, int LPG>
__global__ void __launch_bounds__(256, 2)
layernorm_i16_i8_r4_kernel(const int16_t* __restrict__ x, int64_t rows, int C, const int32_t* __restrict__ bias_int,
                           const ivit_dyadic_t* __restrict__ me, int8_t* __restrict__ out) {
    static_assert(LPG == 32 || LPG == 16 || LPG == 8, "lanes per row group");
    constexpr int R = 4;                                         // rows per warp
    constexpr int RPL = LPG / 8;                                 // rows per lane (= rows per group)
    ptx::grid_dep_wait();
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPG, grp = lane / LPG;
    const int nvec = C >> 3;                                     // == LPG * NVL
    const int64_t warp0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * 8;
    __shared__ LnCol s_c[256 * 4];
    __shared__ int32_t s_b[256 * 4];
    uint4 w[RPL][NVL];
    // my group's rows [rb + grp * RPL, + RPL) of vector j; rows past the end re-read the last row (never stored)
    auto load_vec = [&](int64_t rb, int j) {
        if (rb + R <= rows) {                                    // warp-uniform: the common case has no per-row index math
            const uint4* src = reinterpret_cast<const uint4*>(x + (rb + grp * RPL) * (int64_t)C) + sub + LPG * j;
#pragma unroll
            for (int r = 0; r < RPL; ++r) w[r][j] = __ldg(src + r * nvec);
        } else {
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const int64_t want = rb + grp * RPL + r;
                const int64_t row = want < rows ? want : rows - 1;
                w[r][j] = __ldg(reinterpret_cast<const uint4*>(x + row * (int64_t)C) + sub + LPG * j);
            }
        }
    };
    int64_t rbase = warp0 * R;
    if (rbase < rows) {
#pragma unroll
        for (int j = 0; j < NVL; ++j) load_vec(rbase, j);
    }
    int ok = 1;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int32_t b = bias_int[c];
        const ivit_dyadic_t d = me[c];
        const int tz = __ffs(d.m) - 1;
        const bool f = (d.e >= 32 && d.e <= 62) && (d.e - 1 - tz > 31) && (b > -(1 << 30)) && (b < (1 << 30));
        ok &= f ? 1 : 0;
        LnCol p;
        p.m = d.m; p.sh = d.e - 32;
        p.c = (d.e >= 1 && d.e <= 62) ? ((long long)b * (long long)d.m + (1LL << (d.e - 1))) : 0;
        s_c[(c & 7) * nvec + (c >> 3)] = p;                      // channel 8*vi + u at [u][vi]: lanes read consecutive entries
        s_b[(c & 7) * nvec + (c >> 3)] = b;
    }
    const bool fast = __syncthreads_and(ok) != 0;
    const float inv_c = 1.0f / (float)C;
    const LnCol* sc_lane = s_c + sub;
    const int32_t* sb_lane = s_b + sub;
    for (; rbase < rows; rbase += nwarps * R) {
        const int64_t rnext = rbase + nwarps * R;
        const bool more = rnext < rows;
        // ---- statistics: per-lane partial sums of my rows
        // (vector by vector: the vector requested last -- at the end of the previous iteration -- is touched last)
        int32_t sum[RPL], sqh[RPL], sql[RPL];
        long long ssq[RPL];
#pragma unroll
        for (int r = 0; r < RPL; ++r) { sum[r] = 0; sqh[r] = 0; sql[r] = 0; }
#pragma unroll
        for (int j = 0; j < NVL; ++j) {
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const uint32_t tw[4] = {w[r][j].x, w[r][j].y, w[r][j].z, w[r][j].w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    sum[r] = dp2a_lo_ss(tw[u], 0x0101u, sum[r]);
                    const uint32_t hl = __byte_perm(tw[u], 0u, 0x3120);            // bytes {xl0, xl1, xh0, xh1}
                    sqh[r] = dp2a_hi_ss(tw[u], hl, sqh[r]);
                    sql[r] = dp2a_lo_su(tw[u], hl, sql[r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < RPL; ++r) ssq[r] = (long long)sqh[r] * 256 + (long long)sql[r];
        // transposing butterfly inside the group: every exchange halves the rows a lane holds, until each run of 8 lanes
        // holds one row (row (lane >> 3) & 3 of the warp's four), then a plain reduction over the 8 lanes
        int32_t rs;
        long long rq;
        if constexpr (RPL == 4) {
            int32_t s1[2];
            long long q1[2];
            const bool hi16 = (lane & 16) != 0;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int32_t keep_s = hi16 ? sum[2 + i] : sum[i], send_s = hi16 ? sum[i] : sum[2 + i];
                const long long keep_q = hi16 ? ssq[2 + i] : ssq[i], send_q = hi16 ? ssq[i] : ssq[2 + i];
                s1[i] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, 16);
                q1[i] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, 16);
            }
            const bool hi8 = (lane & 8) != 0;
            rs = (hi8 ? s1[1] : s1[0]) + __shfl_xor_sync(0xffffffffu, hi8 ? s1[0] : s1[1], 8);
            rq = (hi8 ? q1[1] : q1[0]) + __shfl_xor_sync(0xffffffffu, hi8 ? q1[0] : q1[1], 8);
        } else if constexpr (RPL == 2) {
            const bool hi8 = (lane & 8) != 0;
            rs = (hi8 ? sum[1] : sum[0]) + __shfl_xor_sync(0xffffffffu, hi8 ? sum[0] : sum[1], 8);
            rq = (hi8 ? ssq[1] : ssq[0]) + __shfl_xor_sync(0xffffffffu, hi8 ? ssq[0] : ssq[1], 8);
        } else {
            rs = sum[0];
            rq = ssq[0];
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            rs += __shfl_xor_sync(0xffffffffu, rs, o);
            rq += __shfl_xor_sync(0xffffffffu, rq, o);
        }
        // ---- row scalars of my run's row (as in layernorm_i16_i8_kernel) ----
        int32_t qd = (int32_t)floorf((float)rs * inv_c), rem = rs - qd * C;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            if (rem < 0) { qd -= 1; rem += C; }
            if (rem >= C) { qd += 1; rem -= C; }
        }
        if (2 * rem > C || (2 * rem == C && (qd & 1))) qd += 1;
        const long long Vs = rq - (long long)qd * (2LL * (long long)rs - (long long)C * (long long)qd);
        const unsigned long long k = ln_isqrt10((unsigned long long)Vs);
        const int32_t Fm = (int32_t)(k <= 0xffffffffULL ? (2147483647u / (uint32_t)k) : 0u);
        int32_t nmu[RPL], F[RPL];                                // of MY rows: warp row grp * RPL + r lives in lanes 8 * (...)
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            nmu[r] = -__shfl_sync(0xffffffffu, qd, 8 * (grp * RPL + r));
            F[r] = __shfl_sync(0xffffffffu, Fm, 8 * (grp * RPL + r));
        }
        // ---- normalise + per-channel requant: one constant load per channel and RPL rows ----
        const bool whole = rbase + R <= rows;                    // warp-uniform
        uint2* dst = reinterpret_cast<uint2*>(out + (rbase + grp * RPL) * (int64_t)C) + sub;
        auto word_of = [&](int r, int j, int u) -> uint32_t {
            return (u >> 1) == 0 ? w[r][j].x : ((u >> 1) == 1 ? w[r][j].y : ((u >> 1) == 2 ? w[r][j].z : w[r][j].w));
        };
        auto store_vec = [&](int j, const uint32_t (&pk)[RPL][2]) {
            if (whole) {
#pragma unroll
                for (int r = 0; r < RPL; ++r) dst[r * nvec + LPG * j] = make_uint2(pk[r][0], pk[r][1]);
            } else {
#pragma unroll
                for (int r = 0; r < RPL; ++r)
                    if (rbase + grp * RPL + r < rows) dst[r * nvec + LPG * j] = make_uint2(pk[r][0], pk[r][1]);
            }
        };
        auto pack4 = [&](const int32_t (&v)[4]) -> uint32_t {
            uint32_t hi2, lo2;
            asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi2) : "r"(v[3]), "r"(v[2]), "r"(0));
            asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(lo2) : "r"(v[1]), "r"(v[0]), "r"(hi2));
            return lo2;
        };
        if (fast) {
#pragma unroll
            for (int j = 0; j < NVL; ++j) {
                uint32_t pk[RPL][2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {                    // channels 4h .. 4h+3 of the vector
                    int32_t res[RPL][4];
#pragma unroll
                    for (int uu = 0; uu < 4; ++uu) {
                        const int u = 4 * h + uu;
                        const int4 pw = *reinterpret_cast<const int4*>(sc_lane + u * nvec + LPG * j);   // {m, sh, c}
                        const long long pc = (long long)(((unsigned long long)(uint32_t)pw.w << 32) | (uint32_t)pw.z);
#pragma unroll
                        for (int r = 0; r < RPL; ++r) {
                            const int32_t y = dp2a_lo_ss(word_of(r, j, u), (u & 1) ? 0x0100u : 0x0001u, nmu[r]);   // x - mu
                            // floor(y F / 2) from the 32-bit product: k >= floor(sqrt(V)) >= |y| for every element of
                            // the row (V is the sum of the y^2), so |y F| <= k floor((2^31 - 1) / k) < 2^31 -- IMAD + SHF
                            // instead of IMAD.WIDE (twice the pipe time, tools/ubench/pipes.cu) + 64-bit shift
                            int32_t z = ((y * F[r]) >> 1);
                            asm("" : "+r"(z));
                            res[r][uu] = (int32_t)(((long long)z * (long long)pw.x + pc) >> 32) >> pw.y;
                        }
                    }
#pragma unroll
                    for (int r = 0; r < RPL; ++r) pk[r][h] = pack4(res[r]);
                }
                store_vec(j, pk);
                if (more) load_vec(rnext, j);                    // this vector of the next rows: in flight from here on
            }
        } else {
            // general requants (a column outside the fast form): same structure, out-of-line arithmetic; fully unrolled so
            // that the packed rows stay in registers
#pragma unroll
            for (int j = 0; j < NVL; ++j) {
                uint32_t pk[RPL][2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    int32_t res[RPL][4];
#pragma unroll
                    for (int uu = 0; uu < 4; ++uu) {
                        const int u = 4 * h + uu;
                        const LnCol p = sc_lane[u * nvec + LPG * j];
                        const int32_t b = sb_lane[u * nvec + LPG * j];
#pragma unroll
                        for (int r = 0; r < RPL; ++r) {
                            const int32_t y = dp2a_lo_ss(word_of(r, j, u), (u & 1) ? 0x0100u : 0x0001u, nmu[r]);
                            const int32_t z = ((y * F[r]) >> 1);
                            long long o = (long long)z + (long long)b;
                            o = o > 2147483647LL ? 2147483647LL : (o < -2147483648LL ? -2147483648LL : o);
                            res[r][uu] = requant32_general((int32_t)o, p.m, p.sh + 32);
                        }
                    }
#pragma unroll
                    for (int r = 0; r < RPL; ++r) pk[r][h] = pack4(res[r]);
                }
                store_vec(j, pk);
            }
            if (more) {
#pragma unroll
                for (int j = 0; j < NVL; ++j) load_vec(rnext, j);
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// Stem: input QuantAct (fp32 -> int8, vit_quant.py:257) fused with the patch unfold of QuantConv2d
// (kernel == stride, layers_quant.py:190): one pass over the fp32 image, no int8 image round trip.
//   out[(b*Hp + y/p)*Wp + x/p, (c*p + y%p)*p + x%p] = clamp8(RNE(fp32(1/s) * img[b,c,y,x]))
// One thread per 4 horizontally adjacent pixels (p % 4 == 0): 16-byte loads, 4-byte stores.
// ------------------------------------------------------------------------------------
__global__ void quantize_patchify_kernel(const float4* __restrict__ img, const float* __restrict__ scale, int B, int Cin,
                                         int H, int W, int p, int8_t* __restrict__ out) {
    const float inv = __fdiv_rn(1.0f, scale[0]);
    const int W4 = W >> 2, Hp = H / p, Wp = W / p, K = Cin * p * p;
    const int64_t n4 = (int64_t)B * Cin * H * W4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const int x4 = (int)(i % W4);
        int64_t r = i / W4;
        const int y = (int)(r % H); r /= H;
        const int c = (int)(r % Cin);
        const int64_t b = r / Cin;
        const float4 v = __ldg(img + i);
        const int x = x4 * 4;
        const int64_t row = (b * Hp + y / p) * Wp + x / p;
        const int col = (c * p + y % p) * p + x % p;
        char4 o;
        o.x = (signed char)fminf(fmaxf(rintf(__fmul_rn(inv, v.x)), -128.f), 127.f);
        o.y = (signed char)fminf(fmaxf(rintf(__fmul_rn(inv, v.y)), -128.f), 127.f);
        o.z = (signed char)fminf(fmaxf(rintf(__fmul_rn(inv, v.z)), -128.f), 127.f);
        o.w = (signed char)fminf(fmaxf(rintf(__fmul_rn(inv, v.w)), -128.f), 127.f);
        *reinterpret_cast<char4*>(out + row * (int64_t)K + col) = o;
    }
}

// 16 x 16 patches (DeiT/ViT): one thread per (image, channel, row pair, patch column) = 2 x 16 pixels.  The two pixel
// rows of a patch are adjacent in the unfolded row ((c*16 + y%16)*16 + x%16), so the thread reads 2 x 64 contiguous
// bytes and writes ONE aligned 32-byte sector; consecutive lanes take consecutive patch columns of the same image
// rows, i.e. a warp reads whole image rows (the 4-pixel form wrote 4-byte pieces, half-empty sectors: 78 us -> 2.5 TB/s).
__global__ void __launch_bounds__(256)
quantize_patchify16_kernel(const float4* __restrict__ img, const float* __restrict__ scale, int B, int Cin, int H, int W,
                           int8_t* __restrict__ out) {
    const float inv = __fdiv_rn(1.0f, scale[0]);
    const int Wp = W >> 4, Hp = H >> 4, H2 = H >> 1, W4 = W >> 2, K = Cin * 256;
    const int64_t n = (int64_t)B * Cin * H2 * Wp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int px = (int)(i % Wp);
        int64_t r = i / Wp;
        const int y = 2 * (int)(r % H2); r /= H2;
        const int c = (int)(r % Cin);
        const int64_t b = r / Cin;
        const float4* src = img + ((b * Cin + c) * H + y) * (int64_t)W4 + px * 4;
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) { v[j] = __ldg(src + j); v[4 + j] = __ldg(src + W4 + j); }
        uint32_t o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int32_t q0 = __float2int_rn(fminf(fmaxf(__fmul_rn(inv, v[j].x), -128.f), 127.f));
            const int32_t q1 = __float2int_rn(fminf(fmaxf(__fmul_rn(inv, v[j].y), -128.f), 127.f));
            const int32_t q2 = __float2int_rn(fminf(fmaxf(__fmul_rn(inv, v[j].z), -128.f), 127.f));
            const int32_t q3 = __float2int_rn(fminf(fmaxf(__fmul_rn(inv, v[j].w), -128.f), 127.f));
            uint32_t hi;
            asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(q3), "r"(q2), "r"(0));
            asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(o[j]) : "r"(q1), "r"(q0), "r"(hi));
        }
        const int64_t row = (b * Hp + (y >> 4)) * Wp + px;
        uint4* dst = reinterpret_cast<uint4*>(out + row * (int64_t)K + (c * 16 + (y & 15)) * 16);
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
}

// uint8 input: the whole fp32 chain of the reference's eval transform + input QuantAct depends only on (channel, byte):
//   t = u / 255 (ToTensor); x = (t - mean[c]) / std[c] (Normalize, utils/data_utils.py:90-91); q = clamp8(RNE(x * (1/s)))
// -> a Cin x 256 table built per launch by the first threads of each block (same IEEE operations as torch: div, sub, div,
// mul, round-half-even), applied as a byte lookup.  One thread per (image, channel, row pair, patch column) as above.
__global__ void __launch_bounds__(256)
quantize_patchify16_u8_kernel(const uint8_t* __restrict__ img, const float* __restrict__ mean, const float* __restrict__ stdv,
                              const float* __restrict__ scale, int B, int Cin, int H, int W, int8_t* __restrict__ out) {
    __shared__ __align__(16) int8_t s_tab[4][256];
    {
        const float inv = __fdiv_rn(1.0f, scale[0]);
        for (int i = threadIdx.x; i < Cin * 256; i += blockDim.x) {
            const int c = i >> 8, u = i & 255;
            const float t = __fdiv_rn((float)u, 255.0f);
            const float x = __fdiv_rn(__fsub_rn(t, mean[c]), stdv[c]);
            s_tab[c][u] = (int8_t)fminf(fmaxf(rintf(__fmul_rn(inv, x)), -128.f), 127.f);
        }
    }
    __syncthreads();
    const int Wp = W >> 4, Hp = H >> 4, H2 = H >> 1, K = Cin * 256;
    const int64_t n = (int64_t)B * Cin * H2 * Wp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int px = (int)(i % Wp);
        int64_t r = i / Wp;
        const int y = 2 * (int)(r % H2); r /= H2;
        const int c = (int)(r % Cin);
        const int64_t b = r / Cin;
        const uint8_t* src = img + ((b * Cin + c) * H + y) * (int64_t)W + px * 16;
        const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(src)), v1 = __ldg(reinterpret_cast<const uint4*>(src + W));
        const uint32_t in[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        const int8_t* tab = s_tab[c];
        uint32_t o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
            o[j] = (uint32_t)(uint8_t)tab[in[j] & 0xff] | ((uint32_t)(uint8_t)tab[(in[j] >> 8) & 0xff] << 8) |
                   ((uint32_t)(uint8_t)tab[(in[j] >> 16) & 0xff] << 16) | ((uint32_t)(uint8_t)tab[in[j] >> 24] << 24);
        const int64_t row = (b * Hp + (y >> 4)) * Wp + px;
        uint4* dst = reinterpret_cast<uint4*>(out + row * (int64_t)K + (c * 16 + (y & 15)) * 16);
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
}

// ------------------------------------------------------------------------------------
// Stem: cls-token concatenation + position-embedding residual QuantAct (vit_quant.py:259-265), 8 channels per
// thread, scalar dyadics analysed on the host (same unified form as the GEMM epilogue's second stage).
// ------------------------------------------------------------------------------------
struct EmbRq { int32_t m, ls, rs, himask; long long half; int tie; };

__device__ __forceinline__ int32_t emb_rq(const EmbRq& u, int32_t z) {
    const long long t = (long long)(z << u.ls) * (long long)u.m + u.half;
    const int32_t hi = (int32_t)(t >> 32);
    int32_t q = hi >> u.rs;
    if (u.tie) {
        const bool tie = ((uint32_t)t == 0u) && ((hi & u.himask) == 0);
        q -= (int32_t)(tie & (q & 1));
    }
    return q;
}

// One warp per output row (token t of image b), lanes over the 16-byte vectors of the row: the (image, token) split is
// one 32-bit division per ROW (the first version divided a 64-bit element index per vector: 59 us for DeiT-B bs=256
// against an HBM floor of 24 us), and the class-token row -- not a QuantAct output, it may exceed 16 bits and takes the
// exact general requant -- is a warp-uniform branch.
__global__ void embed_tokens_fast_kernel(const int16_t* __restrict__ pe, const int32_t* __restrict__ cls,
                                         const int16_t* __restrict__ pos, int B, int n_tok, int C, EmbRq rq, EmbRq rqp,
                                         ivit_dyadic_t me, int16_t* __restrict__ out) {
    ptx::grid_dep_wait();
    const int C8 = C >> 3;
    const int lane = threadIdx.x & 31;
    const uint32_t rows = (uint32_t)B * (uint32_t)n_tok;                 // < 2^31 (host-checked)
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += nwarps) {
        const uint32_t b = r / (uint32_t)n_tok;
        const int t = (int)(r - b * (uint32_t)n_tok);
        const uint4* prow = reinterpret_cast<const uint4*>(pos + (int64_t)t * C);
        uint4* orow = reinterpret_cast<uint4*>(out + (int64_t)r * C);
        const uint4* xrow = reinterpret_cast<const uint4*>(pe + ((int64_t)b * (n_tok - 1) + (t > 0 ? t - 1 : 0)) * (int64_t)C);
        for (int c8 = lane; c8 < C8; c8 += 32) {
            const uint4 pv = __ldg(prow + c8);
            const uint32_t pw[4] = {pv.x, pv.y, pv.z, pv.w};
            int32_t a[8];
            if (t == 0) {
#pragma unroll
                for (int u = 0; u < 8; ++u) a[u] = requant32_general(cls[c8 * 8 + u], me.m, me.e);
            } else {
                const uint4 v = __ldg(xrow + c8);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    a[2 * u] = emb_rq(rq, (int32_t)(int16_t)(w[u] & 0xffff));
                    a[2 * u + 1] = emb_rq(rq, (int32_t)w[u] >> 16);
                }
            }
            uint32_t o[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int32_t q0 = a[2 * u] + emb_rq(rqp, (int32_t)(int16_t)(pw[u] & 0xffff));
                const int32_t q1 = a[2 * u + 1] + emb_rq(rqp, (int32_t)pw[u] >> 16);
                asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(o[u]) : "r"(q1), "r"(q0));
            }
            orow[c8] = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

static bool make_emb_rq(ivit_dyadic_t d, EmbRq* r) {
    if (d.m == 0 || d.e < 16 || d.e > 62) return false;
    r->m = d.m;
    if (d.e < 32) { r->ls = 32 - d.e; r->rs = 0; r->half = 1LL << 31; }
    else { r->ls = 0; r->rs = d.e - 32; r->half = 1LL << (d.e - 1); }
    r->himask = (1 << r->rs) - 1;
    r->tie = (d.e - 1 - __builtin_ctz((unsigned)d.m) <= 15) ? 1 : 0;
    return true;
}

}  // namespace ivit

using namespace ivit;

extern "C" {

int ivit_shiftgelu_build_lut(ivit_ctx* ctx, int32_t x0, int n, const ivit_dyadic_t* me, int bits, int8_t* lut,
                             ivit_stream stream) {
    IVIT_REQUIRE(ctx && me && lut, "ivit_shiftgelu_build_lut: null pointer");
    IVIT_REQUIRE(bits == 8, "ivit_shiftgelu_build_lut: the table holds int8 outputs (bits == 8)");
    IVIT_REQUIRE(n >= 1 && n <= 30, "ivit_shiftgelu_build_lut: bad n");
    if (!(x0 <= -1 && x0 >= -(1 << 24)))
        return fail(IVIT_ENOTSUP, "ivit_shiftgelu_build_lut: x0=%d outside [-2^24, -1]", x0);
    gelu_lut_build_kernel<<<256, 256, 0, st(stream)>>>(x0, 1.0f / (float)x0, n, me, bits, lut);
    IVIT_LAUNCH_OK("gelu_lut_build_kernel");
    return IVIT_OK;
}

int ivit_shiftgelu_lut(ivit_ctx* ctx, const int8_t* q, int64_t rows, int cols, const int8_t* lut, int8_t* out,
                       ivit_stream stream) {
    IVIT_REQUIRE(ctx && q && lut && out && rows > 0 && cols > 0, "ivit_shiftgelu_lut: bad arguments");
    IVIT_REQUIRE(cols % 16 == 0 && cols <= 16 * 32 * 8, "ivit_shiftgelu_lut: cols must be a multiple of 16, <= 4096");
    IVIT_REQUIRE(((uintptr_t)q % 16) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)lut % 8) == 0,
                 "ivit_shiftgelu_lut: q/out must be 16-byte aligned");
    const int nv = (cols / 16 + 31) / 32;
    // persistent grid: exactly the blocks that are resident at once (a partial second wave would run alone at the end)
#define GL(MAXV) do {                                                                                                    \
        static PerDevice bps_dev;                                                                                        \
        int& bps = bps_dev[ctx->device];                                                                                 \
        if (!bps) IVIT_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, gelu_lut_apply_kernel<MAXV>, 256, 0)); \
        const int64_t want = (rows + 7) / 8, cap = (int64_t)ctx->num_sms * (bps > 0 ? bps : 1);                          \
        IVIT_CUDA_OK(launch_k(gelu_lut_apply_kernel<MAXV>, dim3((unsigned)(want < cap ? want : cap)), dim3(256), 0, st(stream), q, rows, cols, lut, out)); \
    } while (0)
    if (nv <= 2) GL(2); else if (nv <= 4) GL(4); else if (nv <= 6) GL(6); else GL(8);
#undef GL
    IVIT_LAUNCH_OK("gelu_lut_apply_kernel");
    return IVIT_OK;
}

int ivit_quantize_patchify(ivit_ctx* ctx, const float* img, const float* scale, int B, int Cin, int H, int W, int p,
                           int8_t* out, ivit_stream stream) {
    IVIT_REQUIRE(ctx && img && scale && out && B > 0 && Cin > 0 && p > 0, "ivit_quantize_patchify: bad arguments");
    IVIT_REQUIRE(H % p == 0 && W % p == 0 && p % 4 == 0 && W % 4 == 0, "ivit_quantize_patchify: H, W multiples of the patch size, patch % 4 == 0");
    IVIT_REQUIRE(((uintptr_t)img % 16) == 0 && ((uintptr_t)out % 4) == 0, "ivit_quantize_patchify: img must be 16-byte aligned");
    if (p == 16 && ((uintptr_t)out % 16) == 0) {
        const int64_t n = (int64_t)B * Cin * (H / 2) * (W / 16);
        const int64_t blocks = (n + 255) / 256;
        const int grid = (int)(blocks < (int64_t)ctx->num_sms * 8 ? blocks : (int64_t)ctx->num_sms * 8);
        quantize_patchify16_kernel<<<grid, 256, 0, st(stream)>>>((const float4*)img, scale, B, Cin, H, W, out);
        IVIT_LAUNCH_OK("quantize_patchify16_kernel");
        return IVIT_OK;
    }
    const int64_t n4 = (int64_t)B * Cin * H * (W / 4);
    const int64_t blocks = (n4 + 255) / 256;
    const int grid = (int)(blocks < (int64_t)ctx->num_sms * 16 ? blocks : (int64_t)ctx->num_sms * 16);
    quantize_patchify_kernel<<<grid, 256, 0, st(stream)>>>((const float4*)img, scale, B, Cin, H, W, p, out);
    IVIT_LAUNCH_OK("quantize_patchify_kernel");
    return IVIT_OK;
}

int ivit_quantize_patchify_u8(ivit_ctx* ctx, const uint8_t* img, const float* mean, const float* stdv, const float* scale,
                              int B, int Cin, int H, int W, int p, int8_t* out, ivit_stream stream) {
    IVIT_REQUIRE(ctx && img && mean && stdv && scale && out && B > 0, "ivit_quantize_patchify_u8: bad arguments");
    IVIT_REQUIRE(p == 16 && Cin >= 1 && Cin <= 4, "ivit_quantize_patchify_u8: patch 16, at most 4 channels");
    IVIT_REQUIRE(H % 16 == 0 && W % 16 == 0, "ivit_quantize_patchify_u8: H, W multiples of the patch size");
    IVIT_REQUIRE(((uintptr_t)img % 16) == 0 && ((uintptr_t)out % 16) == 0, "ivit_quantize_patchify_u8: 16-byte alignment");
    const int64_t n = (int64_t)B * Cin * (H / 2) * (W / 16);
    const int64_t blocks = (n + 255) / 256;
    const int grid = (int)(blocks < (int64_t)ctx->num_sms * 8 ? blocks : (int64_t)ctx->num_sms * 8);
    quantize_patchify16_u8_kernel<<<grid, 256, 0, st(stream)>>>(img, mean, stdv, scale, B, Cin, H, W, out);
    IVIT_LAUNCH_OK("quantize_patchify16_u8_kernel");
    return IVIT_OK;
}

int ivit_embed_tokens_fast(ivit_ctx* ctx, const int16_t* pe, const int32_t* cls, const int16_t* pos, int B, int n_tok,
                           int C, ivit_dyadic_t me, ivit_dyadic_t me_res, int16_t* out, ivit_stream stream) {
    IVIT_REQUIRE(ctx && pe && cls && pos && out && B > 0 && n_tok > 1 && C > 0 && C % 8 == 0, "ivit_embed_tokens_fast: bad arguments (C % 8 == 0)");
    IVIT_REQUIRE(((uintptr_t)pe % 16) == 0 && ((uintptr_t)pos % 16) == 0 && ((uintptr_t)out % 16) == 0, "ivit_embed_tokens_fast: 16-byte alignment");
    EmbRq rq, rqp;
    if (!make_emb_rq(me, &rq) || !make_emb_rq(me_res, &rqp))
        return fail(IVIT_ENOTSUP, "ivit_embed_tokens_fast: dyadic exponents outside [16, 62]; use ivit_embed_tokens");
    IVIT_REQUIRE((int64_t)B * n_tok < (1LL << 31), "ivit_embed_tokens_fast: more than 2^31 tokens");
    const int64_t blocks = ((int64_t)B * n_tok + 7) / 8;                         // one warp per row, 8 warps per block
    const int grid = (int)(blocks < (int64_t)ctx->num_sms * 8 ? blocks : (int64_t)ctx->num_sms * 8);
    IVIT_CUDA_OK(launch_k(embed_tokens_fast_kernel, dim3(grid), dim3(256), 0, st(stream), pe, cls, pos, B, n_tok, C, rq, rqp, me, out));
    IVIT_LAUNCH_OK("embed_tokens_fast_kernel");
    return IVIT_OK;
}

int ivit_layernorm_i16_i8(ivit_ctx* ctx, const int16_t* x, int64_t rows, int C, const int32_t* bias_int,
                          const ivit_dyadic_t* me, int8_t* out, ivit_stream stream) {
    IVIT_REQUIRE(ctx && x && bias_int && me && out && rows > 0, "ivit_layernorm_i16_i8: bad arguments");
    IVIT_REQUIRE(C % 8 == 0 && C >= 8 && C <= 8 * 32 * 4, "ivit_layernorm_i16_i8: C must be a multiple of 8, <= 1024");
    IVIT_REQUIRE(((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 8) == 0, "ivit_layernorm_i16_i8: x must be 16-byte aligned");
    // 16 lanes per row (two rows per warp), up to 8 vectors of 8 channels per lane (C <= 1024)
    const int nvec = C / 8;
    const int lpr = 16;
    const int rpb = 8 * (32 / lpr);                              // rows per 256-thread block and pass
    const int64_t want = (rows + rpb - 1) / rpb;
    // persistent grid: two resident 256-thread blocks per SM (126 registers with the prefetched second register set).
    // IVIT_LN_VARIANT (experiments, tools/rowops_bench.py): 1 = three blocks per SM (<= 80 registers, no prefetch set),
    // 2 = four blocks per SM (<= 64 registers, no prefetch set)
    static const char* var_env = getenv("IVIT_LN_VARIANT");
    const int variant = var_env ? atoi(var_env) : 0;
    // 3 / 4 (C == 768 only): eight lanes per row (four rows per warp: the per-row scalar work -- reductions, mean, integer
    // square root, reciprocal -- is amortised over twice as many rows), 12 vectors per lane, without / with the prefetch set
    if ((variant == 3 || variant == 4) && C == 768) {
        const int rpb8 = 8 * 4;
        const int64_t want8 = (rows + rpb8 - 1) / rpb8;
        const int grid8 = (int)(want8 < (int64_t)ctx->num_sms * 2 ? want8 : (int64_t)ctx->num_sms * 2);
        if (variant == 3) layernorm_i16_i8_kernel<12, 8, true, 2, false><<<grid8, 256, 0, st(stream)>>>(x, rows, C, bias_int, me, out);
        else layernorm_i16_i8_kernel<12, 8, true, 1, true><<<grid8 / 2 > 0 ? grid8 / 2 : 1, 256, 0, st(stream)>>>(x, rows, C, bias_int, me, out);
        IVIT_LAUNCH_OK("layernorm_i16_i8_kernel");
        return IVIT_OK;
    }
    // four rows per warp, C = 8 * NVL * LPG with LPG = 32 / 16 / 8 lanes per row group and NVL <= 4 vectors per lane:
    // 768 = 3 x 32 (DeiT-B), 1024 = 4 x 32 (ViT-L), 384 = 3 x 16 (DeiT-S), 192 = 3 x 8 (DeiT-T), 512, 256, 128, 64
    // (IVIT_LN_VARIANT=9: the 16-lane kernel below, for A/B timing)
    if (variant == 0 && C % 64 == 0) {
        const int64_t want4 = (rows + 31) / 32;
        const int grid4 = (int)(want4 < (int64_t)ctx->num_sms * 2 ? want4 : (int64_t)ctx->num_sms * 2);
#define LN4(NVLV, LPGV) IVIT_CUDA_OK(launch_k(layernorm_i16_i8_r4_kernel<NVLV, LPGV>, dim3(grid4), dim3(256), 0, st(stream), x, rows, C, bias_int, me, out))
        bool done = true;
        switch (C) {
            case 1024: LN4(4, 32); break;
            case 768: LN4(3, 32); break;
            case 512: LN4(2, 32); break;
            case 256: LN4(1, 32); break;
            case 384: LN4(3, 16); break;
            case 192: LN4(3, 8); break;
            case 128: LN4(2, 8); break;
            case 64: LN4(1, 8); break;
            default: done = false; break;
        }
#undef LN4
        if (done) {
            IVIT_LAUNCH_OK("layernorm_i16_i8_r4_kernel");
            return IVIT_OK;
        }
    }
    const int bps = variant == 1 ? 3 : (variant == 2 ? 4 : 2);
    const int grid = (int)(want < (int64_t)ctx->num_sms * bps ? want : (int64_t)ctx->num_sms * bps);
    const int nv = (nvec + lpr - 1) / lpr;
    const bool full = (nv * lpr == nvec);
#define LNK(NV, LPR, FULLV) do { \
        if (variant == 1) layernorm_i16_i8_kernel<NV, LPR, FULLV, 3, false><<<grid, 256, 0, st(stream)>>>(x, rows, C, bias_int, me, out); \
        else if (variant == 2) layernorm_i16_i8_kernel<NV, LPR, FULLV, 4, false><<<grid, 256, 0, st(stream)>>>(x, rows, C, bias_int, me, out); \
        else IVIT_CUDA_OK(launch_k(layernorm_i16_i8_kernel<NV, LPR, FULLV, 2, true>, dim3(grid), dim3(256), 0, st(stream), x, rows, C, bias_int, me, out)); } while (0)
#define LNF(NV, LPR) do { if (full) LNK(NV, LPR, true); else LNK(NV, LPR, false); } while (0)
    switch (nv) { case 1: LNF(1, 16); break; case 2: LNF(2, 16); break; case 3: LNF(3, 16); break; case 4: LNF(4, 16); break;
                  case 5: LNF(5, 16); break; case 6: LNF(6, 16); break; case 7: LNF(7, 16); break; default: LNF(8, 16); break; }
#undef LNK
#undef LNF
    IVIT_LAUNCH_OK("layernorm_i16_i8_kernel");
    return IVIT_OK;
}

}  // extern "C"
