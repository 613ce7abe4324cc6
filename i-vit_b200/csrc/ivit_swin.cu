// Swin glue folded into row kernels (swin_quant.py:251-301, 328-349, 539-558): no separate roll / window-partition /
// window-reverse / 2x2-merge copies and no eager tensor ops on the hot path.
//
//   layernorm_gather_kernel   IntLayerNorm + per-channel QuantAct (int16 -> int8, same arithmetic as
//                             layernorm_i16_i8_kernel in ivit_fast.cu) whose INPUT rows are gathered through a per-image
//                             row map:
//                               G = 1  out row r <- in row map[r]: the cyclic shift + window partition of the next block
//                                      (torch.roll + window_partition, swin_quant.py:259-271) and the window reverse of
//                                      the previous one (:278-288) composed into one permutation; the gathered int16 row
//                                      is also written out (the residual stream in the new order, read by the proj GEMM's
//                                      residual epilogue)
//                               G = 4  out row r <- the concatenation of four in rows map[4r..4r+3]: PatchMerging's
//                                      strided 2x2 gather + cat (:337-341) feeding its LayerNorm(4C) (:344-345)
//                               IN8 / OUT16: the patch embedding's norm (layers_quant.py:193-195 + swin_quant.py:546) reads the
//                                      8-bit qact_before_norm output and applies TWO 16-bit QuantActs in a row
//                                      (patch_embed.qact per channel, then the model's qact1, scalar) -> int16 stream
//   avgpool_requant_kernel    token average RNE(sum / L) (AdaptiveAvgPool1d on the carrier, :554) + qact3 (:555)
#include <stdlib.h>

#include "ivit_common.cuh"
#include "ivit_internal.h"
#include "ivit_ptx.cuh"

namespace ivit {

struct alignas(16) LnColG { int32_t m; int32_t sh; long long c; };   // fast requant constants: hi32(z0*m + c) >> sh

__device__ __forceinline__ long long g_mul_wide_s32(int32_t a, int32_t b) {
    long long r;
    asm("mul.wide.s32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ int32_t g_dp2a_lo_su(uint32_t a, uint32_t b, int32_t c) {
    int32_t d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// dp2a with signed 16-bit halves of a and signed bytes 2, 3 of b:  a.lo * b.b2 + a.hi * b.b3 + c
__device__ __forceinline__ int32_t g_dp2a_hi_ss(uint32_t a, uint32_t b, int32_t c) {
    int32_t d;
    asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int32_t g_dp2a_lo_ss(uint32_t a, uint32_t b, int32_t c) {
    int32_t d;
    asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// prmt.b32 in its generic form: selector nibbles with bit 3 set replicate the SIGN of the selected byte (the
// __byte_perm intrinsic only honours the low three bits of a nibble)
__device__ __forceinline__ uint32_t g_prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

constexpr int LNG_MAXC = 2048;      // 4 * 512: the widest merge LayerNorm of the Swin zoo (Swin-B, stage 3 -> 4)

// LPR lanes per row (32 / LPR rows per warp), NV 16-byte vectors (8 channels) per lane; C == 8 * NV * LPR when FULL.
// G source rows per output row (1 or 4), each Cs = C / G channels wide.  Statistics and arithmetic: see
// layernorm_i16_i8_kernel (one packed pass with IDP.2A, exact 64-bit variance, closed-form integer square root).
template <int NV, int LPR, bool FULL, int G, bool IO16 = false>
__global__ void __launch_bounds__(256, 2)
layernorm_gather_kernel(const int16_t* __restrict__ x, int64_t rows, int C, const int32_t* __restrict__ rowmap, int L_out,
                        int L_in, const int32_t* __restrict__ bias_int, const ivit_dyadic_t* __restrict__ me,
                        int8_t* __restrict__ out, int16_t* __restrict__ xcopy, ivit_dyadic_t me2 = ivit_dyadic_t{0, 0}) {
    // IO16 (G == 1, no map): x is INT8 [rows, C]; out is INT16 [rows, C] = clamp16(RNE(clamp16(RNE(y*m/2^e)) * m2 / 2^e2))
    constexpr int RPW = 32 / LPR;
    ptx::grid_dep_wait();
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPR, rsel = lane / LPR;
    const int nvec = C >> 3;
    const int nvec_s = nvec / G;                                 // vectors per source row
    const int Cs = C / G;
    const int64_t warp0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * 8;
    __shared__ LnColG s_c[LNG_MAXC];
    __shared__ int32_t s_b[LNG_MAXC];
    auto load_row = [&](int64_t rbase, uint4 (&w)[NV]) {
        int64_t row = rbase + rsel;
        row = row < rows ? row : rows - 1;
        const uint32_t img = (uint32_t)row / (uint32_t)L_out;
        const uint32_t r = (uint32_t)row - img * (uint32_t)L_out;
        const int64_t in0 = (int64_t)img * L_in;
        if constexpr (IO16) {
            // 8 int8 channels per vector -> four words of two sign-extended int16 (PRMT with sign replication)
            const uint2* src = reinterpret_cast<const uint2*>(reinterpret_cast<const int8_t*>(x) + (in0 + r) * (int64_t)C);
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const int vi = sub + LPR * j;
                w[j] = make_uint4(0, 0, 0, 0);
                if (FULL || vi < nvec) {
                    const uint2 v = __ldg(src + vi);
                    w[j] = make_uint4(g_prmt(v.x, 0u, 0x9180u), g_prmt(v.x, 0u, 0xB3A2u),
                                      g_prmt(v.y, 0u, 0x9180u), g_prmt(v.y, 0u, 0xB3A2u));
                }
            }
        } else if constexpr (G == 1) {
            const int sr = rowmap ? __ldg(rowmap + r) : (int)r;
            const uint4* src = reinterpret_cast<const uint4*>(x + (in0 + sr) * (int64_t)C);
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const int vi = sub + LPR * j;
                w[j] = make_uint4(0, 0, 0, 0);
                if (FULL || vi < nvec) w[j] = __ldg(src + vi);
            }
        } else {
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const int vi = sub + LPR * j;
                w[j] = make_uint4(0, 0, 0, 0);
                if (FULL || vi < nvec) {
                    const int g = vi / nvec_s, vo = vi - g * nvec_s;
                    const int sr = __ldg(rowmap + r * G + g);
                    w[j] = __ldg(reinterpret_cast<const uint4*>(x + (in0 + sr) * (int64_t)Cs) + vo);
                }
            }
        }
    };
    uint4 w[NV], wn[NV];
    int64_t rbase = warp0 * RPW;
    if (rbase < rows) load_row(rbase, w);                        // first rows in flight while the constants are staged
    int ok = 1;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int32_t b = bias_int[c];
        const ivit_dyadic_t d = me[c];
        const int tz = __ffs(d.m) - 1;
        const bool f = (d.e >= 32 && d.e <= 62) && (d.e - 1 - tz > 31) && (b > -(1 << 30)) && (b < (1 << 30));
        ok &= f ? 1 : 0;
        LnColG p;
        p.m = d.m; p.sh = d.e - 32;
        p.c = (d.e >= 1 && d.e <= 62) ? ((long long)b * (long long)d.m + (1LL << (d.e - 1))) : 0;
        s_c[(c & 7) * nvec + (c >> 3)] = p;                      // channel 8*vi + u at [u][vi]: conflict-free 16-byte reads
        s_b[(c & 7) * nvec + (c >> 3)] = b;
    }
    const bool fast = __syncthreads_and(ok) != 0;
    const float inv_c = 1.0f / (float)C;
    const UniRq rq2 = make_unirq(me2, 16);                       // IO16: the second (scalar) QuantAct on a 16-bit operand
    for (; rbase < rows; rbase += nwarps * RPW) {
        const int64_t row = rbase + rsel;
        const bool row_ok = row < rows;
        const int64_t rnext = rbase + nwarps * RPW;
        if (rnext < rows) load_row(rnext, wn);                   // in flight during this row's arithmetic
        if (!IO16 && G == 1 && xcopy != nullptr && row_ok) {              // the gathered row in the new order (residual stream)
            uint4* dstx = reinterpret_cast<uint4*>(xcopy + row * (int64_t)C);
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const int vi = sub + LPR * j;
                if (FULL || vi < nvec) dstx[vi] = w[j];
            }
        }
        int32_t sum = 0, sh = 0, sl = 0;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const uint32_t tw[4] = {w[j].x, w[j].y, w[j].z, w[j].w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                sum = g_dp2a_lo_ss(tw[u], 0x0101u, sum);
                const uint32_t hl = __byte_perm(tw[u], 0u, 0x3120);            // bytes {xl0, xl1, xh0, xh1}
                sh = g_dp2a_hi_ss(tw[u], hl, sh);
                sl = g_dp2a_lo_su(tw[u], hl, sl);
            }
        }
        long long ssq = (long long)sh * 256 + (long long)sl;
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, o);
            ssq += __shfl_xor_sync(0xffffffffu, ssq, o);
        }
        // mu = RNE(sum / C): float estimate of the floor quotient + exact integer fix-up (|sum| <= 1536 * 2^15 < 2^26)
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
        const int32_t F = (int32_t)(k <= 0xffffffffULL ? (2147483647u / (uint32_t)k) : 0u);
        uint2* dst = reinterpret_cast<uint2*>(out + (row_ok ? row : rbase) * (int64_t)C);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int vi = sub + LPR * j;
            if ((FULL || vi < nvec) && row_ok) {
                const uint32_t tw[4] = {w[j].x, w[j].y, w[j].z, w[j].w};
                int32_t r[8], z[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int32_t y = g_dp2a_lo_ss(tw[u >> 1], (u & 1) ? 0x0100u : 0x0001u, -mu);   // x - mu
                    z[u] = ((y * F) >> 1);                                     // floor(y * F / 2)
                    asm("" : "+r"(z[u]));
                }
                if (fast) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int4 pw = *reinterpret_cast<const int4*>(&s_c[u * nvec + vi]);
                        const long long pc = (long long)(((unsigned long long)(uint32_t)pw.w << 32) | (uint32_t)pw.z);
                        r[u] = (int32_t)(((long long)z[u] * (long long)pw.x + pc) >> 32) >> pw.y;
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const LnColG p = s_c[u * nvec + vi];
                        long long o = (long long)z[u] + (long long)s_b[u * nvec + vi];
                        o = o > 2147483647LL ? 2147483647LL : (o < -2147483648LL ? -2147483648LL : o);
                        r[u] = requant32_general((int32_t)o, p.m, p.sh + 32);
                    }
                }
                if constexpr (IO16) {
                    uint32_t ow[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int32_t a = clamp_bits<16>(unirq_apply(rq2, clamp_bits<16>(r[2 * u])));
                        const int32_t b = clamp_bits<16>(unirq_apply(rq2, clamp_bits<16>(r[2 * u + 1])));
                        ow[u] = ((uint32_t)a & 0xffffu) | ((uint32_t)b << 16);
                    }
                    reinterpret_cast<uint4*>(reinterpret_cast<int16_t*>(out) + (row_ok ? row : rbase) * (int64_t)C)[vi] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                } else {
                    uint32_t lo, hi, w0, w1;
                    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(r[3]), "r"(r[2]), "r"(0));
                    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(w0) : "r"(r[1]), "r"(r[0]), "r"(hi));
                    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(lo) : "r"(r[7]), "r"(r[6]), "r"(0));
                    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(w1) : "r"(r[5]), "r"(r[4]), "r"(lo));
                    dst[vi] = make_uint2(w0, w1);
                }
            }
        }
        if (rnext < rows) {
#pragma unroll
            for (int j = 0; j < NV; ++j) w[j] = wn[j];
        }
    }
}

// out[b, c] = clamp8(RNE(RNE(sum_t x[b, t, c] / L) * m / 2^e)): one thread per (image, 4 channels); L <= 2^16
__global__ void avgpool_requant_kernel(const int8_t* __restrict__ x, int B, int L, int C, ivit_dyadic_t me,
                                       int8_t* __restrict__ out) {
    const int C4 = C >> 2;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * C4) return;
    const int b = i / C4, c4 = i - b * C4;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(x + (int64_t)b * L * C) + c4;
    int32_t s[4] = {0, 0, 0, 0};
    for (int t = 0; t < L; ++t) {
        const uint32_t v = __ldg(src + (int64_t)t * C4);
#pragma unroll
        for (int u = 0; u < 4; ++u) s[u] += (int32_t)(int8_t)(v >> (8 * u));
    }
    uint32_t o = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        int32_t qd = s[u] / L, rem = s[u] - qd * L;
        if (rem < 0) { qd -= 1; rem += L; }                      // floor division
        if (2 * rem > L || (2 * rem == L && (qd & 1))) qd += 1;  // round half to even
        const int32_t q = clamp_bits<8>(requant32(qd, me.m, me.e));
        o |= (uint32_t)(uint8_t)(int8_t)q << (8 * u);
    }
    reinterpret_cast<uint32_t*>(out + (int64_t)b * C)[c4] = o;
}

// int8 -> int16, 16 values per thread: the 8-bit output of PatchMerging's qact2 (swin_quant.py:347) enters the int16 residual stream
__global__ void widen_i8_i16_kernel(const uint4* __restrict__ x, int64_t n16, uint4* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(x + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t o[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            o[2 * k] = g_prmt(w[k], 0u, 0x9180u);
            o[2 * k + 1] = g_prmt(w[k], 0u, 0xB3A2u);
        }
        out[2 * i] = make_uint4(o[0], o[1], o[2], o[3]);
        out[2 * i + 1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
}

}  // namespace ivit

using namespace ivit;

extern "C" {

int ivit_widen_i8_i16(ivit_ctx* ctx, const int8_t* x, int64_t n, int16_t* out, ivit_stream stream) {
    IVIT_REQUIRE(ctx && x && out && n > 0 && n % 16 == 0, "ivit_widen_i8_i16: n must be a positive multiple of 16");
    IVIT_REQUIRE(((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 16) == 0, "ivit_widen_i8_i16: 16-byte alignment");
    const int64_t n16 = n / 16, blocks = (n16 + 255) / 256;
    const int grid = (int)(blocks < (int64_t)ctx->num_sms * 16 ? blocks : (int64_t)ctx->num_sms * 16);
    widen_i8_i16_kernel<<<grid, 256, 0, st(stream)>>>(reinterpret_cast<const uint4*>(x), n16, reinterpret_cast<uint4*>(out));
    IVIT_LAUNCH_OK("widen_i8_i16_kernel");
    return IVIT_OK;
}

int ivit_layernorm_gather_i16_i8(ivit_ctx* ctx, const int16_t* x, int64_t rows_out, int C, int G, const int32_t* rowmap,
                                 int L_out, int L_in, const int32_t* bias_int, const ivit_dyadic_t* me, int8_t* out,
                                 int16_t* xcopy, ivit_stream stream) {
    IVIT_REQUIRE(ctx && x && bias_int && me && out && rows_out > 0, "ivit_layernorm_gather_i16_i8: bad arguments");
    IVIT_REQUIRE(G == 1 || G == 4, "ivit_layernorm_gather_i16_i8: G must be 1 (row permutation) or 4 (2x2 patch merging)");
    IVIT_REQUIRE(G == 1 || (rowmap != nullptr && xcopy == nullptr), "ivit_layernorm_gather_i16_i8: G = 4 needs a row map and has no copy output");
    IVIT_REQUIRE(C % (8 * G) == 0 && C >= 8 * G && C <= LNG_MAXC, "ivit_layernorm_gather_i16_i8: C must be a multiple of %d, <= %d", 8 * G, LNG_MAXC);
    IVIT_REQUIRE(L_out > 0 && L_in > 0 && rows_out % L_out == 0 && rows_out < (1LL << 31), "ivit_layernorm_gather_i16_i8: rows_out must be a multiple of L_out");
    IVIT_REQUIRE(((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 8) == 0 && ((uintptr_t)xcopy % 16) == 0, "ivit_layernorm_gather_i16_i8: 16-byte alignment");
    const int nvec = C / 8;
    // narrowest lane group that keeps <= 6 vectors per lane: more rows per warp, less per-row redundancy
    int lpr = 4;
    while (lpr < 32 && (nvec + lpr - 1) / lpr > 6) lpr *= 2;
    const int nv = (nvec + lpr - 1) / lpr;
    IVIT_REQUIRE(nv <= 8, "ivit_layernorm_gather_i16_i8: C too wide");
    const bool full = nv * lpr == nvec;
    const int rpb = 8 * (32 / lpr);
    const int64_t want = (rows_out + rpb - 1) / rpb;
    const int grid = (int)(want < (int64_t)ctx->num_sms * 2 ? want : (int64_t)ctx->num_sms * 2);
#define LG_K(NV, LPR, FULLV, GV) IVIT_CUDA_OK(launch_k(layernorm_gather_kernel<NV, LPR, FULLV, GV, false>, dim3(grid), dim3(256), 0, st(stream), x, rows_out, C, rowmap, L_out, L_in, bias_int, me, out, xcopy, ivit_dyadic_t{0, 0}))
#define LG_G(NV, LPR, FULLV) do { if (G == 1) LG_K(NV, LPR, FULLV, 1); else LG_K(NV, LPR, FULLV, 4); } while (0)
#define LG_F(NV, LPR) do { if (full) LG_G(NV, LPR, true); else LG_G(NV, LPR, false); } while (0)
#define LG_L(NV) do { switch (lpr) { case 4: LG_F(NV, 4); break; case 8: LG_F(NV, 8); break; case 16: LG_F(NV, 16); break; default: LG_F(NV, 32); break; } } while (0)
    switch (nv) { case 1: LG_L(1); break; case 2: LG_L(2); break; case 3: LG_L(3); break; case 4: LG_L(4); break;
                  case 5: LG_L(5); break; case 6: LG_L(6); break;
                  case 7: LG_F(7, 32); break; default: LG_F(8, 32); break; }      // C > 1536: one row per warp
#undef LG_L
#undef LG_F
#undef LG_G
#undef LG_K
    IVIT_LAUNCH_OK("layernorm_gather_kernel");
    return IVIT_OK;
}

int ivit_layernorm_i8_i16x2(ivit_ctx* ctx, const int8_t* x, int64_t rows, int C, const int32_t* bias_int,
                            const ivit_dyadic_t* me, ivit_dyadic_t me2, int16_t* out, ivit_stream stream) {
    IVIT_REQUIRE(ctx && x && bias_int && me && out && rows > 0 && rows < (1LL << 31), "ivit_layernorm_i8_i16x2: bad arguments");
    IVIT_REQUIRE(C % 8 == 0 && C >= 8 && C <= LNG_MAXC, "ivit_layernorm_i8_i16x2: C must be a multiple of 8, <= %d", LNG_MAXC);
    IVIT_REQUIRE(((uintptr_t)x % 8) == 0 && ((uintptr_t)out % 16) == 0, "ivit_layernorm_i8_i16x2: alignment");
    const int nvec = C / 8;
    int lpr = 4;
    while (lpr < 32 && (nvec + lpr - 1) / lpr > 6) lpr *= 2;
    const int nv = (nvec + lpr - 1) / lpr;
    IVIT_REQUIRE(nv <= 6, "ivit_layernorm_i8_i16x2: C too wide");
    const bool full = nv * lpr == nvec;
    const int rpb = 8 * (32 / lpr);
    const int64_t want = (rows + rpb - 1) / rpb;
    const int grid = (int)(want < (int64_t)ctx->num_sms * 2 ? want : (int64_t)ctx->num_sms * 2);
    const int L = 1;
#define LQ_K(NV, LPR, FULLV) layernorm_gather_kernel<NV, LPR, FULLV, 1, true><<<grid, 256, 0, st(stream)>>>( \
        reinterpret_cast<const int16_t*>(x), rows, C, nullptr, L, L, bias_int, me, reinterpret_cast<int8_t*>(out), nullptr, me2)
#define LQ_F(NV, LPR) do { if (full) LQ_K(NV, LPR, true); else LQ_K(NV, LPR, false); } while (0)
#define LQ_L(NV) do { switch (lpr) { case 4: LQ_F(NV, 4); break; case 8: LQ_F(NV, 8); break; case 16: LQ_F(NV, 16); break; default: LQ_F(NV, 32); break; } } while (0)
    switch (nv) { case 1: LQ_L(1); break; case 2: LQ_L(2); break; case 3: LQ_L(3); break; case 4: LQ_L(4); break;
                  case 5: LQ_L(5); break; default: LQ_L(6); break; }
#undef LQ_L
#undef LQ_F
#undef LQ_K
    IVIT_LAUNCH_OK("layernorm_gather_kernel (int8 -> int16 x2)");
    return IVIT_OK;
}

int ivit_avgpool_requant_i8(ivit_ctx* ctx, const int8_t* x, int B, int L, int C, ivit_dyadic_t me, int8_t* out,
                            ivit_stream stream) {
    IVIT_REQUIRE(ctx && x && out && B > 0 && L > 0 && L <= 65536 && C > 0 && C % 4 == 0, "ivit_avgpool_requant_i8: bad arguments (C % 4 == 0)");
    IVIT_REQUIRE(((uintptr_t)x % 4) == 0 && ((uintptr_t)out % 4) == 0, "ivit_avgpool_requant_i8: 4-byte alignment");
    const int n = B * (C / 4);
    avgpool_requant_kernel<<<(n + 127) / 128, 128, 0, st(stream)>>>(x, B, L, C, me, out);
    IVIT_LAUNCH_OK("avgpool_requant_kernel");
    return IVIT_OK;
}

}  // extern "C"
