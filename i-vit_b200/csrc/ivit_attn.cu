// Fused integer attention and the generic batched integer matmul (sm_100a).
//
// ivit_attention_i8:  one CTA per (sequence, head).  S = Q K^T -> dyadic requant to int8
// (qact_attn1) -> Shiftmax (IntSoftmax, LUT over d = q - max in [-255, 0]) -> P V -> dyadic
// requant to int8 (attn.qact2).  Scores and probabilities never leave the SM: S lives in the
// accumulator registers of the integer tensor-core MMA (mma.sync m16n8k32 s8), P is re-used
// in place as the A operand of the second MMA (16-bit P split into an unsigned high and low
// byte plane, two u8 x s8 MMAs, recombined (hi << 8) + lo).
// Reference call order: vit_quant.py:59-83 (DeiT), swin_quant.py:121-164 (Swin).
// This mma.sync kernel is the GENERAL path (Swin's relative-position-bias QuantAct and shifted-window mask, 8-bit
// probabilities, head_dim 32, slow-form requants); the DeiT shapes take the tcgen05 kernel of ivit_attn_tc.cu.
//
// ivit_bmm_i32: QuantMatMul.forward's contraction (quant_modules.py:223-228) for the
// operator-level API (raw int32 result, strided batched views, int8 or int16 A).
#include <stdlib.h>

#include "ivit_common.cuh"
#include "ivit_internal.h"

namespace ivit {

IVIT_DEVINL void mma_s8s8(int32_t (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
IVIT_DEVINL void mma_u8s8(int32_t (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct AttnArgs {
    int n_seq, n_tok, H;
    ivit_dyadic_t me_s;
    int32_t x0;
    float inv_x0;
    int n;
    int p_shift;                  // 31 - p_bits + 1  (16 for 16-bit P, 24 for 8-bit P)
    int p_bits;
    ivit_dyadic_t me_o;
    const int8_t* relbias;        // [H, n_tok, n_tok] or null
    ivit_dyadic_t me_s2, me_b;
    const int32_t* mask;          // [n_win, n_tok, n_tok] or null
    int n_win;
    long long half_s, half_o;     // 2^(e-1) of me_s / me_o (FAST form)
    int sum32;                    // E(0) = |x0| << n < 2^23: per-thread partial sums of the exponentials fit 32 bits
};

constexpr int ATT_WARPS = 8;

// D: head dim (32 | 64).  KT: number of 32-token chunks (n_tok <= 32*KT).
// SWIN: relative-position-bias QuantAct and/or shifted-window mask present.  P16: 16-bit probabilities (two byte planes).
// FAST: both scalar requants qualify for the branch-free form (checked on the host, which knows (m, e) by value).
template <int D, int KT, bool SWIN, bool P16, bool FAST>
__global__ void __launch_bounds__(ATT_WARPS * 32, 2)
attention_kernel(const int8_t* __restrict__ qkv, const AttnArgs p, int8_t* __restrict__ out) {
    constexpr int NT = KT * 4;             // 8-column score tiles
    constexpr int KSTR = D + 16;           // bytes; (KSTR/4) mod 32 spreads the 8 fragment rows over distinct banks
    constexpr int VSTR = KT * 32 + 16;
    constexpr int KK = D / 32;             // k-steps of Q K^T
    constexpr int ND = D / 8;              // 8-column output tiles
    // dynamic shared memory: K tile | V^T tile | exponent LUT | per-thread packed scores
    extern __shared__ __align__(16) uint8_t att_smem[];
    int8_t* sK = reinterpret_cast<int8_t*>(att_smem);
    int8_t* sVt = sK + KT * 32 * KSTR;
    int32_t* sE = reinterpret_cast<int32_t*>(sVt + D * VSTR);   // [257]; [256] = saturated value (d <= n*x0), used by masked entries
    uint32_t* sSV = reinterpret_cast<uint32_t*>(sE + 260);       // [ATT_WARPS * NT * 32]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q4 = lane & 3;
    const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
    const int n_tok = p.n_tok;
    const long long ld = 3LL * p.H * D;
    const int8_t* qb = qkv + (long long)b * n_tok * ld + h * D;
    const int8_t* kb = qb + p.H * D;
    const int8_t* vb = qb + 2 * p.H * D;

    // ---- K tile -> smem (zero padded) ----
    for (int idx = tid; idx < KT * 32 * (D / 16); idx += ATT_WARPS * 32) {
        const int j = idx / (D / 16), v = idx % (D / 16);
        uint4 val = make_uint4(0, 0, 0, 0);
        if (j < n_tok) val = *reinterpret_cast<const uint4*>(kb + (long long)j * ld + 16 * v);
        *reinterpret_cast<uint4*>(sK + j * KSTR + 16 * v) = val;
    }
    // ---- V -> smem transposed [d][token slot], token slots permuted inside each 32-chunk so that
    //      the P accumulator fragment of the first MMA is directly the A fragment of the second:
    //      token jj = 8a + 2c + b (a<4, c<4, b<2)  ->  slot = (a<2 ? 0 : 16) + 4c + 2(a&1) + b
    for (int idx = tid; idx < KT * 16 * (D / 4); idx += ATT_WARPS * 32) {
        const int jp = idx / (D / 4), db = idx % (D / 4);     // token pair, group of 4 d
        const int j = 2 * jp;
        uint32_t r0 = 0, r1 = 0;
        if (j < n_tok) r0 = *reinterpret_cast<const uint32_t*>(vb + (long long)j * ld + 4 * db);
        if (j + 1 < n_tok) r1 = *reinterpret_cast<const uint32_t*>(vb + (long long)(j + 1) * ld + 4 * db);
        const int jj = j & 31, chunk = j >> 5;
        const int a = jj >> 3, c = (jj & 7) >> 1;
        const int slot = ((a < 2) ? 0 : 16) + 4 * c + 2 * (a & 1);
#pragma unroll
        for (int dd = 0; dd < 4; ++dd) {
            const uint16_t pr = (uint16_t)(((r0 >> (8 * dd)) & 0xff) | (((r1 >> (8 * dd)) & 0xff) << 8));
            *reinterpret_cast<uint16_t*>(sVt + (4 * db + dd) * VSTR + chunk * 32 + slot) = pr;
        }
    }
    // ---- Shiftmax exponent LUT: sE[k] = int_exp_shift(-k), k = max - q in [0, 255] ----
    for (int k = tid; k < 257; k += ATT_WARPS * 32)
        sE[k] = (int32_t)shiftexp(k < 256 ? -k : -(1 << 20), p.x0, p.inv_x0, p.n);
    __syncthreads();

    const int n_row_tiles = (n_tok + 15) / 16;
    const int nt_full = n_tok >> 3;                   // score tiles with 8 valid columns
    const int nt_used = (n_tok + 7) >> 3;             // tiles that contain at least one valid column
    const int col_lim = n_tok - 2 * q4;               // in the boundary tile: column 8t + (c&1) + 2*q4 is padding iff 8t + (c&1) >= col_lim
    const uint32_t e_sat = (uint32_t)sE[256];
    uint32_t* my_sv = sSV + (warp * NT) * 32 + lane;  // this thread's packed scores: my_sv[t * 32]
    // FAST (host-checked): both scalar requants have 32 <= e <= 62 and no reachable exact tie, so
    //   RNE(z*m/2^e) == hi32(z*m + 2^(e-1)) >> (e-32)   -- three instructions, constants in registers
    const int32_t m_s = p.me_s.m, m_o = p.me_o.m;
    const int sh_s = p.me_s.e - 32, sh_o = p.me_o.e - 32;
    const long long half_s = p.half_s, half_o = p.half_o;
    auto rq_scores = [&](int32_t z) -> int32_t {
        if (FAST) return (int32_t)(((long long)z * (long long)m_s + half_s) >> 32) >> sh_s;
        return requant32_general(z, p.me_s.m, p.me_s.e);
    };
    auto rq_out = [&](int32_t z) -> int32_t {
        if (FAST) return (int32_t)(((long long)z * (long long)m_o + half_o) >> 32) >> sh_o;
        return requant32_general(z, p.me_o.m, p.me_o.e);
    };
    // NOTE on code size: the loops over score tiles / key chunks are deliberately NOT unrolled (per-thread state
    // lives in shared memory) and the Swin / slow-requant variants are separate template instantiations.  The
    // fully unrolled, all-in-one version was ~110 KB of SASS and spent 60 % of its cycles in instruction-fetch
    // stalls (profiles/).
    for (int rt = warp; rt < n_row_tiles; rt += ATT_WARPS) {
        const int r0 = rt * 16 + g, r1 = r0 + 8;
        // ---- Q fragments straight from global ----
        uint32_t aq[KK][4];
#pragma unroll
        for (int kk = 0; kk < KK; ++kk) {
            aq[kk][0] = (r0 < n_tok) ? __ldg(reinterpret_cast<const uint32_t*>(qb + (long long)r0 * ld + 32 * kk + 4 * q4)) : 0u;
            aq[kk][1] = (r1 < n_tok) ? __ldg(reinterpret_cast<const uint32_t*>(qb + (long long)r1 * ld + 32 * kk + 4 * q4)) : 0u;
            aq[kk][2] = (r0 < n_tok) ? __ldg(reinterpret_cast<const uint32_t*>(qb + (long long)r0 * ld + 32 * kk + 16 + 4 * q4)) : 0u;
            aq[kk][3] = (r1 < n_tok) ? __ldg(reinterpret_cast<const uint32_t*>(qb + (long long)r1 * ld + 32 * kk + 16 + 4 * q4)) : 0u;
        }
        // ---- S = Q K^T, requantised tile by tile to int8 (qact_attn1 [+ rel-pos bias, mask]) and packed four per
        //      word: {row r0: col 2q, 2q+1 ; row r1: col 2q, 2q+1}; padding columns / masked entries -> -128 ----
        uint32_t mxw = 0x80808080u;                   // per-byte running max of the packed scores (bytes 0,1: row r0; 2,3: row r1)
        uint32_t masked_bits = 0u;                    // Swin only (NT <= 8): 4 bits per tile
        auto score_tile = [&](int t, bool boundary) {
            int32_t acc[4] = {0, 0, 0, 0};
#pragma unroll
            for (int kk = 0; kk < KK; ++kk) {
                const uint32_t b0 = *reinterpret_cast<const uint32_t*>(sK + (8 * t + g) * KSTR + 32 * kk + 4 * q4);
                const uint32_t b1 = *reinterpret_cast<const uint32_t*>(sK + (8 * t + g) * KSTR + 32 * kk + 16 + 4 * q4);
                mma_s8s8(acc, aq[kk], b0, b1);
            }
            int32_t v[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                v[c] = rq_scores(acc[c]);                                    // clamped to int8 by the saturating pack below
                const bool pad = boundary && ((8 * t + (c & 1)) >= col_lim);
                if (SWIN) {
                    v[c] = clamp_bits<8>(v[c]);
                    const int col = 8 * t + 2 * q4 + (c & 1);
                    const int row = (c < 2) ? r0 : r1;
                    if (!pad && row < n_tok) {
                        if (p.relbias != nullptr) {
                            const int32_t bq = (int32_t)p.relbias[((long long)h * n_tok + row) * n_tok + col];
                            v[c] = clamp_bits<8>(sat_i64_to_i32((long long)requant32_general(v[c], p.me_s2.m, p.me_s2.e) +
                                                                (long long)requant32_general(bq, p.me_b.m, p.me_b.e)));
                        }
                        if (p.mask != nullptr) {
                            // a masked entry (addend RNE(-100/s)) is always past the Shiftmax saturation point for
                            // s < 0.35: remember it, its exponential is E(n*x0)
                            if (p.mask[((long long)(b % p.n_win) * n_tok + row) * n_tok + col] != 0) {
                                masked_bits |= 1u << (4 * (t & 7) + c);
                                v[c] = -128;
                            }
                        }
                    }
                }
                if (pad) v[c] = -128;                                        // never raises the row max
            }
            uint32_t hi2, w;                                                 // saturate to int8 and pack: v0 | v1<<8 | v2<<16 | v3<<24
            asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi2) : "r"(v[3]), "r"(v[2]), "r"(0));
            asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(w) : "r"(v[1]), "r"(v[0]), "r"(hi2));
            mxw = __vmaxs4(mxw, w);
            my_sv[t * 32] = w;
        };
#pragma unroll 1
        for (int t = 0; t < nt_full; ++t) score_tile(t, false);
        if (nt_used > nt_full) score_tile(nt_full, true);
        for (int t = nt_used; t < NT; ++t) my_sv[t * 32] = 0x80808080u;      // unused tiles: valid LUT index, V rows are zero
        mxw = __vmaxs4(mxw, __shfl_xor_sync(0xffffffffu, mxw, 1));
        mxw = __vmaxs4(mxw, __shfl_xor_sync(0xffffffffu, mxw, 2));
        const int32_t mx0 = max((int32_t)(int8_t)(mxw & 0xff), (int32_t)(int8_t)((mxw >> 8) & 0xff));
        const int32_t mx1 = max((int32_t)(int8_t)((mxw >> 16) & 0xff), (int32_t)mxw >> 24);
        // ---- exponentials (LUT over max - q): E for the four entries of packed word w of tile t ----
        const int32_t* sE0 = sE + mx0;
        const int32_t* sE1 = sE + mx1;
        auto expo4 = [&](int t, uint32_t w, uint32_t (&E)[4], bool boundary) {
            E[0] = (uint32_t)sE0[-(int32_t)(int8_t)(w & 0xff)];
            E[1] = (uint32_t)sE0[-(int32_t)(int8_t)((w >> 8) & 0xff)];
            E[2] = (uint32_t)sE1[-(int32_t)(int8_t)((w >> 16) & 0xff)];
            E[3] = (uint32_t)sE1[-((int32_t)w >> 24)];
            if (SWIN) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if ((masked_bits >> (4 * (t & 7) + c)) & 1u) E[c] = e_sat;
            }
            if (boundary) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if ((8 * t + (c & 1)) >= col_lim) E[c] = 0u;
            }
        };
        unsigned long long sum0 = 0, sum1 = 0;
        if (p.sum32) {
            // E <= E(0) = |x0| << n < 2^23 (host-checked): a thread's <= 2*NT terms fit 32 bits
            uint32_t s0 = 0, s1 = 0;
#pragma unroll 2
            for (int t = 0; t < nt_full; ++t) {
                uint32_t E[4];
                expo4(t, my_sv[t * 32], E, false);
                s0 += E[0] + E[1];
                s1 += E[2] + E[3];
            }
            sum0 = s0; sum1 = s1;
        } else {
#pragma unroll 1
            for (int t = 0; t < nt_full; ++t) {
                uint32_t E[4];
                expo4(t, my_sv[t * 32], E, false);
                sum0 += (unsigned long long)E[0] + E[1];
                sum1 += (unsigned long long)E[2] + E[3];
            }
        }
        if (nt_used > nt_full) {
            uint32_t E[4];
            expo4(nt_full, my_sv[nt_full * 32], E, true);
            sum0 += (unsigned long long)E[0] + E[1];
            sum1 += (unsigned long long)E[2] + E[3];
        }
        sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
        sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
        const uint32_t S0 = sum0 > 2147483647ULL ? 2147483647u : (uint32_t)sum0;   // clamp_max_(2**31-1)
        const uint32_t S1 = sum1 > 2147483647ULL ? 2147483647u : (uint32_t)sum1;
        const uint32_t F0 = 2147483647u / (S0 ? S0 : 1u);
        const uint32_t F1 = 2147483647u / (S1 ? S1 : 1u);
        // P = (E*F) >> p_shift == umulhi(E, F << (32 - p_shift)) when F << (32 - p_shift) fits 32 bits
        // (F <= (2^31-1)/E_max and E_max >= |x0| * 2^(n-2) >= 2^13 for n = 15: F < 2^18 ... checked, else two-step form)
        const int lsh = 32 - p.p_shift;
        const bool hi_form = (F0 >> p.p_shift) == 0u && (F1 >> p.p_shift) == 0u;
        const uint32_t F0s = F0 << lsh, F1s = F1 << lsh;

        // ---- O = P V with P = (E * F) >> p_shift (E*F <= S*F < 2^31: 32-bit product), hi/lo byte planes.
        //      No padding checks here: V rows of padding tokens are zero in shared memory. ----
        int32_t ohi[ND][4], olo[ND][4];
#pragma unroll
        for (int nd = 0; nd < ND; ++nd) {
            ohi[nd][0] = ohi[nd][1] = ohi[nd][2] = ohi[nd][3] = 0;
            olo[nd][0] = olo[nd][1] = olo[nd][2] = olo[nd][3] = 0;
        }
        const int kc_used = (n_tok + 31) >> 5;
#pragma unroll 1
        for (int kc = 0; kc < kc_used; ++kc) {
            // tiles 4kc .. 4kc+3; fragment register r: rows (r&1 ? r1 : r0), tiles 4kc + (r>>1)*2 + {0,1}
            uint32_t P[4][4];
#pragma unroll
            for (int tt = 0; tt < 4; ++tt) {
                uint32_t E[4];
                expo4(4 * kc + tt, my_sv[(4 * kc + tt) * 32], E, false);
                if (hi_form) {                                               // warp-divergence free in practice (same for all rows)
                    P[tt][0] = __umulhi(E[0], F0s); P[tt][1] = __umulhi(E[1], F0s);
                    P[tt][2] = __umulhi(E[2], F1s); P[tt][3] = __umulhi(E[3], F1s);
                } else {
                    P[tt][0] = (E[0] * F0) >> p.p_shift; P[tt][1] = (E[1] * F0) >> p.p_shift;
                    P[tt][2] = (E[2] * F1) >> p.p_shift; P[tt][3] = (E[3] * F1) >> p.p_shift;
                }
            }
            uint32_t alo[4], ahi[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int ta = (r >> 1) * 2, cb = (r & 1) * 2;
                alo[r] = __byte_perm(__byte_perm(P[ta][cb], P[ta][cb + 1], 0x0040), __byte_perm(P[ta + 1][cb], P[ta + 1][cb + 1], 0x0040), 0x5410);
                if (P16) ahi[r] = __byte_perm(__byte_perm(P[ta][cb], P[ta][cb + 1], 0x0051), __byte_perm(P[ta + 1][cb], P[ta + 1][cb + 1], 0x0051), 0x5410);
            }
#pragma unroll
            for (int nd = 0; nd < ND; ++nd) {
                const uint32_t b0 = *reinterpret_cast<const uint32_t*>(sVt + (8 * nd + g) * VSTR + 32 * kc + 4 * q4);
                const uint32_t b1 = *reinterpret_cast<const uint32_t*>(sVt + (8 * nd + g) * VSTR + 32 * kc + 16 + 4 * q4);
                mma_u8s8(olo[nd], alo, b0, b1);
                if (P16) mma_u8s8(ohi[nd], ahi, b0, b1);
            }
        }
        // ---- requant (attn.qact2) and store ----
#pragma unroll
        for (int nd = 0; nd < ND; ++nd) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int row = half ? r1 : r0;
                if (row < n_tok) {
                    const int32_t v0 = P16 ? (ohi[nd][2 * half] << 8) + olo[nd][2 * half] : olo[nd][2 * half];
                    const int32_t v1 = P16 ? (ohi[nd][2 * half + 1] << 8) + olo[nd][2 * half + 1] : olo[nd][2 * half + 1];
                    const int32_t o0 = clamp_bits<8>(rq_out(v0));
                    const int32_t o1 = clamp_bits<8>(rq_out(v1));
                    int8_t* dst = out + ((long long)b * n_tok + row) * (long long)(p.H * D) + h * D + 8 * nd + 2 * q4;
                    *reinterpret_cast<uint16_t*>(dst) = (uint16_t)((o0 & 0xff) | ((o1 & 0xff) << 8));
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// generic strided batched integer matmul, raw int32 result (operator-level QuantMatMul)
// ------------------------------------------------------------------------------------
// 64 x 64 output tile per 256-thread block, 4 x 4 outputs per thread (rows 4*ty + i, columns tx + 16*j), K in steps of
// 32 through shared memory as int32 with a 36-word row pitch: 16-byte loads of four k values, 64 IMADs per eight loads
// (the first version, one output per thread from 16 x 16 tiles, took 26 % of an operator-level DeiT forward).
template <typename TA>
__global__ void __launch_bounds__(256, 3)
bmm_i32_kernel(const TA* __restrict__ A, long long lda, long long sa,
               const int8_t* __restrict__ B, long long ldb, long long sb, int trans_b,
               int M, int N, int K, int32_t* __restrict__ C, long long ldc, long long sc) {
    constexpr int TP = 36;                                    // row pitch in words: 16-byte aligned, conflict-free for 8 lanes
    __shared__ __align__(16) int32_t tA[64 * TP], tB[64 * TP];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long bi = blockIdx.z;
    const TA* Ab = A + bi * sa;
    const int8_t* Bb = B + bi * sb;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    int32_t acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0;
    for (int k0 = 0; k0 < K; k0 += 32) {
        // stage: 64 x 32 elements of A and of B ([n][k]) each, 8 per thread; k fastest for row-major operands
#pragma unroll 2
        for (int e = 0; e < 8; ++e) {
            const int idx = tid + 256 * e, r = idx >> 5, k = idx & 31;
            tA[r * TP + k] = (m0 + r < M && k0 + k < K) ? (int32_t)Ab[(long long)(m0 + r) * lda + k0 + k] : 0;
            int32_t bv = 0;
            if (trans_b) {                                    // B is [N, K] row-major
                if (n0 + r < N && k0 + k < K) bv = (int32_t)Bb[(long long)(n0 + r) * ldb + k0 + k];
                tB[r * TP + k] = bv;
            } else {                                          // B is [K, N] row-major: n fastest in memory
                const int kk = idx >> 6, nn = idx & 63;
                if (n0 + nn < N && k0 + kk < K) bv = (int32_t)Bb[(long long)(k0 + kk) * ldb + n0 + nn];
                tB[nn * TP + kk] = bv;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
            int4 av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = *reinterpret_cast<const int4*>(&tA[(4 * ty + i) * TP + k]);
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = *reinterpret_cast<const int4*>(&tB[(tx + 16 * j) * TP + k]);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    acc[i][j] += av[i].x * bv[j].x + av[i].y * bv[j].y + av[i].z * bv[j].z + av[i].w * bv[j].w;
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = m0 + 4 * ty + i;
        if (row < M) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int col = n0 + tx + 16 * j;
                if (col < N) C[bi * sc + (long long)row * ldc + col] = acc[i][j];
            }
        }
    }
}

template <int D, int KT, bool SWIN, bool P16, bool FAST>
static int launch_attention(int grid, const int8_t* qkv, const AttnArgs& a, int8_t* out, cudaStream_t s) {
    constexpr int SMEM = KT * 32 * (D + 16) + D * (KT * 32 + 16) + 260 * 4 + ATT_WARPS * KT * 4 * 32 * 4;
    auto kern = attention_kernel<D, KT, SWIN, P16, FAST>;
    static PerDevice attr_set;
    int dev = 0;
    IVIT_CUDA_OK(cudaGetDevice(&dev));
    if (!attr_set[dev]) {
        IVIT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_set[dev] = 1;
    }
    kern<<<grid, ATT_WARPS * 32, SMEM, s>>>(qkv, a, out);
    IVIT_LAUNCH_OK("attention_kernel");
    return IVIT_OK;
}

// branch-free requant is exact iff 32 <= e <= 62 and no exact tie is reachable for |z| < 2^zbits
static bool fast_dyadic(ivit_dyadic_t d, int zbits) {
    if (d.m == 0 || d.e < 32 || d.e > 62) return false;
    const int tz = __builtin_ctz((unsigned)d.m);
    return d.e - 1 - tz > zbits;
}

template <int D, int KT>
static int dispatch_attention(int grid, const int8_t* qkv, const AttnArgs& a, int8_t* out, cudaStream_t s, bool swin, bool p16, bool fast) {
    // instantiations actually used by the models: DeiT (no bias/mask, 16-bit P) and Swin (bias/mask, 8-bit P)
    if (!swin && p16) return fast ? launch_attention<D, KT, false, true, true>(grid, qkv, a, out, s)
                                  : launch_attention<D, KT, false, true, false>(grid, qkv, a, out, s);
    if (!swin && !p16) return fast ? launch_attention<D, KT, false, false, true>(grid, qkv, a, out, s)
                                   : launch_attention<D, KT, false, false, false>(grid, qkv, a, out, s);
    if (p16) return launch_attention<D, KT, true, true, false>(grid, qkv, a, out, s);
    return fast ? launch_attention<D, KT, true, false, true>(grid, qkv, a, out, s)
                : launch_attention<D, KT, true, false, false>(grid, qkv, a, out, s);
}

}  // namespace ivit

using namespace ivit;

extern "C" int ivit_attention_i8(ivit_ctx* ctx, const int8_t* qkv, const ivit_attn_params* ap, int8_t* out,
                                 ivit_stream stream) {
    IVIT_REQUIRE(ctx && qkv && ap && out, "ivit_attention_i8: null pointer");
    IVIT_REQUIRE(ap->n_seq > 0 && ap->n_tok > 0 && ap->n_heads > 0, "ivit_attention_i8: bad shape");
    IVIT_REQUIRE(ap->head_dim == 64 || ap->head_dim == 32, "ivit_attention_i8: head_dim must be 32 or 64");
    IVIT_REQUIRE(ap->p_bits == 16 || ap->p_bits == 8, "ivit_attention_i8: p_bits must be 8 or 16");
    IVIT_REQUIRE(ap->n >= 1 && ap->n <= 16, "ivit_attention_i8: n must be in [1,16] (int32 exponent LUT)");
    IVIT_REQUIRE(((uintptr_t)qkv % 16) == 0 && ((uintptr_t)out % 2) == 0, "ivit_attention_i8: qkv must be 16-byte aligned");
    if (!(ap->x0 <= -1 && ap->x0 >= -65535))
        return fail(IVIT_ENOTSUP, "ivit_attention_i8: x0=%d outside [-65535, -1]", ap->x0);
    if (ap->mask) IVIT_REQUIRE(ap->n_win > 0 && ap->n_seq % ap->n_win == 0, "ivit_attention_i8: n_seq must be a multiple of n_win");
    AttnArgs a;
    a.n_seq = ap->n_seq; a.n_tok = ap->n_tok; a.H = ap->n_heads;
    a.me_s = ap->me_s; a.x0 = ap->x0; a.inv_x0 = 1.0f / (float)ap->x0; a.n = ap->n;
    a.p_bits = ap->p_bits; a.p_shift = 31 - ap->p_bits + 1;
    a.me_o = ap->me_o; a.relbias = ap->relbias; a.me_s2 = ap->me_s2; a.me_b = ap->me_b;
    a.mask = ap->mask; a.n_win = ap->n_win > 0 ? ap->n_win : 1;
    const int grid = ap->n_seq * ap->n_heads;
    cudaStream_t s = st(stream);
    if (ap->n_tok > 224) return fail(IVIT_ENOTSUP, "ivit_attention_i8: n_tok=%d > 224 not supported", ap->n_tok);
    const bool swin = (ap->relbias != nullptr) || (ap->mask != nullptr);
    const bool p16 = ap->p_bits > 8;
    // |Q.K| <= 64 * 128 * 128 = 2^20 ; |P.V| <= 2^15 * 128 = 2^22
    const bool fast = fast_dyadic(ap->me_s, 21) && fast_dyadic(ap->me_o, 23);
    a.sum32 = (((long long)(-ap->x0)) << ap->n) < (1LL << 23) ? 1 : 0;
    a.half_s = (ap->me_s.e >= 1 && ap->me_s.e <= 62) ? (1LL << (ap->me_s.e - 1)) : 0;
    a.half_o = (ap->me_o.e >= 1 && ap->me_o.e <= 62) ? (1LL << (ap->me_o.e - 1)) : 0;
    // DeiT path on the tcgen05 tensor cores when the preconditions hold: the pipelined one-CTA-per-SM kernel
    // (ivit_attn_pipe.cu; IVIT_ATTN_PIPE=0 disables) or round 1's two-CTA kernel (ivit_attn_tc.cu; IVIT_ATTN_TC=0 disables)
    {
        static const char* tc_env = getenv("IVIT_ATTN_TC");
        static const char* pipe_env = getenv("IVIT_ATTN_PIPE");
        const long long e0 = ((long long)(-ap->x0)) << ap->n;              // largest exponential, E(0)
        const bool common = ap->head_dim == 64 && !swin && p16 && fast && ap->n_tok <= 224 && ap->n == 15 &&
                            ctx->encode_tiled != nullptr;
        if (common && !(pipe_env && pipe_env[0] == '0') && ap->n_tok >= 49 && e0 < (1LL << 31))
            return launch_attention_pipe(ctx, qkv, ap, a.half_s, a.half_o, out, s);
        const bool tc = !(tc_env && tc_env[0] == '0') && common && e0 >= (1LL << 15) && e0 < (1LL << 23);
        if (tc) return launch_attention_tc(ctx, qkv, ap, a.half_s, a.half_o, out, s);
    }
    if (swin && ap->n_tok > 64) return fail(IVIT_ENOTSUP, "ivit_attention_i8: bias/mask path supports n_tok <= 64 (window attention)");
    int rc;
    if (ap->head_dim == 64) rc = (ap->n_tok <= 64) ? dispatch_attention<64, 2>(grid, qkv, a, out, s, swin, p16, fast)
                                                   : dispatch_attention<64, 7>(grid, qkv, a, out, s, swin, p16, fast);
    else rc = (ap->n_tok <= 64) ? dispatch_attention<32, 2>(grid, qkv, a, out, s, swin, p16, fast)
                                : dispatch_attention<32, 7>(grid, qkv, a, out, s, swin, p16, fast);
    return rc;
}

extern "C" int ivit_bmm_i32(ivit_ctx* ctx, const void* A, int a_dtype, int64_t lda, int64_t sa, const int8_t* B,
                            int64_t ldb, int64_t sb, int trans_b, int64_t batch, int M, int N, int K,
                            int32_t* C, int64_t ldc, int64_t sc, ivit_stream stream) {
    IVIT_REQUIRE(ctx && A && B && C && batch > 0 && M > 0 && N > 0 && K > 0, "ivit_bmm_i32: bad arguments");
    IVIT_REQUIRE(a_dtype == IVIT_I8 || a_dtype == IVIT_I16, "ivit_bmm_i32: a_dtype must be I8 or I16");
    // the batch index rides in gridDim.z (<= 65535): larger batches (Swin-B bs=256 stage 1: 256*64*4 windows x heads) are
    // launched in z-chunks
    for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
        const int64_t nb = batch - b0 < 65535 ? batch - b0 : 65535;
        dim3 grid((N + 63) / 64, (M + 63) / 64, (unsigned)nb);
        const int es = a_dtype == IVIT_I8 ? 1 : 2;
        const void* Ab = (const char*)A + b0 * sa * es;
        const int8_t* Bb = B + b0 * sb;
        int32_t* Cb = C + b0 * sc;
        if (a_dtype == IVIT_I8)
            bmm_i32_kernel<int8_t><<<grid, 256, 0, st(stream)>>>((const int8_t*)Ab, lda, sa, Bb, ldb, sb, trans_b, M, N, K, Cb, ldc, sc);
        else
            bmm_i32_kernel<int16_t><<<grid, 256, 0, st(stream)>>>((const int16_t*)Ab, lda, sa, Bb, ldb, sb, trans_b, M, N, K, Cb, ldc, sc);
    }
    IVIT_LAUNCH_OK("bmm_i32_kernel");
    return IVIT_OK;
}
