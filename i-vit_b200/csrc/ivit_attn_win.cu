// Fused Swin window attention on the 5th-generation tensor cores (tcgen05, sm_100a): 7 x 7 = 49-token windows,
// head_dim 32, 8-bit probabilities, relative-position bias and shifted-window mask.
//
//   S = Q K^T -> qact_attn1 (dyadic, int8) -> qact2 with the relative-position bias as identity (swin_quant.py:149:
//   clamp8(RNE(s * m2 / 2^e2) + RNE(bias * mb / 2^eb))) -> mask (-100 on the carrier, :151-155) -> Shiftmax, 8 bit
//   (quant_modules.py:469-497) -> P V -> qact3 (:164).          reference call order: swin_quant.py:121-164
//
// One MMA tile holds TWO windows: window A in rows / keys [0, 49), window B in [64, 113) (64-row TMA boxes straight out
// of the packed qkv activations; the 15 rows behind a window are the next window's tokens and never used).  The score
// tile is 128 x 128 with the two windows on its diagonal blocks; a thread owns one query row and reads only its own
// window's 49 columns from TMEM, so the off-diagonal blocks are never looked at, and the probability tile keeps zeros
// there (written once), which makes P V block-diagonal too.
//
// head_dim 32 rides on the 64-byte-row / 64B-swizzle operand layout that the DeiT kernels use: a TMA box is 64 bytes
// wide = TWO adjacent heads; the score MMA of head hh takes the K = 32 slice at byte offset 32 * hh of each row (the
// same descriptor advance the DeiT kernel uses for its second k-step), and P V runs with N = 64 (both heads' channels;
// the other head's 32 columns are ignored).  Q / K / V are therefore loaded once per (window pair, head pair).
//
// Four CTAs per SM (4 warps, 128 TMEM columns, ~54 KB shared memory each): the MMA hand-overs of one CTA are covered by
// the softmax arithmetic of the other three.  One thread = one query row = 49 scores: row max and row sum are
// thread-local, there is no cross-warp exchange at all.
#include <stdlib.h>

#include "ivit_common.cuh"
#include "ivit_internal.h"
#include "ivit_ptx.cuh"

namespace ivit {

constexpr int WA_N = 49;                   // tokens per window
constexpr int WA_NW = 13;                  // packed score words per row (49 bytes)
constexpr int WA_BSTRIDE = 50;             // int16 per bias row in shared memory (4-byte aligned rows)
constexpr int WA_LUTC = 4;                 // copies of the exponent table

struct WinAttnArgs {
    int n_win, H, C;                       // windows in total, heads, channels (= 32 * H)
    int n_wp, n_hp;                        // window pairs, head pairs
    int32_t m_s, sh_s;                     // qact_attn1: hi32(z*m + half) >> sh
    long long half_s;
    int32_t m_2, e_2;                      // qact2 (scores branch): (a*m + half) >> e, 8 <= e <= 62, no reachable tie
    long long half_2;
    int32_t m_o, sh_o;                     // qact3
    long long half_o;
    int32_t x0;
    float inv_x0;
    int n;
    const int16_t* bias_rq;                // [H][49][49]: RNE(bias * mb / 2^eb)
    const unsigned long long* mask_bits;   // [n_win_img][49] or null: bit j of row i = key j masked
    int n_win_img;
};

constexpr int WA_SQ = 0;                   // [128 rows x 64 B]
constexpr int WA_SK = 8192;
constexpr int WA_SV = 16384;
constexpr int WA_SP = 24576;               // [128 rows x 128 B] probabilities, K-major, 128B swizzle
constexpr int WA_SB = 40960;               // bias of the two heads of the current pair: [2][49][50] int16
constexpr int WA_SE = WA_SB + 2 * WA_N * WA_BSTRIDE * 2 + 8;   // exponent table [256][4] uint32 (16-byte aligned)
constexpr int WA_BAR = ((WA_SE + 256 * WA_LUTC * 4 + 15) / 16) * 16;
constexpr int WA_SMEM = WA_BAR + 64 + 1024;
static_assert(WA_SE % 8 == 0, "table alignment");
static_assert(4 * (WA_SMEM + 1024) <= 228 * 1024, "four CTAs per SM");

__device__ __forceinline__ uint64_t wa_desc_sw64(uint32_t smem_addr) {
    uint64_t d = 0;                         // 64-byte rows, 64B swizzle, 8-row atoms 512 B apart (see ivit_attn_tc.cu)
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}

// E2HI: e_2 >= 32 (the shifted product is read from the high word)
template <bool E2HI>
__global__ void __launch_bounds__(128, 4)
window_attention_kernel(const __grid_constant__ CUtensorMap tmap, const WinAttnArgs p, int8_t* __restrict__ out) {
    extern __shared__ uint8_t wa_smem_raw[];
    const uint32_t base = (ptx::smem_u32(wa_smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = wa_smem_raw + (base - ptx::smem_u32(wa_smem_raw));
    const uint32_t sQ = base + WA_SQ, sK = base + WA_SK, sV = base + WA_SV, sP = base + WA_SP;
    int16_t* sB = reinterpret_cast<int16_t*>(smem + WA_SB);
    uint32_t* sE = reinterpret_cast<uint32_t*>(smem + WA_SE);
    const uint32_t bar = base + WA_BAR;
    const uint32_t qk_full = bar, v_full = bar + 8, s_full = bar + 16, p_ready = bar + 24, o_full = bar + 32, o_done = bar + 40;
    volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + WA_BAR + 48);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int units = p.n_wp * p.n_hp;                       // (head pair, window pair), head pair major
    const int u0 = (int)(((long long)units * blockIdx.x) / gridDim.x);
    const int u1 = (int)(((long long)units * (blockIdx.x + 1)) / gridDim.x);

    if (tid == 0) {
        ptx::prefetch_tensormap(&tmap);
        ptx::mbar_init(qk_full, 1);
        ptx::mbar_init(v_full, 1);
        ptx::mbar_init(s_full, 1);
        ptx::mbar_init(p_ready, 4);
        ptx::mbar_init(o_full, 1);
        ptx::mbar_init(o_done, 4);
        ptx::fence_barrier_init();
    }
    if (warp == 0) {
        ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), 128);
        ptx::tmem_relinquish();
    }
    // probability tile: zero once (off-diagonal window blocks and the 15 padding keys of each window stay zero)
    for (int i = tid; i < 128 * 128 / 16; i += 128) reinterpret_cast<uint4*>(smem + WA_SP)[i] = make_uint4(0, 0, 0, 0);
    // exponent table: sE[k][copy] = int_exp_shift(-k), k = max - q in [0, 255]
    for (int k = tid; k < 256; k += 128) {
        const uint32_t e = (uint32_t)shiftexp(-k, p.x0, p.inv_x0, p.n);
#pragma unroll
        for (int j = 0; j < WA_LUTC; ++j) sE[k * WA_LUTC + j] = e;
    }
    ptx::fence_proxy_async();                                // the zeros are read by the MMA (async proxy)
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const uint32_t idesc_s = ptx::umma_idesc_i8(128, 128, 1, 1);
    const uint32_t idesc_pv = ptx::umma_idesc_i8(128, 64, 0, 1) | (1u << 16);   // A = unsigned P; B (V) N-major
    const uint32_t emin = (uint32_t)(-p.x0);                 // int_exp_shift of a masked entry (t clamped at n * x0)

    const int wsel = warp >> 1;                              // my window of the pair
    const int irow = (warp & 1) * 32 + lane;                 // my row inside the window (valid below 49)
    const int trow = warp * 32 + lane;                       // my row of the tile
    const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint32_t sE_lane = ptx::smem_u32(sE) + 4u * ((uint32_t)lane & (WA_LUTC - 1));
    // my 64-byte half of my P row: 16-byte chunks 4*wsel .. 4*wsel+3, swizzled with the row index
    const uint32_t sP_row = sP + (uint32_t)trow * 128u;

    auto load_qk = [&](int u) {
        const int hp = u / p.n_wp, wp = u - hp * p.n_wp;
        ptx::mbar_arrive_expect_tx(qk_full, 4 * 64 * 64);
        ptx::tma_load_2d(sQ, &tmap, qk_full, hp * 64, (2 * wp) * WA_N);
        ptx::tma_load_2d(sQ + 4096, &tmap, qk_full, hp * 64, (2 * wp + 1) * WA_N);
        ptx::tma_load_2d(sK, &tmap, qk_full, p.C + hp * 64, (2 * wp) * WA_N);
        ptx::tma_load_2d(sK + 4096, &tmap, qk_full, p.C + hp * 64, (2 * wp + 1) * WA_N);
    };
    auto load_v = [&](int u) {
        const int hp = u / p.n_wp, wp = u - hp * p.n_wp;
        ptx::mbar_arrive_expect_tx(v_full, 2 * 64 * 64);
        ptx::tma_load_2d(sV, &tmap, v_full, 2 * p.C + hp * 64, (2 * wp) * WA_N);
        ptx::tma_load_2d(sV + 4096, &tmap, v_full, 2 * p.C + hp * 64, (2 * wp + 1) * WA_N);
    };
    if (tid == 0 && u0 < u1) {
        load_qk(u0);
        load_v(u0);
    }

    int cur_hp = -1;
    uint32_t g = 0, it = 0;                                  // g: (unit, head) steps = barrier phases; it: units
#pragma unroll 1
    for (int u = u0; u < u1; ++u, ++it) {
        const int hp = u / p.n_wp, wp = u - hp * p.n_wp;
        const int nh = (2 * hp + 1 < p.H) ? 2 : 1;           // heads of this pair that exist
        if (hp != cur_hp) {                                  // bias of the (up to) two heads -> shared memory (rare)
            __syncthreads();
            for (int i = tid; i < nh * WA_N * WA_N; i += 128) {
                const int hh = i / (WA_N * WA_N), r = i - hh * WA_N * WA_N;
                sB[hh * WA_N * WA_BSTRIDE + (r / WA_N) * WA_BSTRIDE + (r % WA_N)] = __ldg(p.bias_rq + (long long)(2 * hp + hh) * WA_N * WA_N + r);
            }
            __syncthreads();
            cur_hp = hp;
        }
        const int win = 2 * wp + wsel;                       // my window
        // warp_ok is WARP-UNIFORM (tcgen05.ld is .sync.aligned: never under a divergent branch): every warp of an existing
        // window has valid rows (32 < 49).  Lanes behind the window (irow >= 49) run along on the next window's tokens;
        // only their loads of per-row tables and their stores are suppressed.
        const bool warp_ok = win < p.n_win;
        const bool row_ok = warp_ok && irow < WA_N;
        const int brow_i = irow < WA_N ? irow : WA_N - 1;
        unsigned long long mrow = 0;
        if (p.mask_bits != nullptr && row_ok) mrow = __ldg(p.mask_bits + (long long)(win % p.n_win_img) * WA_N + irow);
#pragma unroll 1
        for (int hh = 0; hh < nh; ++hh, ++g) {
            if (tid == 0) {
                if (g > 0) ptx::mbar_wait(o_done, (g - 1u) & 1u);               // TMEM columns are free again
                if (hh == 0) ptx::mbar_wait(qk_full, it & 1u);
                ptx::tc_fence_after();
                ptx::mma_i8_ss(tmem_base, wa_desc_sw64(sQ) + (uint64_t)(2 * hh), wa_desc_sw64(sK) + (uint64_t)(2 * hh), idesc_s, 0u);
                ptx::mma_commit(s_full);
            }
            __syncwarp();
            ptx::mbar_wait(s_full, g & 1u);
            ptx::tc_fence_after();
            if (tid == 0 && hh == nh - 1 && u + 1 < u1) load_qk(u + 1);        // the last score MMA has read Q and K
            // ---- pass 1: scores -> qact_attn1 -> qact2 (+ bias) -> int8, four per register ----
            uint32_t sc[WA_NW];
            if (warp_ok) {
                uint32_t r[3][16], r48;
                const uint32_t c0 = (uint32_t)(64 * wsel);
                ptx::tmem_ld_32x32b_x16(t_row + c0, r[0]);
                ptx::tmem_ld_32x32b_x16(t_row + c0 + 16, r[1]);
                ptx::tmem_ld_32x32b_x16(t_row + c0 + 32, r[2]);
                ptx::tmem_ld_32x32b_x1(t_row + c0 + 48, r48);
                ptx::tmem_ld_wait();
                const int16_t* brow = sB + hh * WA_N * WA_BSTRIDE + brow_i * WA_BSTRIDE;
                auto rq = [&](uint32_t s, int32_t b) -> int32_t {
                    int32_t a = (int32_t)(((long long)(int32_t)s * (long long)p.m_s + p.half_s) >> 32) >> p.sh_s;
                    a = max(min(a, 127), -128);                                    // qact_attn1: 8 bit
                    const long long t = (long long)a * (long long)p.m_2 + p.half_2;
                    int32_t a2;
                    if constexpr (E2HI) a2 = (int32_t)(t >> 32) >> (p.e_2 - 32);
                    else a2 = (int32_t)__funnelshift_r((uint32_t)t, (uint32_t)(t >> 32), p.e_2);
                    return a2 + b;                                                 // clamped to 8 bits by the pack below
                };
#pragma unroll
                for (int k = 0; k < 12; ++k) {
                    const uint32_t b01 = *reinterpret_cast<const uint32_t*>(brow + 4 * k);
                    const uint32_t b23 = *reinterpret_cast<const uint32_t*>(brow + 4 * k + 2);
                    const int32_t v0 = rq(r[k >> 2][4 * (k & 3)], (int32_t)(int16_t)(b01 & 0xffff));
                    const int32_t v1 = rq(r[k >> 2][4 * (k & 3) + 1], (int32_t)b01 >> 16);
                    const int32_t v2 = rq(r[k >> 2][4 * (k & 3) + 2], (int32_t)(int16_t)(b23 & 0xffff));
                    const int32_t v3 = rq(r[k >> 2][4 * (k & 3) + 3], (int32_t)b23 >> 16);
                    uint32_t hi2;
                    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi2) : "r"(v3), "r"(v2), "r"(0));
                    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(sc[k]) : "r"(v1), "r"(v0), "r"(hi2));
                }
                {
                    int32_t v48 = rq(r48, (int32_t)brow[48]);
                    v48 = max(min(v48, 127), -128);
                    sc[12] = 0x80808000u | (uint32_t)(uint8_t)(int8_t)v48;         // bytes 1..3: padding, -128
                }
                // masked keys (shifted windows): out of the max (forced to -128), exponential = int_exp_shift floor
                if (mrow != 0) {
#pragma unroll
                    for (int k = 0; k < WA_NW; ++k) {
                        const uint32_t b4 = (uint32_t)(mrow >> (4 * k)) & 0xfu;
                        const uint32_t bm = ((b4 * 0x00204081u) & 0x01010101u) * 0xffu;   // bit i -> byte i
                        sc[k] = (sc[k] & ~bm) | (0x80808080u & bm);
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < WA_NW; ++k) sc[k] = 0x80808080u;
            }
            ptx::tc_fence_before();                          // my tcgen05.ld of S are complete: P V may overwrite the columns
            uint32_t pk[16];                                 // my 64 bytes of the probability row (49 + 15 zeros)
#pragma unroll
            for (int k = 0; k < 16; ++k) pk[k] = 0;
            if (warp_ok) {
                // row max (16x2 SIMD max over bytes 3/1 and, shifted, 2/0)
                uint32_t mo = 0x80008000u, me2 = 0x80008000u;
#pragma unroll
                for (int k = 0; k < WA_NW; ++k) {
                    mo = __vmaxs2(mo, sc[k]);
                    me2 = __vmaxs2(me2, sc[k] << 8);
                }
                const int32_t mxs = max(max((int32_t)mo >> 24, (int32_t)(mo << 16) >> 24), max((int32_t)me2 >> 24, (int32_t)(me2 << 16) >> 24));
                // ---- pass 2: exponentials from the table (address = table + 16 * (max - q) + 4 * copy: one IDP.4A), row sum ----
                const int32_t pEq = (int32_t)(sE_lane + (uint32_t)(4 * WA_LUTC) * (uint32_t)mxs);
                uint32_t E[4 * WA_NW];
#pragma unroll
                for (int k = 0; k < WA_NW; ++k) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        int32_t addr;
                        asm("dp4a.s32.s32 %0, %1, %2, %3;" : "=r"(addr) : "r"(sc[k]), "r"((uint32_t)(0x100 - 4 * WA_LUTC) << (8 * i)), "r"(pEq));
                        asm("ld.shared.u32 %0, [%1];" : "=r"(E[4 * k + i]) : "r"(addr));
                    }
                }
                if (mrow != 0) {
#pragma unroll
                    for (int j = 0; j < WA_N; ++j)
                        if ((mrow >> j) & 1ULL) E[j] = emin;
                }
                uint32_t sum = 0;                            // 49 exponentials below 2^26 (host-checked)
#pragma unroll
                for (int j = 0; j < WA_N; ++j) sum += E[j];
                const uint32_t S32 = sum > 2147483647u ? 2147483647u : sum;         // clamp_max_(2**31-1)
                const uint32_t F = 2147483647u / (S32 ? S32 : 1u);                  // <= 65535 (E(0) >= 2^15)
                const uint32_t Fs = F << 8;                                         // P = (E*F) >> 24 == umulhi(E, F << 8)
                // ---- pass 3: 8-bit probabilities ----
#pragma unroll
                for (int k = 0; k < 12; ++k) {
                    const uint32_t P01 = __umulhi(E[4 * k + 1], Fs) * 256u + __umulhi(E[4 * k], Fs);
                    const uint32_t P23 = __umulhi(E[4 * k + 3], Fs) * 256u + __umulhi(E[4 * k + 2], Fs);
                    pk[k] = __byte_perm(P01, P23, 0x5410);
                }
                pk[12] = __umulhi(E[48], Fs);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint32_t a = sP_row + ((((uint32_t)(4 * wsel + c)) ^ ((uint32_t)trow & 7u)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pk[4 * c]), "r"(pk[4 * c + 1]), "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3]) : "memory");
            }
            ptx::fence_proxy_async();                        // P written through the generic proxy -> visible to the MMA
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(p_ready);
            if (tid == 0) {
                ptx::mbar_wait(p_ready, g & 1u);
                if (hh == 0) ptx::mbar_wait(v_full, it & 1u);
                ptx::tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)               // 128 keys (both windows; block-diagonal P) in steps of 32
                    ptx::mma_i8_ss(tmem_base, ptx::umma_desc_k_sw128(sP) + (uint64_t)(2 * kk), wa_desc_sw64(sV + (uint32_t)(kk * 2048)), idesc_pv, kk ? 1u : 0u);
                ptx::mma_commit(o_full);
            }
            __syncwarp();
            // ---- output rows: O -> qact3 -> int8 (my head's 32 of the 64 channel columns) ----
            ptx::mbar_wait(o_full, g & 1u);
            ptx::tc_fence_after();
            if (tid == 0 && hh == nh - 1 && u + 1 < u1) load_v(u + 1);        // the last P V MMA has read V
            if (warp_ok) {
                uint32_t o[2][16];
                ptx::tmem_ld_32x32b_x16(t_row + (uint32_t)(32 * hh), o[0]);
                ptx::tmem_ld_32x32b_x16(t_row + (uint32_t)(32 * hh + 16), o[1]);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(o_done);
                uint32_t ow[8];
#pragma unroll
                for (int w = 0; w < 8; ++w) {
                    int32_t q[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        q[e] = (int32_t)(((long long)(int32_t)o[w >> 2][4 * (w & 3) + e] * (long long)p.m_o + p.half_o) >> 32) >> p.sh_o;
                    uint32_t hi2;
                    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi2) : "r"(q[3]), "r"(q[2]), "r"(0));
                    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(ow[w]) : "r"(q[1]), "r"(q[0]), "r"(hi2));
                }
                if (row_ok) {
                    uint4* dst = reinterpret_cast<uint4*>(out + ((long long)win * WA_N + irow) * (long long)p.C + (2 * hp + hh) * 32);
                    dst[0] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                    dst[1] = make_uint4(ow[4], ow[5], ow[6], ow[7]);
                }
            } else {
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(o_done);
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 128);
    }
}

typedef CUresult (*EncodeTiledFnW)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool wa_fast(ivit_dyadic_t d, int zbits) {            // hi32(z*m + half) >> (e-32) exact: 32 <= e <= 62, no reachable tie
    if (d.m == 0 || d.e < 32 || d.e > 62) return false;
    return d.e - 1 - __builtin_ctz((unsigned)d.m) > zbits;
}

}  // namespace ivit

using namespace ivit;

extern "C" int ivit_window_attention_i8(ivit_ctx* ctx, const int8_t* qkv, const ivit_winattn_params* wp, int8_t* out,
                                        ivit_stream stream) {
    IVIT_REQUIRE(ctx && qkv && wp && out && wp->bias_rq, "ivit_window_attention_i8: null pointer");
    IVIT_REQUIRE(wp->n_win > 0 && wp->n_heads > 0, "ivit_window_attention_i8: bad shape");
    if (wp->n_tok != WA_N || wp->head_dim != 32 || wp->p_bits != 8 || wp->n != 15)
        return fail(IVIT_ENOTSUP, "ivit_window_attention_i8: 49-token windows, head_dim 32, 8-bit probabilities, n = 15 only "
                                  "(use ivit_attention_i8)");
    IVIT_REQUIRE(((uintptr_t)qkv % 16) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)wp->bias_rq % 2) == 0, "ivit_window_attention_i8: alignment");
    if (wp->mask_bits) IVIT_REQUIRE(wp->n_win_img > 0 && wp->n_win % wp->n_win_img == 0, "ivit_window_attention_i8: n_win must be a multiple of n_win_img");
    // preconditions of the fast forms; the general kernel (ivit_attention_i8) covers everything else
    const long long e0 = ((long long)(-wp->x0)) << wp->n;
    const bool ok_s = wa_fast(wp->me_s, 21) && wa_fast(wp->me_o, 23);
    const bool ok_2 = wp->me_s2.m != 0 && wp->me_s2.e >= 8 && wp->me_s2.e <= 62 &&
                      (wp->me_s2.e - 1 - __builtin_ctz((unsigned)wp->me_s2.m) > 8);
    // a masked entry must saturate Shiftmax whatever the row holds: t(add + 255) <= n * x0 with t(d) <= 23 d / 16 + 1
    const bool ok_m = wp->mask_bits == nullptr || ((23LL * ((long long)wp->mask_add + 255)) / 16 + 1 <= (long long)wp->n * wp->x0);
    if (!(wp->x0 <= -1 && e0 < (1LL << 26) && ok_s && ok_2 && ok_m && ctx->encode_tiled))
        return fail(IVIT_ENOTSUP, "ivit_window_attention_i8: scales outside the fast-form domain (x0=%d, e_s=%d, e_2=%d, e_o=%d, mask_add=%d); "
                                  "use ivit_attention_i8", wp->x0, wp->me_s.e, wp->me_s2.e, wp->me_o.e, wp->mask_add);
    WinAttnArgs a;
    a.n_win = wp->n_win; a.H = wp->n_heads; a.C = 32 * wp->n_heads;
    a.n_wp = (wp->n_win + 1) / 2; a.n_hp = (wp->n_heads + 1) / 2;
    a.m_s = wp->me_s.m; a.sh_s = wp->me_s.e - 32; a.half_s = 1LL << (wp->me_s.e - 1);
    a.m_2 = wp->me_s2.m; a.e_2 = wp->me_s2.e; a.half_2 = 1LL << (wp->me_s2.e - 1);
    a.m_o = wp->me_o.m; a.sh_o = wp->me_o.e - 32; a.half_o = 1LL << (wp->me_o.e - 1);
    a.x0 = wp->x0; a.inv_x0 = 1.0f / (float)wp->x0; a.n = wp->n;
    a.bias_rq = wp->bias_rq; a.mask_bits = (const unsigned long long*)wp->mask_bits; a.n_win_img = wp->n_win_img > 0 ? wp->n_win_img : 1;
    // 2D uint8 tensor map over the packed qkv activations [n_win * 49, 3C]; box {64 B, 64 rows}, 64B swizzle.  Rows past
    // the end (the second window of an odd last pair, the 15 rows behind the last window) read as zeros.
    CUtensorMap tm;
    {
        cuuint64_t gdim[2] = {(cuuint64_t)(3 * a.C), (cuuint64_t)wp->n_win * WA_N};
        cuuint64_t gstride[1] = {(cuuint64_t)(3 * a.C)};
        cuuint32_t box[2] = {64, 64};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = reinterpret_cast<EncodeTiledFnW>(ctx->encode_tiled)(
            &tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<int8_t*>(qkv), gdim, gstride, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(IVIT_ECUDA, "cuTensorMapEncodeTiled (window qkv) failed (CUresult %d)", (int)r);
    }
    static PerDevice attr_set;
    if (!attr_set[ctx->device]) {
        IVIT_CUDA_OK(cudaFuncSetAttribute(window_attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WA_SMEM));
        IVIT_CUDA_OK(cudaFuncSetAttribute(window_attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WA_SMEM));
        IVIT_CUDA_OK(cudaFuncSetAttribute(window_attention_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        IVIT_CUDA_OK(cudaFuncSetAttribute(window_attention_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr_set[ctx->device] = 1;
    }
    const int units = a.n_wp * a.n_hp;
    const int grid = units < 4 * ctx->num_sms ? units : 4 * ctx->num_sms;      // persistent: four resident CTAs per SM
    if (a.e_2 >= 32) window_attention_kernel<true><<<grid, 128, WA_SMEM, st(stream)>>>(tm, a, out);
    else window_attention_kernel<false><<<grid, 128, WA_SMEM, st(stream)>>>(tm, a, out);
    IVIT_LAUNCH_OK("window_attention_kernel");
    return IVIT_OK;
}
