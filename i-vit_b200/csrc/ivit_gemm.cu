// INT8 GEMM on the 5th-generation tensor cores (tcgen05.mma kind::i8, sm_100a) with the dyadic
// requantisation of the following QuantAct fused into the epilogue.
//
//   acc[i,j] = sum_k A[i,k] * W[j,k] (+ bias[j])              QuantLinear.forward, quant_modules.py:93-97
//   out[i,j] = clamp(RNE(acc * m[j] / 2^e[j]) ...)            QuantAct / fixedpoint_mul, quant_utils.py:192-253
//
// Structure (one persistent CTA per SM, 352 threads; for M >= 512 and 256-wide tiles the CTAs of a TPC run as a pair,
// tcgen05 cta_group::2, on a 256 x 256 output tile -- see the kernel's comment):
//   warp 0      TMA producer: A tile [128 x 128 B] and W tile [BN (or BN/2 in a pair) x 128 B] per k-block into a
//               STAGES-deep shared-memory ring (128-byte swizzle), mbarrier full/empty pairs
//   warp 1      TMEM allocator + MMA issuer: one thread issues 4 x tcgen05.mma (K = 32 each) per
//               k-block into one of two TMEM accumulators (128 lanes x BN int32 columns each)
//   warp 2      auxiliary warp: per-column requant constants of the tile two ahead -> shared memory, requant form vote
//   warps 3-10  epilogue (two warps per TMEM lane group, half of the tile's columns each):
//               tcgen05.ld 16/32 columns at a time (prefetched one chunk ahead) -> per-channel dyadic requant
//               (+ second stage + int16 residual, which arrives by TMA for the K <= 1024 shapes) in registers ->
//               128B-swizzled staging tile -> one TMA store per warp.  Runs concurrently with the MMAs of the next
//               tile (double-buffered accumulator); no CTA-wide barrier in the loop.
// Both operands are K-major (row-major A [M,K], row-major W [N,K]), so no transposes anywhere.
#include <stdio.h>
#include <stdlib.h>

#include "ivit_common.cuh"
#include "ivit_internal.h"
#include "ivit_ptx.cuh"

namespace ivit {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 128;          // bytes == int8 elements per k-block (one 128 B swizzle row)
constexpr int GEMM_UMMA_K = 32;       // K per tcgen05.mma for 8-bit operands
// epilogue warps per TMEM lane group (each takes 1/WPG of the tile's columns).  4 (16 epilogue warps, 96 registers)
// was measured slower than 2 for every epilogue, including the residual one (profiles/gemm_bench_r18.log).
__host__ __device__ constexpr int gemm_wpg(int mode) { return (void)mode, 2; }
// TMA producer, MMA issuer, auxiliary warp, epilogue warps
__host__ __device__ constexpr int gemm_threads(int mode) { return 96 + 128 * gemm_wpg(mode); }

enum GemmMode { GM_RAW_I32 = 0, GM_CARRIER = 1, GM_RQ_I8 = 2, GM_RQ_I16 = 3 };

// A scalar dyadic applied to 16-bit operands (|z| < 2^15), analysed on the host (it knows (m, e) by value).
// Unified branch-free form for 16 <= e <= 62:
//     t = (z << ls) * m + half';   q = hi32(t) >> rs        with (ls, rs, half') = (32-e, 0, 2^31)   for e <  32
//                                                                               (0, e-32, 2^(e-1))  for e >= 32
// (scaling z by 2^(32-e) scales numerator and denominator alike, |z << ls| < 2^31).  `tie` says whether an exact
// .5 tie is reachable (0 <= e-1-ctz(m) <= 15); then tie <=> lo32(t) == 0 && (hi32(t) & himask) == 0 and q -= q & 1.
// kind 0: e outside [16, 62] -> general out-of-line form.
struct ScalarRq {
    int32_t m, e;
    int32_t ls, rs, himask;
    long long half;
    int kind;                         // 0 general, 1 unified form
    int tie;                          // unified form needs the tie-to-even correction
    // fixed-shift form for 16 <= e <= 47 (|z| < 2^15):  q = hi32((z << 16) * m + 2^(e+15)) >> (e - 16).  The operand
    // shifted left by 16 is what a packed pair of int16 gives for free (high half: mask, low half: one shift).
    int32_t rs16;
    long long half16;
};
static ScalarRq make_scalar_rq(ivit_dyadic_t d) {
    ScalarRq r;
    r.m = d.m; r.e = d.e; r.ls = 0; r.rs = 0; r.himask = 0; r.half = 0; r.kind = 0; r.tie = 0; r.rs16 = 0; r.half16 = 0;
    if (d.e >= 16 && d.e <= 47) { r.rs16 = d.e - 16; r.half16 = 1LL << (d.e + 15); }
    if (d.m != 0 && d.e >= 16 && d.e <= 62) {
        r.kind = 1;
        if (d.e < 32) { r.ls = 32 - d.e; r.rs = 0; r.half = 1LL << 31; }
        else { r.ls = 0; r.rs = d.e - 32; r.half = 1LL << (d.e - 1); }
        r.himask = (1 << r.rs) - 1;
        const int t = d.e - 1 - __builtin_ctz((unsigned)d.m);     // t < 0: z*m/2^e is an integer, nothing to round
        r.tie = (t >= 0 && t <= 15) ? 1 : 0;
    }
    return r;
}
template <bool TIE>
__device__ __forceinline__ int32_t scalar_rq_fast(const ScalarRq& u, int32_t z) {
    const long long t = (long long)(z << u.ls) * (long long)u.m + u.half;
    const int32_t hi = (int32_t)(t >> 32);
    int32_t q = hi >> u.rs;
    if (TIE) {
        const bool tie = ((uint32_t)t == 0u) && ((hi & u.himask) == 0);
        q -= (int32_t)(tie & (q & 1));
    }
    return q;
}
__device__ __forceinline__ int32_t add_sat_s32(int32_t a, int32_t b) {
    int32_t r;
    asm("add.sat.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}

// The IVIT_GEMM_DEBUG diagnostics (skip epilogue work / operand loads / residual loads / output stores) are compiled in only
// with -DIVIT_GEMM_DIAG (python i-vit_b200/csrc/build.py --diag): even never-taken branches in the chunk loop cost the
// production kernel several per cent (register allocation / code layout).
#ifdef IVIT_GEMM_DIAG
#define IVIT_DBG(args, bit) ((args).debug & (bit))
#else
#define IVIT_DBG(args, bit) 0
#endif

struct GemmArgs {
    int M, N, K;
    int mode_bits;                    // clamp bits for requant modes
    const int32_t* bias;
    const ivit_dyadic_t* me;
    const void* residual;             // int16 [M, res_ld] (GM_RQ_I16 only)
    long long res_ld;
    int two_stage;
    ScalarRq rq2, rqr;                // second-stage / residual dyadics, pre-analysed on the host (known by value)
    int acc_bits;                     // |acc + bias| < 2^acc_bits (tie analysis of the per-column requant)
    int scalar_mode;                  // 1: rq2/rqr unified form, no ties; 2: unified with tie correction; 0: general;
                                      // 3: mode 1 with both stages present and a wrap-free 32-bit sum (straight-line)
                                      // 4: mode 3 with both exponents <= 47 (fixed-shift form on packed int16 pairs)
    const float* scale;
    void* out;
    long long out_ld;
    int res_async;                    // residual rows are 16-byte aligned: prefetched into the staging tile with cp.async
    int debug;                        // IVIT_GEMM_DEBUG diagnostics (wrong results): 1 epilogue does no work, 2 producer loads nothing,
                                      // 4 residual reads as zero (no global loads), 8 no TMA store of the output
};

struct alignas(16) ColParam {         // per output column, staged in shared memory per tile
    int32_t m;                        // dyadic multiplier (or the fp32 scale bits for GM_CARRIER)
    int32_t sh;                       // e - 32
    long long c;                      // bias*m + 2^(e-1): the fast requant is hi32(acc*m + c) >> sh
};

// one 16-byte (warp-broadcast) shared-memory load per column: the epilogue is bounded by shared-memory wavefronts
__device__ __forceinline__ ColParam ld_col_param(const ColParam* p) {
    const int4 t = *reinterpret_cast<const int4*>(p);
    ColParam r;
    r.m = t.x; r.sh = t.y;
    r.c = (long long)(((unsigned long long)(uint32_t)t.w << 32) | (uint32_t)t.z);
    return r;
}

__device__ __forceinline__ uint32_t pack_sat_s8x4(int32_t a, int32_t b, int32_t c, int32_t d) {
    uint32_t hi, r;
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(d), "r"(c), "r"(0));
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(a), "r"(hi));
    return r;                                                      // a | b<<8 | c<<16 | d<<24, each saturated to int8
}
__device__ __forceinline__ uint32_t pack_sat_s16x2(int32_t lo, int32_t hi) {
    uint32_t r;
    asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(r) : "r"(hi), "r"(lo));
    return r;
}

// Residual prefetch for one CW-column chunk of one row (int16): CW/8 x 16-byte loads issued early so that
// their latency overlaps the requant arithmetic of the previous chunk.
template <int CW>
__device__ __forceinline__ void load_residual(const GemmArgs& args, int row, bool row_ok, int ncol0, uint32_t (&rr)[CW / 2]) {
#pragma unroll
    for (int j = 0; j < CW / 2; ++j) rr[j] = 0u;
    if (!args.residual || !row_ok || ncol0 >= args.N || (IVIT_DBG(args, 4))) return;   // debug & 4: diagnostics, residual reads as zero
    const int16_t* res = reinterpret_cast<const int16_t*>(args.residual) + (long long)row * args.res_ld + ncol0;
    if (ncol0 + CW <= args.N && ((reinterpret_cast<uintptr_t>(res) & 15) == 0)) {
#pragma unroll
        for (int j = 0; j < CW / 8; ++j) {
            const uint4 t = __ldg(reinterpret_cast<const uint4*>(res) + j);
            rr[4 * j] = t.x; rr[4 * j + 1] = t.y; rr[4 * j + 2] = t.z; rr[4 * j + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < CW; ++j) {
            const uint32_t v = (ncol0 + j < args.N) ? (uint32_t)(uint16_t)res[j] : 0u;
            rr[j >> 1] |= v << (16 * (j & 1));
        }
    }
}

// One CW-column chunk of one output row: registers r[] hold the int32 accumulators.
// 16-byte store into the 128B-swizzled staging tile: logical (row, byte offset inside the BN*ES-byte tile row)
__device__ __forceinline__ void sts_swz(uint32_t out_base, int trow, int byte_off, uint4 v) {
    const uint32_t box = (uint32_t)byte_off >> 7, chunk = ((uint32_t)byte_off >> 4) & 7u;
    const uint32_t addr = out_base + box * (uint32_t)(GEMM_BM * 128) + (uint32_t)trow * 128u + ((chunk ^ ((uint32_t)trow & 7u)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ uint32_t swz_addr(uint32_t out_base, int trow, int byte_off) {
    const uint32_t box = (uint32_t)byte_off >> 7, chunk = ((uint32_t)byte_off >> 4) & 7u;
    return out_base + box * (uint32_t)(GEMM_BM * 128) + (uint32_t)trow * 128u + ((chunk ^ ((uint32_t)trow & 7u)) << 4);
}
// 16 bytes global -> shared, asynchronous (LDGSTS); src_bytes = 0 zero-fills without reading
__device__ __forceinline__ void cp_async_16(uint32_t dst_smem, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// residual of one CW-column chunk of my row, from the staging tile (where residual_prefetch put it, in the layout of
// the output that later overwrites it in place)
template <int CW>
__device__ __forceinline__ void load_residual_smem(uint32_t out_base, int trow, int tcol, uint32_t (&rr)[CW / 2]) {
#pragma unroll
    for (int j = 0; j < CW / 8; ++j) {
        const uint32_t a = swz_addr(out_base, trow, (tcol + 8 * j) * 2);
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(rr[4 * j]), "=r"(rr[4 * j + 1]), "=r"(rr[4 * j + 2]), "=r"(rr[4 * j + 3]) : "r"(a) : "memory");
    }
}

template <int MODE, int CW, bool TS>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&r)[CW], const uint32_t (&rr)[CW / 2],
                                               const ColParam* __restrict__ cp, const int32_t* __restrict__ cb,
                                               const GemmArgs& args, int row, bool row_ok, int ncol0, int fast,
                                               uint32_t out_base, int trow, int tcol) {
    const bool full_chunk = (ncol0 + CW <= args.N);
    if (MODE == GM_RAW_I32 || MODE == GM_CARRIER) {
        uint32_t o[CW];
#pragma unroll
        for (int j = 0; j < CW; ++j) {
            const int32_t v = (int32_t)r[j] + cb[j];
            o[j] = (MODE == GM_RAW_I32) ? (uint32_t)v : __float_as_uint(__fmul_rn(__int2float_rn(v), __int_as_float(cp[j].m)));
        }
        if (!row_ok) return;
        uint32_t* dst = reinterpret_cast<uint32_t*>(args.out) + (long long)row * args.out_ld + ncol0;
        if (full_chunk && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < CW; j += 4) *reinterpret_cast<uint4*>(dst + j) = make_uint4(o[j], o[j + 1], o[j + 2], o[j + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < CW; ++j)
                if (ncol0 + j < args.N) dst[j] = o[j];
        }
        return;
    }
    // ---- first-stage per-channel requant: q = RNE((acc + bias) * m / 2^e) ----
    // tile mode (uniform): 1 = every column has 32 <= e <= 62 and none can reach an exact tie
    //                      2 = every column has 32 <= e <= 62, some can tie -> correction on all elements
    //                      0 = general out-of-line form
    int32_t q[CW];
    if (fast == 1) {
#pragma unroll
        for (int j = 0; j < CW; ++j) {
            const ColParam p = ld_col_param(cp + j);
            const long long t = (long long)(int32_t)r[j] * (long long)p.m + p.c;
            q[j] = (int32_t)(t >> 32) >> p.sh;
        }
    } else if (fast == 2) {
#pragma unroll
        for (int j = 0; j < CW; ++j) {
            const ColParam p = ld_col_param(cp + j);
            const long long t = (long long)(int32_t)r[j] * (long long)p.m + p.c;
            const int32_t hi = (int32_t)(t >> 32);
            int32_t v = hi >> p.sh;
            const bool tie = ((uint32_t)t == 0u) && ((hi & ((1 << p.sh) - 1)) == 0);
            q[j] = v - (int32_t)(tie & (v & 1));
        }
    } else {
#pragma unroll
        for (int j = 0; j < CW; ++j) {
            const ColParam p = ld_col_param(cp + j);
            q[j] = requant32_general((int32_t)r[j] + cb[j], p.m, p.sh + 32);
        }
    }
    if (MODE == GM_RQ_I8) {
        if (TS) {
#pragma unroll
            for (int j = 0; j < CW; j += 16)
                sts_swz(out_base, trow, tcol + j, make_uint4(
                    pack_sat_s8x4(q[j], q[j + 1], q[j + 2], q[j + 3]), pack_sat_s8x4(q[j + 4], q[j + 5], q[j + 6], q[j + 7]),
                    pack_sat_s8x4(q[j + 8], q[j + 9], q[j + 10], q[j + 11]), pack_sat_s8x4(q[j + 12], q[j + 13], q[j + 14], q[j + 15])));
            return;
        }
        if (!row_ok) return;
        int8_t* dst = reinterpret_cast<int8_t*>(args.out) + (long long)row * args.out_ld + ncol0;
        if (full_chunk && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < CW; j += 16)
                *reinterpret_cast<uint4*>(dst + j) = make_uint4(
                    pack_sat_s8x4(q[j], q[j + 1], q[j + 2], q[j + 3]), pack_sat_s8x4(q[j + 4], q[j + 5], q[j + 6], q[j + 7]),
                    pack_sat_s8x4(q[j + 8], q[j + 9], q[j + 10], q[j + 11]), pack_sat_s8x4(q[j + 12], q[j + 13], q[j + 14], q[j + 15]));
        } else {
#pragma unroll
            for (int j = 0; j < CW; ++j)
                if (ncol0 + j < args.N) dst[j] = (int8_t)clamp_bits<8>(q[j]);
        }
        return;
    }
    // ---- GM_RQ_I16 (16-bit clamp): optional second (scalar) stage and int16 residual ----
    //   single stage: clamp(RNE(z*me) + RNE(res*res_me))       one QuantAct with identity
    //   two stage   : q1 = clamp(RNE(z*me)) is a QuantAct output; a second QuantAct adds the residual
    //                 (attn.qact3 -> Block.qact2, mlp.qact2 -> Block.qact4; vit_quant.py:85,135,141)
    if (!TS && !row_ok) return;
    const bool has_res = args.residual != nullptr;
    auto resid = [&](int j) -> int32_t {
        return (j & 1) ? ((int32_t)rr[j >> 1] >> 16) : (int32_t)(int16_t)(rr[j >> 1] & 0xffff);
    };
    if (args.scalar_mode == 4) {
        // scalar_mode 3 with both exponents <= 47: the 16-bit clamp of the first stage is the saturating pack of two
        // results into one word, and both that word and the packed residual word yield their operands already shifted
        // left by 16 with one instruction each (mask / shift) -- 11 instead of 14 instructions per element.
#pragma unroll
        for (int j = 0; j < CW; j += 2) {
            const uint32_t d = pack_sat_s16x2(q[j], q[j + 1]);
            const uint32_t w = rr[j >> 1];
            const int32_t zl = (int32_t)(d << 16), zh = (int32_t)(d & 0xffff0000u);
            const int32_t rl = (int32_t)(w << 16), rh = (int32_t)(w & 0xffff0000u);
            q[j] = ((int32_t)(((long long)zl * (long long)args.rq2.m + args.rq2.half16) >> 32) >> args.rq2.rs16) +
                   ((int32_t)(((long long)rl * (long long)args.rqr.m + args.rqr.half16) >> 32) >> args.rqr.rs16);
            q[j + 1] = ((int32_t)(((long long)zh * (long long)args.rq2.m + args.rq2.half16) >> 32) >> args.rq2.rs16) +
                       ((int32_t)(((long long)rh * (long long)args.rqr.m + args.rqr.half16) >> 32) >> args.rqr.rs16);
        }
    } else if (args.scalar_mode == 3) {
        // the residual-block epilogue of the models (attn.proj / mlp.fc2): two-stage + residual, both scalar dyadics in
        // unified form without reachable ties, and |each term| < 2^30 so that the 32-bit sum cannot wrap (host-checked).
        // One straight-line block per chunk: no per-element conditions, the compiler interleaves the 16 chains.
#pragma unroll
        for (int j = 0; j < CW; ++j)
            q[j] = scalar_rq_fast<false>(args.rq2, clamp_bits<16>(q[j])) + scalar_rq_fast<false>(args.rqr, resid(j));
    } else if (args.scalar_mode == 1) {                            // both scalar dyadics in unified form, no ties
#pragma unroll
        for (int j = 0; j < CW; ++j) {
            int32_t v = q[j];
            if (args.two_stage) v = scalar_rq_fast<false>(args.rq2, clamp_bits<16>(v));
            if (has_res) v = add_sat_s32(v, scalar_rq_fast<false>(args.rqr, resid(j)));
            q[j] = v;
        }
    } else if (args.scalar_mode == 2) {                            // unified form with tie-to-even correction
#pragma unroll
        for (int j = 0; j < CW; ++j) {
            int32_t v = q[j];
            if (args.two_stage) v = scalar_rq_fast<true>(args.rq2, clamp_bits<16>(v));
            if (has_res) v = add_sat_s32(v, scalar_rq_fast<true>(args.rqr, resid(j)));
            q[j] = v;
        }
    } else {                                                       // general
#pragma unroll
        for (int j = 0; j < CW; ++j) {
            int32_t v = q[j];
            if (args.two_stage) v = requant32_general(clamp_bits<16>(v), args.rq2.m, args.rq2.e);
            if (has_res) v = sat_i64_to_i32((long long)v + (long long)requant32_general(resid(j), args.rqr.m, args.rqr.e));
            q[j] = v;
        }
    }
    if (TS) {
#pragma unroll
        for (int j = 0; j < CW; j += 8)
            sts_swz(out_base, trow, (tcol + j) * 2, make_uint4(pack_sat_s16x2(q[j], q[j + 1]), pack_sat_s16x2(q[j + 2], q[j + 3]),
                                                               pack_sat_s16x2(q[j + 4], q[j + 5]), pack_sat_s16x2(q[j + 6], q[j + 7])));
        return;
    }
    int16_t* dst = reinterpret_cast<int16_t*>(args.out) + (long long)row * args.out_ld + ncol0;
    if (full_chunk && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
        for (int j = 0; j < CW; j += 8)
            *reinterpret_cast<uint4*>(dst + j) = make_uint4(pack_sat_s16x2(q[j], q[j + 1]), pack_sat_s16x2(q[j + 2], q[j + 3]),
                                                            pack_sat_s16x2(q[j + 4], q[j + 5]), pack_sat_s16x2(q[j + 6], q[j + 7]));
    } else {
#pragma unroll
        for (int j = 0; j < CW; ++j)
            if (ncol0 + j < args.N) dst[j] = (int16_t)clamp_bits<16>(q[j]);
    }
}

template <int CW>
__device__ __forceinline__ void tmem_ld_chunk(uint32_t taddr, uint32_t (&r)[CW]) {
    if constexpr (CW == 32) ptx::tmem_ld_32x32b_x32(taddr, r);
    else ptx::tmem_ld_32x32b_x16(taddr, r);
}

// OUT_ES: bytes per output element staged through shared memory for the TMA store (0: direct global stores)
// PAIR: the CTA holds only its half of the W tile (the other half lives in the peer CTA of the pair)
template <int BN, int STAGES, int OUT_ES, bool PAIR, bool RES = false>
struct GemmSmem {
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK;
    static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * GEMM_BK;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int OUT_BYTES = GEMM_BM * BN * OUT_ES;                     // [BN*OUT_ES/128 boxes][128 rows][128 B], 128B-swizzled
    static constexpr int RES_BYTES = RES ? GEMM_BM * BN * 2 : 0;                // int16 residual tile, same box layout (output mode 3)
    static constexpr int PARAM_BYTES = 2 * BN * ((int)sizeof(ColParam) + 4);   // ColParam[2][BN] + int32 bias[2][BN]
    static constexpr int BAR_BYTES = 512;                                       // pipeline barriers [0, 256), residual barriers [256, 512)
    static constexpr int TOTAL = STAGES * STAGE_BYTES + OUT_BYTES + RES_BYTES + PARAM_BYTES + BAR_BYTES + 1024;  // +1024 alignment slack
};

// PAIR = true: CTA-pair kernel (tcgen05 cta_group::2).  The two CTAs of a 2-CTA cluster (the two SMs of a TPC) work on
// two consecutive m-tiles of the same n-tile as ONE 256 x BN MMA: each CTA loads its own 128 rows of A and HALF of the
// W tile, the leader (even rank) issues the MMAs for both, each CTA gets its 128 accumulator rows in its own TMEM and
// runs its own epilogue.  Per MMA an SM's shared memory then feeds 128 x 32 B of A + BN/2 x 32 B of W instead of
// 128 x 32 + BN x 32: the un-paired kernel saturates the SM's shared-memory data pipe (tensor-core operand reads +
// epilogue LDS/STS ~ 93 % of peak wavefronts, profiles/ncu_full_r1g), which is what bounds it, not the tensor pipe.
// OM (output mode): 0 direct global stores from registers; 1 output staged in shared memory and written by TMA;
//                   2 direct stores, the staging tile instead holds the int16 residual, prefetched per tile by cp.async
//                   3 as 1, and the int16 residual tile arrives by TMA in its own shared-memory tile (32-row x 128-byte
//                     boxes, one pair per epilogue warp, refilled for the next tile as soon as a box has been consumed).
//                     Row-per-thread global loads of the residual touch 32 different lines per instruction: reading the
//                     residual as zeros (IVIT_GEMM_DEBUG=4) took 18 of proj's 70 us away, a deeper register prefetch none.
template <int BN, int STAGES, int MODE, int OM, bool PAIR>
__global__ void __launch_bounds__(gemm_threads(MODE), 1)
gemm_i8_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                       const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res,
                       const GemmArgs args) {
    constexpr bool TS = (OM == 1 || OM == 3), RSM = (OM == 2), RTMA = (OM == 3);
    static_assert(!RSM || MODE == GM_RQ_I16, "residual staging belongs to the 16-bit epilogue");
    static_assert(!RTMA || (MODE == GM_RQ_I16 && BN == 256), "TMA residual: 16-bit epilogue, 256-wide tiles");
    constexpr int OUT_ES = (OM == 0) ? 0 : (MODE == GM_RQ_I8 ? 1 : 2);
    using S = GemmSmem<BN, STAGES, OUT_ES, PAIR, RTMA>;
    constexpr uint32_t TMEM_COLS = 2 * BN;            // double-buffered accumulator (256 or 512)
    constexpr int WPG = gemm_wpg(MODE);
    constexpr int EPI_WARPS = 4 * WPG;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - ptx::smem_u32(smem_raw));

    const uint32_t stage_base = smem_base;
    const uint32_t out_base = smem_base + STAGES * S::STAGE_BYTES;             // 1024-aligned (stage sizes are)
    const uint32_t res_base = out_base + S::OUT_BYTES;                         // 1024-aligned (OUT_BYTES is a multiple)
    ColParam* col_params = reinterpret_cast<ColParam*>(smem + STAGES * S::STAGE_BYTES + S::OUT_BYTES + S::RES_BYTES);
    int32_t* col_bias = reinterpret_cast<int32_t*>(col_params + 2 * BN);
    const uint32_t bar_base = smem_base + STAGES * S::STAGE_BYTES + S::OUT_BYTES + S::RES_BYTES + S::PARAM_BYTES;
    auto res_bar = [&](int ew, int h) { return bar_base + 256u + 8u * (uint32_t)(ew * 2 + h); };   // residual box h of epilogue warp ew landed
    // barrier layout (8 B each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], pfull[2], sfull[2], then tmem ptr + flags
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
    auto pfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 4 + s); };     // column constants staged (aux -> epilogue)
    auto sfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 6 + s); };     // epilogue warps done with a constants buffer (-> aux)
    volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + STAGES * S::STAGE_BYTES + S::OUT_BYTES + S::RES_BYTES + S::PARAM_BYTES + 8 * (2 * STAGES + 8));
    int* fast_flag = reinterpret_cast<int*>(const_cast<uint32_t*>(tmem_ptr_smem) + 2);   // [2] one per accumulator stage

    ptx::grid_dep_wait();                             // A / residual are the predecessor's outputs (programmatic dependent launch)
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    constexpr int CS = PAIR ? 2 : 1;                             // CTAs per cluster, stacked along M
    const int ry = PAIR ? (int)ptx::cluster_ctarank() : 0;       // 0 = leader
    const int tiles_m = (args.M + GEMM_BM - 1) / GEMM_BM;
    const int tiles_n = (args.N + BN - 1) / BN;
    const int tiles_m_c = (tiles_m + CS - 1) / CS;
    const int num_tiles = tiles_m_c * tiles_n;                   // cluster tiles (CS CTA tiles each)
    const int tile_first = (int)blockIdx.x / CS, tile_step = (int)gridDim.x / CS;
    const int num_kb = (args.K + GEMM_BK - 1) / GEMM_BK;
    // CTA tile of cluster tile `t`: an out-of-range m-tile of the peer (odd tile count) still runs the whole pipeline
    // (TMA zero-fills, stores are clipped) so that the pair stays in lock step.
    auto tile_m0 = [&](int t) { return ((t / tiles_n) * CS + ry) * GEMM_BM; };
    auto tile_n0 = [&](int t) { return (t % tiles_n) * BN; };

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmap_a);
        ptx::prefetch_tensormap(&tmap_b);
        if (TS) ptx::prefetch_tensormap(&tmap_out);
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(full_bar(s), 1);                // PAIR: only the leader's is used (both CTAs' loads signal it)
            ptx::mbar_init(empty_bar(s), 1);               // one tcgen05.commit arrival (PAIR: multicast by the leader)
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(tfull_bar(s), 1);
            ptx::mbar_init(tempty_bar(s), CS * EPI_WARPS); // one arrive per epilogue warp (PAIR: of both CTAs, on the leader's)
            ptx::mbar_init(pfull_bar(s), 1);
            ptx::mbar_init(sfull_bar(s), EPI_WARPS);
        }
        if (RTMA)
            for (int i = 0; i < 2 * EPI_WARPS; ++i) ptx::mbar_init(res_bar(i >> 1, i & 1), 1);
        if (RTMA) ptx::prefetch_tensormap(&tmap_res);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        if (PAIR) {
            ptx::tmem_alloc_pair(ptx::smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), TMEM_COLS);
            ptx::tmem_relinquish_pair();
        } else {
            ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), TMEM_COLS);
            ptx::tmem_relinquish();
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (PAIR) ptx::cluster_sync();                         // the peer's barriers and TMEM exist before any remote arrive / pair MMA
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t full0 = PAIR ? ptx::mapa(full_bar(0), 0) : 0u;   // leader's full[0] in the cluster window
            for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
                const int m0 = tile_m0(tile), n0 = tile_n0(tile);
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(empty_bar(stage), phase ^ 1u);    // the MMAs that read this stage have retired
                    const uint32_t a_dst = stage_base + stage * S::STAGE_BYTES;
                    const uint32_t b_dst = a_dst + S::A_BYTES;
                    if (IVIT_DBG(args, 2)) {                            // diagnostics: mainloop without operand traffic
                        if (ry == 0) ptx::mbar_arrive(full_bar(stage));
                    } else if (PAIR) {
                        // the leader's barrier collects the bytes of both CTAs' loads
                        if (ry == 0) ptx::mbar_arrive_expect_tx(full_bar(stage), 2 * S::STAGE_BYTES);
                        ptx::tma_load_2d_pair(a_dst, &tmap_a, full0 + 8u * stage, kb * GEMM_BK, m0);
                        ptx::tma_load_2d_pair(b_dst, &tmap_b, full0 + 8u * stage, kb * GEMM_BK, n0 + ry * (BN / 2));
                    } else {
                        ptx::mbar_arrive_expect_tx(full_bar(stage), S::STAGE_BYTES);
                        ptx::tma_load_2d(a_dst, &tmap_a, full_bar(stage), kb * GEMM_BK, m0);
                        ptx::tma_load_2d(b_dst, &tmap_b, full_bar(stage), kb * GEMM_BK, n0);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (PAIR: leader CTA only) =================
        if (lane == 0 && ry == 0) {
            constexpr uint32_t idesc = ptx::umma_idesc_i8(CS * GEMM_BM, BN, 1, 1);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
                if (PAIR) ptx::mbar_wait_cluster(tempty_bar(as), aphase ^ 1u);
                else ptx::mbar_wait(tempty_bar(as), aphase ^ 1u);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(full_bar(stage), phase);
                    ptx::tc_fence_after();
                    const uint32_t a_addr = stage_base + stage * S::STAGE_BYTES;
                    const uint32_t b_addr = a_addr + S::A_BYTES;
                    const uint64_t a_desc = ptx::umma_desc_k_sw128(a_addr);
                    const uint64_t b_desc = ptx::umma_desc_k_sw128(b_addr);
#pragma unroll
                    for (int k = 0; k < GEMM_BK / GEMM_UMMA_K; ++k) {
                        // advance along K inside the 128 B swizzle row: +32 B == +2 in the (addr >> 4) field
                        if (PAIR) ptx::mma_i8_ss_pair(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                                                      (kb | k) != 0 ? 1u : 0u);
                        else ptx::mma_i8_ss(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                                            (kb | k) != 0 ? 1u : 0u);
                    }
                    // frees the smem slot when the MMAs retire (PAIR: in both CTAs)
                    if (PAIR) ptx::mma_commit_pair_mc(empty_bar(stage), (uint16_t)3);
                    else ptx::mma_commit(empty_bar(stage));
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                // accumulator complete -> epilogue (PAIR: of both CTAs)
                if (PAIR) ptx::mma_commit_pair_mc(tfull_bar(as), (uint16_t)3);
                else ptx::mma_commit(tfull_bar(as));
                if (++as == 2) { as = 0; aphase ^= 1u; }
            }
        }
    } else if (warp == 2) {
        // ================= auxiliary warp: per-tile column constants =================
        // Keeps the staging (global loads + analysis) off the epilogue warps' critical path: they never execute a CTA
        // barrier, only mbarrier waits that have normally completed long before.  Constants of tile it live in buffer it & 1.
        // The global loads for tile it+2 are issued BEFORE waiting for the epilogue warps to finish tile it (whose buffer
        // they will overwrite), so that only the analysis + shared-memory stores remain once the buffer is free.
        constexpr int CPL = BN / 32;                             // columns per lane
        int32_t pb[CPL];
        ivit_dyadic_t pd[CPL];
        float pscale[CPL];
        auto load_params = [&](int tile) {
            const int n0 = tile_n0(tile);
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                const int n = n0 + lane + 32 * j;
                pb[j] = 0; pd[j].m = 0; pd[j].e = 63; pscale[j] = 0.f;
                if (n < args.N) {
                    if (args.bias) pb[j] = __ldg(args.bias + n);
                    if (MODE == GM_RQ_I8 || MODE == GM_RQ_I16) pd[j] = args.me[n];
                    else if (MODE == GM_CARRIER) pscale[j] = __ldg(args.scale + n);
                }
            }
        };
        auto store_params = [&](int buf) {
            ColParam* cp = col_params + buf * BN;
            int32_t* cb = col_bias + buf * BN;
            int ok = 1, any_tie = 0;
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                ColParam p;
                p.m = 0; p.sh = 31; p.c = 0;
                const int32_t b = pb[j];
                if (MODE == GM_RQ_I8 || MODE == GM_RQ_I16) {
                    const ivit_dyadic_t d = pd[j];
                    if (d.m != 0) {
                        // fast form: t = acc*m + (bias*m + 2^(e-1)); q = hi32(t) >> (e-32), needs 32 <= e <= 62
                        // (whole tile).  An exact tie z*m = (2k+1)*2^(e-1) needs v2(z) = e-1-ctz(m); with
                        // |z| < 2^acc_bits it is unreachable when e-1-ctz(m) >= acc_bits, else the whole tile
                        // gets the tie-to-even correction.  (m == 0: column past N, or a zero multiplier -> q = 0.)
                        const int tz = __ffs(d.m) - 1;
                        const bool in_range = (d.e >= 32 && d.e <= 62);
                        ok &= in_range ? 1 : 0;
                        any_tie |= (in_range && (d.e - 1 - tz < args.acc_bits)) ? 1 : 0;
                        p.m = d.m;
                        p.sh = d.e - 32;
                        if (d.e >= 1 && d.e <= 62) p.c = (long long)b * (long long)d.m + (1LL << (d.e - 1));
                    }
                } else if (MODE == GM_CARRIER) {
                    p.m = __float_as_int(pscale[j]);
                }
                cp[lane + 32 * j] = p;
                cb[lane + 32 * j] = b;
            }
            const bool all_ok = __all_sync(0xffffffffu, ok != 0);
            const bool some_tie = __any_sync(0xffffffffu, any_tie != 0);
            if (lane == 0) fast_flag[buf] = !all_ok ? 0 : (some_tie ? 2 : 1);
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(pfull_bar(buf));
        };
        {
            int t = tile_first;
            if (t < num_tiles) { load_params(t); store_params(0); }
            t += tile_step;
            if (t < num_tiles) { load_params(t); store_params(1); }
        }
        int it = 0;
        for (int tile = tile_first; tile < num_tiles; tile += tile_step, ++it) {
            const int buf = it & 1;
            const int nxt = tile + 2 * tile_step;
            if (nxt < num_tiles) load_params(nxt);
            ptx::mbar_wait(sfull_bar(buf), (uint32_t)((it >> 1) & 1));   // every epilogue warp is done with the constants of tile `it`
            if (nxt < num_tiles) store_params(buf);
        }
    } else {
        // ================= epilogue (warps 3..) =================
        // TMEM lane group is fixed by (warp % 4); the warps that share a lane group split the tile's columns.
        const int ew = warp - 3;
        const int lane_group = warp & 3;              // TMEM lanes [32*lane_group, +32) are accessible to this warp
        const int col_part = ew >> 2;                 // which 1/WPG of the tile's columns
        constexpr int CW = (MODE == GM_RQ_I16) ? 16 : 32;
        const uint32_t tempty_leader = PAIR ? ptx::mapa(tempty_bar(0), 0) : 0u;   // the leader's MMA warp owns both accumulators
        // output mode 3: residual box h (64 columns) of this warp's 32 rows for `tile`, by TMA into the residual tile
        auto issue_res = [&](int tile, int h) {
            if constexpr (RTMA) {
                const int bx = col_part * 2 + h;
                const uint32_t dst = res_base + (uint32_t)(bx * GEMM_BM * 128 + lane_group * 32 * 128);
                ptx::mbar_arrive_expect_tx(res_bar(ew, h), 32 * 128);
                ptx::tma_load_2d(dst, &tmap_res, res_bar(ew, h), (tile_n0(tile) + bx * 64) * 2, tile_m0(tile) + lane_group * 32);
            }
        };
        if (RTMA && lane == 0 && tile_first < num_tiles) { issue_res(tile_first, 0); issue_res(tile_first, 1); }
        int it = 0;
        for (int tile = tile_first; tile < num_tiles; tile += tile_step, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (uint32_t)((it >> 1) & 1);
            const int m0 = tile_m0(tile), n0 = tile_n0(tile);
            const ColParam* cp = col_params + as * BN;
            const int32_t* cb = col_bias + as * BN;

            const int row = m0 + lane_group * 32 + lane;
            const bool row_ok = row < args.M;
            constexpr int CPART = BN / WPG;
            const int c_begin = col_part * CPART;
            const int c_end = min(c_begin + CPART, args.N - n0);         // exclusive, may be <= c_begin
            const uint32_t t_row = tmem_base + ((uint32_t)(lane_group * 32) << 16) + (uint32_t)(as * BN);

            // boxes (128 bytes wide) per epilogue warp; 0: WPB = 2 warps of a lane group share one box (128-wide tiles)
            // and pair up through a 64-thread named barrier around its store (the even one of the two issues it)
            constexpr int BOX_COLS = 128 / (OUT_ES ? OUT_ES : 1);
            constexpr int NBOX_W = CPART / BOX_COLS;
            constexpr int WPB = NBOX_W > 0 ? 1 : BOX_COLS / CPART;     // warps per box
            static_assert(!TS || WPB <= 2, "at most two epilogue warps per 128-byte box");
            const int box_bar = 1 + lane_group * 2 + (col_part >> 1);   // named barrier of my box (WPB == 2)
            const bool box_leader = (NBOX_W > 0) || ((col_part & 1) == 0);
            auto wait_staging_free = [&]() {                             // my previous tile's TMA stores have read my staging rows
                if (lane == 0 && box_leader) ptx::tma_store_wait_read<0>();
                if (NBOX_W == 0) asm volatile("bar.sync %0, 64;" ::"r"(box_bar) : "memory");
                else __syncwarp();
            };
            const int trow = lane_group * 32 + lane;
            // OM == 2: the residual of my row segment -- ALL of it -- is requested now, asynchronously (cp.async, 16 bytes
            // each), into my own rows of the staging tile: one memory round trip per tile, overlapped with the wait for
            // the accumulator, instead of one exposed round trip per 16-column chunk.  The rows are private to this
            // thread (it consumed the previous tile's copy itself), so no synchronisation is involved.
            constexpr bool res_smem = RSM || RTMA;
            const uint32_t rsm_base = RTMA ? res_base : out_base;
            const uint32_t res_parity = (uint32_t)(it & 1);
            const int next_tile = tile + tile_step;
            // RTMA: box h must have landed before its first read; once its last chunk has been consumed the next tile's
            // box is requested into the same place (every lane's reads are complete: their values have been used)
            auto res_wait = [&](int h) { if constexpr (RTMA) ptx::mbar_wait(res_bar(ew, h), res_parity); };
            auto res_refill = [&](int h) {
                if constexpr (RTMA) {
                    __syncwarp();
                    if (lane == 0 && next_tile < num_tiles) { ptx::fence_proxy_async(); issue_res(next_tile, h); }
                }
            };
            uint32_t ra[CW], rb[CW], resa[CW / 2], resb[CW / 2];
            if (RSM) {
                const int16_t* rsrc = reinterpret_cast<const int16_t*>(args.residual) + (long long)(row_ok ? row : 0) * args.res_ld + n0;
#pragma unroll
                for (int j = 0; j < CPART / 8; ++j) {
                    const int c = c_begin + 8 * j;
                    const bool ok = row_ok && (n0 + c + 8 <= args.N);     // N % 8 == 0 on this path: all or nothing
                    cp_async_16(swz_addr(out_base, trow, c * 2), ok ? (const void*)(rsrc + c) : args.residual, ok ? 16 : 0);
                }
                cp_async_commit();
            } else if (!RTMA && MODE == GM_RQ_I16 && c_begin < c_end) {
                load_residual<CW>(args, row, row_ok, n0 + c_begin, resa);
            }

            ptx::mbar_wait(pfull_bar(as), aphase);                       // this tile's column constants are staged
            const int fast = fast_flag[as];
            ptx::mbar_wait(tfull_bar(as), aphase);
            ptx::tc_fence_after();

            const bool work = (c_begin < c_end) && !IVIT_DBG(args, 1);    // debug & 1: diagnostics, epilogue does nothing
            if (work) {
                tmem_ld_chunk<CW>(t_row + (uint32_t)c_begin, ra);
                ptx::tmem_ld_wait();
            }
            if (RSM) {
                cp_async_wait_all();                                     // my own copies: no barrier needed to read them back
                if (work) load_residual_smem<CW>(out_base, trow, c_begin, resa);
            } else if (TS) {
                wait_staging_free();
            }
            if (RTMA) {
                res_wait(0);
                if (work) load_residual_smem<CW>(rsm_base, trow, c_begin, resa);
            }
            if (work) {
#pragma unroll 1
                for (int c0 = c_begin; c0 < c_end; c0 += 2 * CW) {
                    const bool has1 = (c0 + CW) < c_end;
                    const bool has2 = (c0 + 2 * CW) < c_end;
                    // output mode 3 (full tiles only): chunks 0-3 of my column part lie in residual box 0, 4-7 in box 1
                    const int k = (c0 - c_begin) / CW;                   // even chunk index of this pair
                    if (has1) {
                        tmem_ld_chunk<CW>(t_row + (uint32_t)(c0 + CW), rb);
                        if (MODE == GM_RQ_I16) {
                            if (res_smem) load_residual_smem<CW>(rsm_base, trow, c0 + CW, resb);
                            else load_residual<CW>(args, row, row_ok, n0 + c0 + CW, resb);
                        }
                    }
                    epilogue_chunk<MODE, CW, TS>(ra, resa, cp + c0, cb + c0, args, row, row_ok, n0 + c0, fast,
                                                 out_base, lane_group * 32 + lane, c0);
                    ptx::tmem_ld_wait();
                    if (has1) {
                        if (has2) {
                            tmem_ld_chunk<CW>(t_row + (uint32_t)(c0 + 2 * CW), ra);
                            if (MODE == GM_RQ_I16) {
                                if (RTMA && k + 2 == 4) res_wait(1);      // the next pair starts box 1
                                if (res_smem) load_residual_smem<CW>(rsm_base, trow, c0 + 2 * CW, resa);
                                else load_residual<CW>(args, row, row_ok, n0 + c0 + 2 * CW, resa);
                            }
                        }
                        epilogue_chunk<MODE, CW, TS>(rb, resb, cp + c0 + CW, cb + c0 + CW, args, row, row_ok, n0 + c0 + CW, fast,
                                                     out_base, lane_group * 32 + lane, c0 + CW);
                        ptx::tmem_ld_wait();
                    }
                    if (RTMA && (k == 2 || k == 6)) res_refill(k == 2 ? 0 : 1);   // chunks 3 / 7 done: box consumed
                }
            } else if (RTMA) {
                // diagnostics path (no epilogue work): keep the residual pipeline's phases in step
                res_wait(1);
                res_refill(0);
                res_refill(1);
            }
            // release the accumulator back to the MMA warp and the constants buffer to the auxiliary warp
            ptx::tc_fence_before();
            if (TS) ptx::fence_proxy_async();                            // generic-proxy smem writes -> async proxy (TMA store)
            __syncwarp();
            if (lane == 0) {
                // constants buffer first: a warp can only reach tile it+2 (same buffer) after the MMA warp has seen
                // every warp's accumulator release for tile it, which this orders after the buffer release
                ptx::mbar_arrive(sfull_bar(as));
                if (PAIR) ptx::mbar_arrive_cluster(tempty_leader + 8u * as);
                else ptx::mbar_arrive(tempty_bar(as));
            }
            if (TS) {
                // my 32 rows x CPART columns of the staged tile -> global: one TMA store per 128-byte-wide box
                // (32-row boxes; coalesced, asynchronous, clips the M / N tails).  Each warp stores and later waits
                // for its own rows only, one tile later: no CTA-wide synchronisation around the store.
                if (NBOX_W == 0) asm volatile("bar.sync %0, 64;" ::"r"(box_bar) : "memory");
                if (lane == 0 && box_leader && !IVIT_DBG(args, 8)) {      // debug & 8: diagnostics, nothing is stored
                    const int r0 = m0 + lane_group * 32;
                    constexpr int NB = NBOX_W > 0 ? NBOX_W : 1;
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
                        const int bx = (NBOX_W > 0 ? col_part * NBOX_W : col_part / WPB) + b;
                        if (n0 + bx * BOX_COLS < args.N && r0 < args.M)
                            ptx::tma_store_2d(&tmap_out, out_base + (uint32_t)(bx * GEMM_BM * 128 + lane_group * 32 * 128),
                                              (n0 + bx * BOX_COLS) * OUT_ES, r0);   // byte-typed map
                    }
                    ptx::tma_store_commit();
                }
            }
        }
        if (TS && lane == 0) ptx::tma_store_wait<0>();               // all my stores complete before the CTA exits
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (PAIR) ptx::cluster_sync();                         // no CTA exits while its peer may still use its barriers / smem / TMEM
    if (warp == 1) {
        ptx::tc_fence_after();
        if (PAIR) ptx::tmem_dealloc_pair(tmem_base, TMEM_COLS);
        else ptx::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------
// Debug-only SIMT implementation (dp4a), selected by IVIT_GEMM_IMPL=simt.  It exists so that
// the tensor-core kernel can be cross-checked ON THE GPU at full problem sizes (the CPU oracle
// is too slow there); it is never the default path.
// ------------------------------------------------------------------------------------
__global__ void gemm_i8_simt_raw(const int8_t* __restrict__ A, long long lda, const int8_t* __restrict__ W,
                                 int M, int N, int K, int32_t* __restrict__ C, long long ldc) {
    __shared__ int32_t sa[32][9], sb[32][9];          // 32 rows x 32 bytes (+pad)
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 32 threads
    const int row = blockIdx.y * 32 + ty, col = blockIdx.x * 32 + tx;
    int32_t acc = 0;
    for (int k0 = 0; k0 < K; k0 += 32) {
        // each thread loads one byte-quad for A and B tiles
        if (tx < 8) {
            const int r = blockIdx.y * 32 + ty;
            int32_t v = 0;
            if (r < M) {
                const int kk = k0 + tx * 4;
                const int8_t* p = A + (long long)r * lda + kk;
                uint32_t b = 0;
                for (int u = 0; u < 4; ++u) if (kk + u < K) b |= ((uint32_t)(uint8_t)p[u]) << (8 * u);
                v = (int32_t)b;
            }
            sa[ty][tx] = v;
        } else if (tx < 16) {
            const int t = tx - 8;
            const int r = blockIdx.x * 32 + ty;
            int32_t v = 0;
            if (r < N) {
                const int kk = k0 + t * 4;
                const int8_t* p = W + (long long)r * K + kk;
                uint32_t b = 0;
                for (int u = 0; u < 4; ++u) if (kk + u < K) b |= ((uint32_t)(uint8_t)p[u]) << (8 * u);
                v = (int32_t)b;
            }
            sb[ty][t] = v;
        }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < 8; ++t) acc = __dp4a(sa[ty][t], sb[tx][t], acc);
        __syncthreads();
    }
    if (row < M && col < N) C[(long long)row * ldc + col] = acc;
}

__global__ void gemm_simt_epilogue(const int32_t* __restrict__ acc, GemmArgs a, int mode) {
    const long long n = (long long)a.M * a.N;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / a.N), c = (int)(i % a.N);
        const int32_t v = acc[i] + (a.bias ? a.bias[c] : 0);
        if (mode == GM_RAW_I32) { reinterpret_cast<int32_t*>(a.out)[(long long)r * a.out_ld + c] = v; continue; }
        if (mode == GM_CARRIER) { reinterpret_cast<float*>(a.out)[(long long)r * a.out_ld + c] = __fmul_rn(__int2float_rn(v), a.scale[c]); continue; }
        const ivit_dyadic_t d = a.me[c];
        int32_t q = requant32(v, d.m, d.e);
        if (mode == GM_RQ_I8) { reinterpret_cast<int8_t*>(a.out)[(long long)r * a.out_ld + c] = (int8_t)clamp_bits_rt(q, a.mode_bits); continue; }
        if (a.two_stage) q = requant32(clamp_bits_rt(q, a.mode_bits), a.rq2.m, a.rq2.e);
        long long s = q;
        if (a.residual) s += requant64((long long)reinterpret_cast<const int16_t*>(a.residual)[(long long)r * a.res_ld + c], a.rqr.m, a.rqr.e);
        reinterpret_cast<int16_t*>(a.out)[(long long)r * a.out_ld + c] = (int16_t)clamp_i64_bits(s, a.mode_bits);
    }
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2D uint8 tensor map: dims {inner = K bytes, outer = rows}, row pitch ld bytes, box {128, box_rows}, SWIZZLE_128B.
int make_tmap_2d_u8(ivit_ctx* ctx, CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_bytes,
                    uint32_t box_inner, uint32_t box_outer) {
    if (!ctx->encode_tiled) return fail(IVIT_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t gdim[2] = {inner, outer};
    cuuint64_t gstride[1] = {ld_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled)(
        tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), gdim, gstride, box, estr,
        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(IVIT_ECUDA, "cuTensorMapEncodeTiled failed (CUresult %d): inner=%llu outer=%llu ld=%llu",
                                       (int)r, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld_bytes);
    return IVIT_OK;
}

template <int BN, int STAGES, int MODE, int OM, bool PAIR>
static int launch_gemm(ivit_ctx* ctx, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const GemmArgs& ga,
                       cudaStream_t s, const CUtensorMap* tres = nullptr) {
    constexpr int OUT_ES = (OM == 0) ? 0 : (MODE == GM_RQ_I8 ? 1 : 2);
    constexpr int CS = PAIR ? 2 : 1;
    using S = GemmSmem<BN, STAGES, OUT_ES, PAIR, OM == 3>;
    const CUtensorMap& tr = tres ? *tres : ta;
    static_assert(S::TOTAL <= 227 * 1024, "shared memory budget");
    auto kern = gemm_i8_tcgen05_kernel<BN, STAGES, MODE, OM, PAIR>;
    static PerDevice attr_dev, clusters_dev;          // per instantiation and device
    int& attr_set = attr_dev[ctx->device];
    int& max_clusters = clusters_dev[ctx->device];
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(gemm_threads(MODE));
    cfg.dynamicSmemBytes = S::TOTAL;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (CS > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = CS; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
    }
    const int na_cluster = na;
    if (pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    if (!attr_set) {
        IVIT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
        if (CS > 1) {
            cfg.gridDim = dim3(CS * ctx->num_sms);
            cfg.numAttrs = na_cluster;                           // the occupancy query takes the cluster attribute only
            IVIT_CUDA_OK(cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg));
            cfg.numAttrs = na;
            if (max_clusters < 1) return fail(IVIT_ECUDA, "gemm: no co-resident cluster of %d CTAs fits", CS);
        }
        attr_set = 1;
    }
    const int tiles_m = (ga.M + GEMM_BM - 1) / GEMM_BM, tiles_n = (ga.N + BN - 1) / BN;
    const int ctiles = ((tiles_m + CS - 1) / CS) * tiles_n;
    const int cap = (CS > 1) ? max_clusters : ctx->num_sms;
    const int nclusters = ctiles < cap ? ctiles : cap;
    cfg.gridDim = dim3(nclusters * CS);
    IVIT_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, ta, tb, to, tr, ga));
    IVIT_LAUNCH_OK("gemm_i8_tcgen05_kernel");
    return IVIT_OK;
}

template <int MODE>
static int dispatch_bn(ivit_ctx* ctx, const int8_t* A, int64_t lda, const int8_t* W, const GemmArgs& ga, cudaStream_t s) {
    // wide tiles when N is large enough to fill them; 128-wide otherwise (less padding waste)
    const bool wide = (ga.N % 256 == 0) || ga.N >= 1024;
    // CTA pairs (cta_group::2, 256 x 256 MMA over two SMs) when there is enough work to keep every pair busy
    static const char* pair_env = getenv("IVIT_GEMM_PAIR");
    const bool pair = wide && ga.M >= 4 * GEMM_BM && !(pair_env && pair_env[0] == '0');
    CUtensorMap ta, tb, to;
    int rc = make_tmap_2d_u8(ctx, &ta, A, (uint64_t)ga.K, (uint64_t)ga.M, (uint64_t)lda, GEMM_BK, GEMM_BM);
    if (rc) return rc;
    // W box: each CTA of a pair loads half of the 256-row W tile
    rc = make_tmap_2d_u8(ctx, &tb, W, (uint64_t)ga.K, (uint64_t)ga.N, (uint64_t)ga.K, GEMM_BK, (wide && !pair) ? 256 : 128);
    if (rc) return rc;
    if constexpr (MODE == GM_RQ_I8 || MODE == GM_RQ_I16) {
        // Output staged through shared memory and written by TMA when the destination allows it
        // (16-byte aligned base and row pitch, full-width clamp); otherwise direct stores.
        constexpr int ES = (MODE == GM_RQ_I8) ? 1 : 2;
        const bool ts = ((uintptr_t)ga.out % 16 == 0) && ((ga.out_ld * ES) % 16 == 0) && ga.mode_bits == 8 * ES && ga.N % 8 == 0 &&
                        (!ga.residual || (((uintptr_t)ga.residual % 4) == 0 && ga.res_ld % 2 == 0));
        if constexpr (MODE == GM_RQ_I16) {
            // residual epilogue: direct 16-byte stores, residual prefetched through shared memory (cp.async)
            // (measured slower than the TMA-store form with register residuals, profiles/gemm_bench_r20.log: opt-in)
            static const char* rsm_env = getenv("IVIT_GEMM_RSM");
            const bool rsm = rsm_env && rsm_env[0] == '1' && ga.residual && ga.res_async && ((uintptr_t)ga.out % 16 == 0) &&
                             ((ga.out_ld * ES) % 16 == 0) && ga.mode_bits == 16 && ga.N % 8 == 0 && wide;
            if (rsm) {
                to = ta;
                if (pair) return launch_gemm<256, 4, MODE, 2, true>(ctx, ta, tb, to, ga, s);
                return launch_gemm<256, 3, MODE, 2, false>(ctx, ta, tb, to, ga, s);
            }
        }
        if (ts) {
            // byte-typed view of the output: inner dim = N*ES bytes, box = 128 bytes x 32 rows (one epilogue warp), 128B swizzle
            rc = make_tmap_2d_u8(ctx, &to, ga.out, (uint64_t)ga.N * ES, (uint64_t)ga.M, (uint64_t)ga.out_ld * ES, 128, GEMM_BM / 4);
            if (rc) return rc;
            if constexpr (MODE == GM_RQ_I16) {
                // residual tile by TMA (output mode 3): pairs, full 256-wide tiles, 16-byte aligned residual rows.  The extra
                // 64 KB tile leaves two operand stages, enough for the K <= 1024 shapes whose epilogue sets the time
                // (attn.proj); IVIT_GEMM_RTMA=0 disables, =2 takes it for every K.
                static const char* rt_env = getenv("IVIT_GEMM_RTMA");
                const int rt = rt_env ? atoi(rt_env) : 1;
                if (pair && rt && ga.residual && ga.N % 256 == 0 && ((uintptr_t)ga.residual % 16) == 0 && (ga.res_ld * 2) % 16 == 0 &&
                    (ga.K <= 1024 || rt == 2)) {
                    CUtensorMap tr;
                    rc = make_tmap_2d_u8(ctx, &tr, ga.residual, (uint64_t)ga.N * 2, (uint64_t)ga.M, (uint64_t)ga.res_ld * 2, 128, GEMM_BM / 4);
                    if (rc) return rc;
                    return launch_gemm<256, 2, MODE, 3, true>(ctx, ta, tb, to, ga, s, &tr);
                }
            }
            if (pair) return launch_gemm<256, (ES == 1 ? 5 : 4), MODE, 1, true>(ctx, ta, tb, to, ga, s);
            if (wide) return launch_gemm<256, 3, MODE, 1, false>(ctx, ta, tb, to, ga, s);
            return launch_gemm<128, 5, MODE, 1, false>(ctx, ta, tb, to, ga, s);
        }
    }
    to = ta;
    if (pair) return launch_gemm<256, 6, MODE, 0, true>(ctx, ta, tb, to, ga, s);
    if (wide) return launch_gemm<256, 4, MODE, 0, false>(ctx, ta, tb, to, ga, s);
    return launch_gemm<128, 6, MODE, 0, false>(ctx, ta, tb, to, ga, s);
}

}  // namespace ivit

using namespace ivit;

extern "C" int ivit_gemm_i8(ivit_ctx* ctx, const int8_t* A, int64_t lda, const int8_t* W, int64_t M, int64_t N,
                            int64_t K, const ivit_gemm_epilogue* epi, void* out, ivit_stream stream) {
    IVIT_REQUIRE(ctx && A && W && epi && out, "ivit_gemm_i8: null pointer");
    IVIT_REQUIRE(M > 0 && N > 0 && K > 0 && M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31), "ivit_gemm_i8: bad shape");
    IVIT_REQUIRE(K % 16 == 0 && lda % 16 == 0 && lda >= K, "ivit_gemm_i8: K and lda must be multiples of 16 (TMA row pitch), lda >= K");
    IVIT_REQUIRE(((uintptr_t)A % 16) == 0 && ((uintptr_t)W % 16) == 0, "ivit_gemm_i8: A and W must be 16-byte aligned");
    IVIT_REQUIRE(epi->out_ld >= N, "ivit_gemm_i8: out_ld < N");
    GemmArgs ga;
    ga.M = (int)M; ga.N = (int)N; ga.K = (int)K;
    ga.mode_bits = epi->bits;
    ga.bias = epi->bias; ga.me = epi->me;
    ga.residual = epi->residual; ga.res_ld = epi->res_ld;
    ga.two_stage = epi->two_stage;
    ga.rq2 = make_scalar_rq(epi->me2);                // operands are clamped 16-bit values
    ga.rqr = make_scalar_rq(epi->res_me);
    ga.acc_bits = (epi->acc_bits > 0 && epi->acc_bits <= 31) ? epi->acc_bits : 31;
    {
        const bool use2 = epi->two_stage != 0, user = epi->residual != nullptr;
        const bool uni = (!use2 || ga.rq2.kind == 1) && (!user || ga.rqr.kind == 1);
        const bool tie = (use2 && ga.rq2.tie) || (user && ga.rqr.tie);
        ga.scalar_mode = !uni ? 0 : (tie ? 2 : 1);
        // |RNE(z*m/2^e)| <= 2^15 * 2^31 / 2^e + 1 < 2^30 for e >= 18: the sum of the two terms fits 32 bits
        if (ga.scalar_mode == 1 && use2 && user && ga.rq2.e >= 18 && ga.rqr.e >= 18) ga.scalar_mode = 3;
        static const char* sm_env = getenv("IVIT_GEMM_SM4");
        if (ga.scalar_mode == 3 && ga.rq2.e <= 47 && ga.rqr.e <= 47 && !(sm_env && atoi(sm_env) == 0)) ga.scalar_mode = 4;
    }
    ga.scale = epi->scale; ga.out = out; ga.out_ld = epi->out_ld;
    ga.res_async = (epi->residual && ((uintptr_t)epi->residual % 16) == 0 && epi->res_ld % 8 == 0) ? 1 : 0;
    static const char* dbg_env = getenv("IVIT_GEMM_DEBUG");
    ga.debug = dbg_env ? atoi(dbg_env) : 0;
    int mode;
    switch (epi->mode) {
        case IVIT_EPI_RAW_I32:
            IVIT_REQUIRE(epi->out_dtype == IVIT_I32, "ivit_gemm_i8: RAW epilogue writes int32");
            mode = GM_RAW_I32; break;
        case IVIT_EPI_CARRIER:
            IVIT_REQUIRE(epi->out_dtype == IVIT_F32 && epi->scale, "ivit_gemm_i8: CARRIER epilogue needs scale[] and fp32 out");
            mode = GM_CARRIER; break;
        case IVIT_EPI_REQUANT:
            IVIT_REQUIRE(epi->me != nullptr, "ivit_gemm_i8: REQUANT epilogue needs me[N]");
            if (epi->bits == 8) {
                IVIT_REQUIRE(epi->out_dtype == IVIT_I8 && !epi->residual && !epi->two_stage,
                             "ivit_gemm_i8: 8-bit requant writes int8, no residual");
                mode = GM_RQ_I8;
            } else {
                IVIT_REQUIRE(epi->bits == 16 && epi->out_dtype == IVIT_I16, "ivit_gemm_i8: requant bits must be 8 or 16 (int8/int16 out)");
                IVIT_REQUIRE(!epi->residual || (epi->res_dtype == IVIT_I16 && epi->res_ld >= N), "ivit_gemm_i8: residual must be int16 with res_ld >= N");
                mode = GM_RQ_I16;
            }
            break;
        default:
            return fail(IVIT_EINVAL, "ivit_gemm_i8: unknown epilogue mode %d", epi->mode);
    }
    cudaStream_t s = st(stream);
    static const char* impl = getenv("IVIT_GEMM_IMPL");
    if (impl && impl[0] == 's') {                      // debug cross-check path, see gemm_i8_simt_raw
        int32_t* tmp = nullptr;
        IVIT_CUDA_OK(cudaMallocAsync(&tmp, sizeof(int32_t) * (size_t)M * (size_t)N, s));
        dim3 g((unsigned)((N + 31) / 32), (unsigned)((M + 31) / 32));
        gemm_i8_simt_raw<<<g, 1024, 0, s>>>(A, lda, W, (int)M, (int)N, (int)K, tmp, N);
        gemm_simt_epilogue<<<ctx->num_sms * 4, 256, 0, s>>>(tmp, ga, mode);
        IVIT_LAUNCH_OK("gemm_i8_simt");
        IVIT_CUDA_OK(cudaFreeAsync(tmp, s));
        return IVIT_OK;
    }
    switch (mode) {
        case GM_RAW_I32: return dispatch_bn<GM_RAW_I32>(ctx, A, lda, W, ga, s);
        case GM_CARRIER: return dispatch_bn<GM_CARRIER>(ctx, A, lda, W, ga, s);
        case GM_RQ_I8:   return dispatch_bn<GM_RQ_I8>(ctx, A, lda, W, ga, s);
        default:         return dispatch_bn<GM_RQ_I16>(ctx, A, lda, W, ga, s);
    }
}
