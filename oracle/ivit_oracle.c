/*
 * ivit_oracle.c -- CPU restatement of the I-ViT integer-only inference operators.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (i-vit_b200/) may link,
 * import or call this file.  It is used by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py as the CHECKER.
 *
 * Every function restates one reference function in exact integer arithmetic
 * (int64 / __int128), citing the reference file:line it follows (paths relative
 * to the reference checkout, zkkli/I-ViT @ 380ba99).  Parity is PINNED: the
 * golden vectors under tests/golden/ were produced by executing the reference's
 * own unmodified modules (tests/golden/make_golden.py) and this file reproduces
 * them bit-for-bit (tests/test_oracle_golden.py).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off, no fast-math: several
 * static quantities are defined by fp32 / fp64 IEEE operations).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define IVO_API __attribute__((visibility("default")))

typedef __int128 i128;

static inline int64_t clamp64(int64_t v, int64_t lo, int64_t hi) {
    return v < lo ? lo : (v > hi ? hi : v);
}

/* floor division for signed operands */
static inline int64_t floordiv64(int64_t a, int64_t b) {
    int64_t q = a / b, r = a % b;
    if (r != 0 && ((r < 0) != (b < 0))) q -= 1;
    return q;
}

/* round-half-to-even of num/den, den > 0 (torch.round of an exact rational) */
static inline int64_t div_rne(i128 num, i128 den) {
    i128 q = num / den, r = num % den;
    if (r < 0) { q -= 1; r += den; }             /* floor, 0 <= r < den */
    i128 twice = 2 * r;
    if (twice > den || (twice == den && (q & 1))) q += 1;
    return (int64_t)q;
}

/* RNE( p / 2^e ) for any integer e; exact. */
static inline int64_t shift_rne(i128 p, int e) {
    if (p == 0) return 0;
    if (e <= 0) {
        if (e < -62) return p > 0 ? INT64_MAX : INT64_MIN;
        i128 v = p << (-e);
        if (v > (i128)INT64_MAX) return INT64_MAX;
        if (v < (i128)INT64_MIN) return INT64_MIN;
        return (int64_t)v;
    }
    if (e >= 126) return 0;
    return div_rne(p, ((i128)1) << e);
}

/* ---------------------------------------------------------------------------
 * A.1  symmetric_linear_quantization_params   quant_utils.py:51-69
 *      n = 2^(b-1)-1 ; s = max(max(-min,max)/n, fp32 eps)   (fp32 arithmetic)
 * ------------------------------------------------------------------------- */
IVO_API float ivo_sym_scale(int bits, float min_val, float max_val) {
    float n = (float)((1LL << (bits - 1)) - 1);
    float mx = (-min_val > max_val) ? -min_val : max_val;
    float s = mx / n;
    const float eps = 1.1920928955078125e-07f;      /* torch.finfo(float32).eps */
    return s < eps ? eps : s;
}

/* ---------------------------------------------------------------------------
 * A.2  SymmetricQuantFunction.forward + linear_quantize
 *      quant_utils.py:48 (round(1./scale * input + zero_point)), :90-92 (clamp)
 *      fp32 reciprocal, fp32 multiply, RNE, clamp to [-n-1, n].
 *      scale has `ns` entries; element i uses scale[(i / inner) % ns]
 *      (inner=row length for per-row weight scales, ns=1 for a scalar).
 * ------------------------------------------------------------------------- */
IVO_API void ivo_quantize_f32(const float* x, int64_t n_elem, const float* scale,
                              int64_t ns, int64_t inner, int bits, int64_t* out) {
    /* torch.clamp(x, -n-1, n) on an fp32 tensor: the bounds are applied in fp32,
       so for bits=32 the upper bound 2^31-1 becomes 2^31 (quant_utils.py:92). */
    double nd = (double)((1LL << (bits - 1)) - 1);
    float hi = (float)nd, lo = (float)(-nd - 1.0);
    for (int64_t i = 0; i < n_elem; ++i) {
        float s = scale[ns == 1 ? 0 : (i / inner) % ns];
        volatile float inv = 1.0f / s;
        volatile float v = inv * x[i];
        float r = rintf(v + 0.0f);
        if (r < lo) r = lo;
        if (r > hi) r = hi;
        out[i] = (int64_t)r;
    }
}

/* ---------------------------------------------------------------------------
 * batch_frexp   quant_utils.py:150-175   (+ the ratio of fixedpoint_mul :221-223)
 *   r = fp64(s_in) / fp64(fp32(s_out)); (mant, ex) = frexp(r);
 *   m = round_half_away(mant * 2^31); e = 31 - ex.
 *   Returned un-normalised (|m| in [2^30, 2^31]) exactly as the reference.
 * ------------------------------------------------------------------------- */
IVO_API void ivo_dyadic(const float* s_in, int64_t n, float s_out, int64_t* m, int64_t* e) {
    for (int64_t i = 0; i < n; ++i) {
        double r = (double)s_in[i] / (double)s_out;
        int ex;
        double mant = frexp(r, &ex);
        double sc = mant * 2147483648.0;             /* exact */
        m[i] = (int64_t)llround(sc);                 /* Decimal ROUND_HALF_UP == half away from 0 */
        e[i] = 31 - (int64_t)ex;
    }
}

/* ---------------------------------------------------------------------------
 * A.4  fixedpoint_mul.forward   quant_utils.py:192-253
 *   out = clamp( RNE(z*m / 2^e) [+ RNE(w*m1 / 2^e1)] , -n-1, n )
 *   z:[rows, cols]; m/e have me_len (1 or cols) entries; the residual uses
 *   m1/e1 (me1_len entries) and w with wrows rows broadcast along rows
 *   (wrows divides rows: wrows == rows, or one sequence's worth of rows for pos_embed,
 *   vit_quant.py:264-265, broadcast over the batch).
 *   Exact 64x64->128-bit product (the reference forms the product in fp64,
 *   quant_utils.py:229; `*n_fp64_diff` counts elements where that fp64
 *   evaluation would differ from the exact one).
 * ------------------------------------------------------------------------- */
static inline int64_t requant_fp64(int64_t z, int64_t m, int64_t e) {
    double o = (double)z * (double)m;
    o = nearbyint(o / pow(2.0, (double)e));
    if (o > 9.2e18) return INT64_MAX;
    if (o < -9.2e18) return INT64_MIN;
    return (int64_t)o;
}

IVO_API void ivo_requant(const int64_t* z, int64_t rows, int64_t cols,
                         const int64_t* m, const int64_t* e, int64_t me_len,
                         const int64_t* w, int64_t wrows,
                         const int64_t* m1, const int64_t* e1, int64_t me1_len,
                         int bits, int64_t* out, int64_t* n_fp64_diff) {
    int64_t n = (1LL << (bits - 1)) - 1;
    int do_clamp = (bits == 4 || bits == 8 || bits == 16 || bits == 32);
    int64_t diff = 0;
    for (int64_t r = 0; r < rows; ++r)
        for (int64_t c = 0; c < cols; ++c) {
            int64_t i = r * cols + c;
            int64_t mm = m[me_len == 1 ? 0 : c], ee = e[me_len == 1 ? 0 : c];
            int64_t o = shift_rne((i128)z[i] * (i128)mm, (int)ee);
            if (o != requant_fp64(z[i], mm, ee)) diff++;
            if (w) {
                int64_t wi = w[(r % wrows) * cols + c];              /* periodic broadcast */
                int64_t mm1 = m1[me1_len == 1 ? 0 : c], ee1 = e1[me1_len == 1 ? 0 : c];
                int64_t o1 = shift_rne((i128)wi * (i128)mm1, (int)ee1);
                if (o1 != requant_fp64(wi, mm1, ee1)) diff++;
                i128 s = (i128)o + (i128)o1;
                o = s > (i128)INT64_MAX ? INT64_MAX : (s < (i128)INT64_MIN ? INT64_MIN : (int64_t)s);
            }
            out[i] = do_clamp ? clamp64(o, -n - 1, n) : o;
        }
    if (n_fp64_diff) *n_fp64_diff = diff;
}

/* ---------------------------------------------------------------------------
 * A.3  QuantLinear / QuantConv2d / QuantMatMul integer contraction
 *      quant_modules.py:93-97, 224-228, 325-330
 *   acc[i,j] = sum_k a[i,k]*w[j,k] (+ b[j])           (A [M,K], W [N,K])
 * ------------------------------------------------------------------------- */
IVO_API void ivo_gemm_nt(const int32_t* a, const int32_t* w, const int64_t* bias,
                         int64_t M, int64_t N, int64_t K, int64_t* out) {
    for (int64_t i = 0; i < M; ++i) {
        const int32_t* ar = a + i * K;
        for (int64_t j = 0; j < N; ++j) {
            const int32_t* wr = w + j * K;
            int64_t acc = 0;
            for (int64_t k = 0; k < K; ++k) acc += (int64_t)ar[k] * (int64_t)wr[k];
            out[i * N + j] = acc + (bias ? bias[j] : 0);
        }
    }
}

/* int8 fast path of the same contraction (used for whole-model oracles and as the CPU
 * baseline of bench.py): rows are split over `ivo_set_threads` host threads, the dot
 * product is auto-vectorised per ISA (function multi-versioning, safe on any x86-64). */
static int g_threads = 1;
IVO_API void ivo_set_threads(int n) { g_threads = n < 1 ? 1 : (n > 256 ? 256 : n); }
IVO_API int ivo_get_threads(void) { return g_threads; }

#if defined(__x86_64__)
__attribute__((target_clones("arch=x86-64-v4", "avx2", "default")))
#endif
static void gemm_rows_i8(const int8_t* a, const int8_t* w, const int32_t* bias,
                         int64_t r0, int64_t r1, int64_t N, int64_t K, int32_t* out) {
    for (int64_t i = r0; i < r1; ++i) {
        const int8_t* ar = a + i * K;
        for (int64_t j = 0; j < N; ++j) {
            const int8_t* wr = w + j * K;
            int32_t acc = 0;
            for (int64_t k = 0; k < K; ++k) acc += (int32_t)ar[k] * (int32_t)wr[k];
            out[i * N + j] = acc + (bias ? bias[j] : 0);
        }
    }
}

typedef struct {
    const int8_t* a; const int8_t* w; const int32_t* bias;
    int64_t r0, r1, N, K; int32_t* out;
} gemm_job;

static void* gemm_worker(void* p) {
    gemm_job* j = (gemm_job*)p;
    gemm_rows_i8(j->a, j->w, j->bias, j->r0, j->r1, j->N, j->K, j->out);
    return NULL;
}

IVO_API void ivo_gemm_nt_i8(const int8_t* a, const int8_t* w, const int32_t* bias,
                            int64_t M, int64_t N, int64_t K, int32_t* out) {
    int T = g_threads;
    if (T > M) T = (int)M;
    if (T <= 1 || M * N * K < (1LL << 22)) {
        gemm_rows_i8(a, w, bias, 0, M, N, K, out);
        return;
    }
    pthread_t th[256];
    gemm_job jobs[256];
    for (int t = 0; t < T; ++t) {
        jobs[t] = (gemm_job){a, w, bias, M * t / T, M * (t + 1) / T, N, K, out};
        pthread_create(&th[t], NULL, gemm_worker, &jobs[t]);
    }
    for (int t = 0; t < T; ++t) pthread_join(th[t], NULL);
}

/* ---------------------------------------------------------------------------
 * int_exp_shift   quant_modules.py:410-423 (IntGELU) == :469-481 (IntSoftmax)
 *   t = d + floor(d/2) - floor(d/16); t = max(t, n*x0); k = floor(t/x0);
 *   r = t - x0*k; E = max(floor((r/2 - x0) * 2^(n-k)), 0)
 *   (r/2 - x0)*2^(n-k) == (r - 2*x0) * 2^(n-k-1): exact, floor handles n-k-1 = -1.
 *   Saturated at 2^100 (only reachable for degenerate scales, x0 ~ -1).
 * ------------------------------------------------------------------------- */
static inline i128 shiftexp(int64_t d, int64_t x0, int n) {
    int64_t t = d + floordiv64(d, 2) - floordiv64(d, 16);
    int64_t lim = (int64_t)n * x0;
    if (t < lim) t = lim;
    int64_t k = floordiv64(t, x0);
    int64_t r = t - x0 * k;
    i128 base = (i128)r - 2 * (i128)x0;             /* > 0 */
    int64_t sh = (int64_t)n - k - 1;
    i128 E;
    if (sh >= 0) E = (sh > 100) ? (((i128)1) << 100) : (base << sh);
    else E = base >> 1;                              /* sh == -1 (k == n): floor(base/2) */
    if (E < 0) E = 0;
    if (E > (((i128)1) << 100)) E = ((i128)1) << 100;
    return E;
}

/* x0 = floor(-1 / s) in fp32   quant_modules.py:414 / :473 */
IVO_API int64_t ivo_x0(float s) {
    volatile float q = -1.0f / s;
    return (int64_t)floorf(q);
}
/* IntGELU's sigmoid scale: fp32(s * 1.702)   quant_modules.py:427 */
IVO_API float ivo_gelu_sig_scale(float s) {
    volatile float r = s * 1.702f;
    return r;
}

/* ---------------------------------------------------------------------------
 * A.5  IntSoftmax.forward (Shiftmax)   quant_modules.py:483-497
 *   row-wise over the last dim; n = 15; out bits b (16 DeiT / 8 Swin)
 * ------------------------------------------------------------------------- */
IVO_API void ivo_shiftmax(const int64_t* q, int64_t rows, int64_t cols, int64_t x0,
                          int n, int out_bits, int64_t* out) {
    i128* E = (i128*)malloc(sizeof(i128) * (size_t)cols);
    for (int64_t r = 0; r < rows; ++r) {
        const int64_t* x = q + r * cols;
        int64_t mx = x[0];
        for (int64_t c = 1; c < cols; ++c) if (x[c] > mx) mx = x[c];
        i128 S = 0;
        for (int64_t c = 0; c < cols; ++c) { E[c] = shiftexp(x[c] - mx, x0, n); S += E[c]; }
        const i128 cap = 2147483647;                 /* clamp_max_(2**31-1)  :491 */
        if (S > cap) S = cap;
        i128 F = cap / S;                            /* floor((2**31-1)/sum) :492 */
        int sh = 31 - out_bits + 1;
        for (int64_t c = 0; c < cols; ++c) out[r * cols + c] = (int64_t)((E[c] * F) >> sh);
    }
    free(E);
}

/* ---------------------------------------------------------------------------
 * A.6  IntGELU.forward (ShiftGELU)   quant_modules.py:425-445
 *   x0 is derived from fp32(s*1.702) (ivo_gelu_sig_scale + ivo_x0); n = 23; b = 8
 *   out = q * sigma   (scale s/128)
 * ------------------------------------------------------------------------- */
IVO_API void ivo_shiftgelu(const int64_t* q, int64_t rows, int64_t cols, int64_t x0,
                           int n, int out_bits, int64_t* out) {
    for (int64_t r = 0; r < rows; ++r) {
        const int64_t* x = q + r * cols;
        int64_t mx = x[0];
        for (int64_t c = 1; c < cols; ++c) if (x[c] > mx) mx = x[c];
        i128 Em = shiftexp(-mx, x0, n);              /* e^(-x_max)   :434 */
        const i128 cap = 2147483647;
        int sh = 31 - out_bits + 1;
        for (int64_t c = 0; c < cols; ++c) {
            i128 E = shiftexp(x[c] - mx, x0, n);     /* e^(x-x_max)  :432 */
            i128 S = E + Em;
            if (S > cap) S = cap;                    /* :437 */
            i128 F = cap / S;                        /* :438 */
            i128 sig = (E * F) >> sh;                /* :439 */
            i128 o = (i128)x[c] * sig;               /* :442 */
            out[r * cols + c] = o > (i128)INT64_MAX ? INT64_MAX : (o < (i128)INT64_MIN ? INT64_MIN : (int64_t)o);
        }
    }
}

/* ---------------------------------------------------------------------------
 * A.7  IntLayerNorm.forward   quant_modules.py:353-386
 *   mu = RNE(sum/C) (x_int.mean then round_ste :360); y = q - mu; V = sum y^2;
 *   k = 2^16; 10x: k = floor((k + floor(V/k))/2) (:366-370);
 *   F = floor((2^31-1)/k); y' = floor(y*F/2); out = y' + b_q[c]
 * ------------------------------------------------------------------------- */
IVO_API void ivo_layernorm(const int64_t* q, int64_t rows, int64_t cols,
                           const int64_t* bias_int, int64_t* out) {
    for (int64_t r = 0; r < rows; ++r) {
        const int64_t* x = q + r * cols;
        i128 sum = 0;
        for (int64_t c = 0; c < cols; ++c) sum += x[c];
        int64_t mu = div_rne(sum, (i128)cols);
        i128 V = 0;
        for (int64_t c = 0; c < cols; ++c) { i128 y = (i128)x[c] - mu; V += y * y; }
        i128 k = 65536;
        for (int it = 0; it < 10; ++it) k = (k + V / k) / 2;
        i128 F = ((i128)2147483647) / k;
        for (int64_t c = 0; c < cols; ++c) {
            i128 y = (i128)x[c] - mu;
            i128 p = y * F;
            i128 h = p >> 1;                         /* floor(p/2): arithmetic shift */
            out[r * cols + c] = (int64_t)(h + (bias_int ? bias_int[c] : 0));
        }
    }
}

/* ---------------------------------------------------------------------------
 * Swin tail: AdaptiveAvgPool1d on dequantised values followed by QuantAct
 * (swin_quant.py:554-555) -- integer reading z = RNE(sum_q / L).
 * ------------------------------------------------------------------------- */
IVO_API void ivo_avgpool_rne(const int64_t* q, int64_t batch, int64_t L, int64_t C, int64_t* out) {
    for (int64_t b = 0; b < batch; ++b)
        for (int64_t c = 0; c < C; ++c) {
            i128 s = 0;
            for (int64_t l = 0; l < L; ++l) s += q[(b * L + l) * C + c];
            out[b * C + c] = div_rne(s, (i128)L);
        }
}

IVO_API int ivo_version(void) { return 1; }
