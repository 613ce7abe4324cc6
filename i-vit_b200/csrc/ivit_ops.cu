// Elementwise and row operators of the I-ViT integer path (sm_100a).
//
// All of these are HBM-bound integer/byte kernels: one warp per row for the row
// reductions (warp-shuffle max / sum over int32/int64 partials), 16-byte vector
// loads/stores on the narrow int8/int16 tensors, per-channel dyadic tables read through
// the read-only path.  Integer semantics: SURVEY.md Appendix A; each kernel cites the
// reference function it replaces.
#include <math.h>

#include "ivit_common.cuh"
#include "ivit_internal.h"

namespace ivit {

static inline int grid_for(int64_t work, int per_block, int num_sms, int waves = 8) {
    int64_t g = (work + per_block - 1) / per_block;
    int64_t cap = (int64_t)num_sms * waves;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// ------------------------------------------------------------------------------------
// batch_frexp   quant_utils.py:150-175 (+ :221-228)
// ------------------------------------------------------------------------------------
__global__ void dyadic_kernel(const float* __restrict__ s_in, int n, const float* __restrict__ s_out,
                              ivit_dyadic_t* __restrict__ out) {
    const double so = (double)s_out[0];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double r = (double)s_in[i] / so;
        int ex;
        const double mant = frexp(r, &ex);           // |mant| in [0.5, 1)
        long long m = llround(mant * 2147483648.0);  // half away from zero == Decimal ROUND_HALF_UP
        long long e = 31 - (long long)ex;
        if (m == 2147483648LL) { m = 1073741824LL; e -= 1; }   // normalise 2^31 -> fits int32
        if (m == 0) { e = 63; }                      // zero / denormal ratio: result is 0
        e = e > 63 ? 63 : (e < -1 ? -1 : e);
        out[i].m = (int32_t)m;
        out[i].e = (int32_t)e;
    }
}

// ------------------------------------------------------------------------------------
// SymmetricQuantFunction.forward   quant_utils.py:48, 90-92
// ------------------------------------------------------------------------------------
__global__ void quantize_f32_kernel(const float* __restrict__ x, int64_t n, const float* __restrict__ scale,
                                    int64_t ns, int64_t inner, int bits, int out_dtype, void* __restrict__ out) {
    const float hi = (float)(((long long)1 << (bits - 1)) - 1);   // fp32 bounds, as torch.clamp on fp32
    const float lo = -hi - 1.0f;
    const float inv0 = __fdiv_rn(1.0f, scale[0]);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float inv = (ns == 1) ? inv0 : __fdiv_rn(1.0f, scale[(i / inner) % ns]);
        float v = rintf(__fmul_rn(inv, x[i]));
        v = fminf(fmaxf(v, lo), hi);
        store_int(out, out_dtype, i, (int32_t)v);
    }
}

// Vectorised scalar-scale int8 specialisation (image -> int8, the only large instance).
__global__ void quantize_f32_i8_vec4(const float4* __restrict__ x, int64_t n4, const float* __restrict__ scale,
                                     char4* __restrict__ out) {
    const float inv = __fdiv_rn(1.0f, scale[0]);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = x[i];
        char4 o;
        o.x = (signed char)fminf(fmaxf(rintf(__fmul_rn(inv, v.x)), -128.f), 127.f);
        o.y = (signed char)fminf(fmaxf(rintf(__fmul_rn(inv, v.y)), -128.f), 127.f);
        o.z = (signed char)fminf(fmaxf(rintf(__fmul_rn(inv, v.z)), -128.f), 127.f);
        o.w = (signed char)fminf(fmaxf(rintf(__fmul_rn(inv, v.w)), -128.f), 127.f);
        out[i] = o;
    }
}

// ------------------------------------------------------------------------------------
// carrier <-> integer      quant_utils.py:220 ; the `* scaling_factor` of every operator
// ------------------------------------------------------------------------------------
__global__ void carrier_to_int_kernel(const float* __restrict__ x, int64_t n, int cols,
                                      const float* __restrict__ s, int s_len, int out_dtype,
                                      void* __restrict__ out) {
    const float lim_hi = out_dtype == IVIT_I8 ? 127.f : (out_dtype == IVIT_I16 ? 32767.f : 2147483520.f);
    const float lim_lo = out_dtype == IVIT_I8 ? -128.f : (out_dtype == IVIT_I16 ? -32768.f : -2147483648.f);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float sc = s[s_len == 1 ? 0 : (int)(i % cols)];
        float v = rintf(__fdiv_rn(x[i], sc));
        v = fminf(fmaxf(v, lim_lo), lim_hi);
        store_int(out, out_dtype, i, (int32_t)v);
    }
}

// fp64 carrier (IntLayerNorm outputs reach 2^30 and do not fit the 24-bit mantissa of an fp32
// carrier; the reference's modules produce an fp64 carrier there when their input is exact)
__global__ void carrier64_to_int_kernel(const double* __restrict__ x, int64_t n, int cols,
                                        const float* __restrict__ s, int s_len, int out_dtype,
                                        void* __restrict__ out) {
    const double lim_hi = out_dtype == IVIT_I8 ? 127.0 : (out_dtype == IVIT_I16 ? 32767.0 : 2147483647.0);
    const double lim_lo = out_dtype == IVIT_I8 ? -128.0 : (out_dtype == IVIT_I16 ? -32768.0 : -2147483648.0);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double sc = (double)s[s_len == 1 ? 0 : (int)(i % cols)];
        double v = rint(x[i] / sc);
        v = fmin(fmax(v, lim_lo), lim_hi);
        store_int(out, out_dtype, i, (int32_t)v);
    }
}

__global__ void int_to_carrier64_kernel(const void* __restrict__ q, int q_dtype, int64_t n, int cols,
                                        const float* __restrict__ s, int s_len, double* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double sc = (double)s[s_len == 1 ? 0 : (int)(i % cols)];
        out[i] = (double)load_int(q, q_dtype, i) * sc;
    }
}

__global__ void int_to_carrier_kernel(const void* __restrict__ q, int q_dtype, int64_t n, int cols,
                                      const float* __restrict__ s, int s_len, float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float sc = s[s_len == 1 ? 0 : (int)(i % cols)];
        out[i] = __fmul_rn(__int2float_rn(load_int(q, q_dtype, i)), sc);
    }
}

// ------------------------------------------------------------------------------------
// fixedpoint_mul.forward   quant_utils.py:192-253     (general form; the hot instances are
// fused into the GEMM / LayerNorm / GELU / attention kernels)
// ------------------------------------------------------------------------------------
__global__ void requant_kernel(const void* __restrict__ z, int z_dtype, int64_t rows, int cols,
                               const ivit_dyadic_t* __restrict__ me, int me_len,
                               const void* __restrict__ w, int w_dtype, int64_t w_rows,
                               const ivit_dyadic_t* __restrict__ me1, int me1_len,
                               int bits, int out_dtype, void* __restrict__ out) {
    const int64_t n = rows * (int64_t)cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % cols);
        const ivit_dyadic_t d = me[me_len == 1 ? 0 : c];
        long long o = requant64((long long)load_int(z, z_dtype, i), d.m, d.e);
        if (w != nullptr) {
            const ivit_dyadic_t d1 = me1[me1_len == 1 ? 0 : c];
            const int64_t wi = (w_rows == rows) ? i : (i % (w_rows * (int64_t)cols));   // periodic broadcast
            o += requant64((long long)load_int(w, w_dtype, wi), d1.m, d1.e);
        }
        store_int(out, out_dtype, i, clamp_i64_bits(o, bits));
    }
}

// ------------------------------------------------------------------------------------
// IntLayerNorm.forward   quant_modules.py:353-386   (+ optional fused QuantAct)
// One warp per row; the row is cached in registers (C <= 32*MAXV).
// ------------------------------------------------------------------------------------
template <int MAXV>
__global__ void __launch_bounds__(256)
layernorm_kernel(const void* __restrict__ x, int x_dtype, int64_t rows, int C,
                 const int32_t* __restrict__ bias_int, const ivit_dyadic_t* __restrict__ me,
                 int bits, int out_dtype, void* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int64_t warp0 = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * warps_per_block;
    for (int64_t row = warp0; row < rows; row += nwarps) {
        const int64_t base = row * (int64_t)C;
        int32_t v[MAXV];
        long long sum = 0;
#pragma unroll
        for (int j = 0; j < MAXV; ++j) {
            const int c = lane + 32 * j;
            v[j] = (c < C) ? load_int(x, x_dtype, base + c) : 0;
            sum += v[j];
        }
        sum = warp_sum_i64(sum);
        // mu = RNE(sum / C)          quant_modules.py:360 (mean, then round_ste)
        long long qd = sum / C, rem = sum % C;
        if (rem < 0) { qd -= 1; rem += C; }
        const long long twice = 2 * rem;
        if (twice > C || (twice == C && (qd & 1))) qd += 1;
        const int32_t mu = (int32_t)qd;
        // V = sum (q - mu)^2         :361-363
        unsigned long long V = 0;
#pragma unroll
        for (int j = 0; j < MAXV; ++j) {
            const int c = lane + 32 * j;
            const long long y = (c < C) ? (long long)v[j] - mu : 0;
            v[j] = (int32_t)y;
            V += (unsigned long long)(y * y);
        }
        V = (unsigned long long)warp_sum_i64((long long)V);
        // 10 integer Newton steps from 2^16     :366-370 (no early exit; V = 0 -> k = 64)
        const unsigned long long k = ln_isqrt10(V);
        const long long F = (long long)(2147483647ULL / k);      // :372
#pragma unroll
        for (int j = 0; j < MAXV; ++j) {
            const int c = lane + 32 * j;
            if (c < C) {
                long long o = (((long long)v[j] * F) >> 1);       // floor(y*F/2)   :373
                o += bias_int ? (long long)bias_int[c] : 0;       // :382
                o = o > 2147483647LL ? 2147483647LL : (o < -2147483648LL ? -2147483648LL : o);
                if (me != nullptr) {
                    const ivit_dyadic_t d = me[c];
                    o = clamp_i64_bits(requant64(o, d.m, d.e), bits);
                }
                store_int(out, out_dtype, base + c, (int32_t)o);
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// IntSoftmax.forward (Shiftmax)   quant_modules.py:469-497.  One warp per row.
// ------------------------------------------------------------------------------------
template <int MAXV>
__global__ void __launch_bounds__(256)
shiftmax_kernel(const void* __restrict__ q, int q_dtype, int64_t rows, int cols, int32_t x0,
                float inv_x0, int n, int out_bits, int out_dtype, void* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int64_t warp0 = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * wpb;
    const int sh = 31 - out_bits + 1;
    for (int64_t row = warp0; row < rows; row += nwarps) {
        const int64_t base = row * (int64_t)cols;
        int32_t v[MAXV];
        int32_t mx = INT32_MIN;
#pragma unroll
        for (int j = 0; j < MAXV; ++j) {
            const int c = lane + 32 * j;
            v[j] = (c < cols) ? load_int(q, q_dtype, base + c) : INT32_MIN;
            mx = v[j] > mx ? v[j] : mx;
        }
        mx = warp_max_i32(mx);
        long long E[MAXV];
        long long S = 0;
#pragma unroll
        for (int j = 0; j < MAXV; ++j) {
            const int c = lane + 32 * j;
            E[j] = (c < cols) ? shiftexp(v[j] - mx, x0, inv_x0, n) : 0;
            S += E[j];
        }
        S = warp_sum_i64(S);
        S = S > 2147483647LL ? 2147483647LL : S;                  // clamp_max_(2**31-1)  :491
        const long long F = 2147483647LL / S;                     // :492
#pragma unroll
        for (int j = 0; j < MAXV; ++j) {
            const int c = lane + 32 * j;
            if (c < cols) store_int(out, out_dtype, base + c, (int32_t)((E[j] * F) >> sh));   // :493
        }
    }
}

// ------------------------------------------------------------------------------------
// IntGELU.forward (ShiftGELU)   quant_modules.py:410-445   (+ optional fused scalar QuantAct)
// One warp per row, row streamed twice (max pass, then compute pass; the second read hits L1/L2).
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
shiftgelu_kernel(const void* __restrict__ q, int q_dtype, int64_t rows, int cols, int32_t x0,
                 float inv_x0, int n, const ivit_dyadic_t* __restrict__ me, int bits,
                 int out_dtype, void* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int64_t warp0 = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * wpb;
    ivit_dyadic_t d = {0, 0};
    if (me != nullptr) d = me[0];
    for (int64_t row = warp0; row < rows; row += nwarps) {
        const int64_t base = row * (int64_t)cols;
        int32_t mx = INT32_MIN;
        for (int c = lane; c < cols; c += 32) {
            const int32_t t = load_int(q, q_dtype, base + c);
            mx = t > mx ? t : mx;
        }
        mx = warp_max_i32(mx);
        const long long Em = shiftexp(-mx, x0, inv_x0, n);        // e^(-x_max)   :434
        for (int c = lane; c < cols; c += 32) {
            const int32_t x = load_int(q, q_dtype, base + c);
            const long long E = shiftexp(x - mx, x0, inv_x0, n);  // e^(x-x_max)  :432
            long long S = E + Em;
            S = S > 2147483647LL ? 2147483647LL : S;              // :437
            const long long F = 2147483647LL / S;                 // :438
            const long long sig = (E * F) >> (31 - 8 + 1);        // :439 (output_bit = 8)
            long long o = (long long)x * sig;                     // :442
            if (me != nullptr) o = clamp_i64_bits(requant64(o, d.m, d.e), bits);
            o = o > 2147483647LL ? 2147483647LL : (o < -2147483648LL ? -2147483648LL : o);
            store_int(out, out_dtype, base + c, (int32_t)o);
        }
    }
}

// ------------------------------------------------------------------------------------
// patch unfold (QuantConv2d with kernel == stride as a GEMM)   layers_quant.py:172-177,190-191
// ------------------------------------------------------------------------------------
__global__ void patchify_kernel(const int8_t* __restrict__ x, int B, int Cin, int H, int W, int p,
                                int8_t* __restrict__ out) {
    const int Hp = H / p, Wp = W / p;
    const int K = Cin * p * p;
    const int64_t n = (int64_t)B * Hp * Wp * K;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int kk = (int)(i % K);
        const int64_t r = i / K;
        const int v = kk % p, u = (kk / p) % p, c = kk / (p * p);
        const int j = (int)(r % Wp), ii = (int)((r / Wp) % Hp), b = (int)(r / ((int64_t)Wp * Hp));
        out[i] = x[(((int64_t)b * Cin + c) * H + ii * p + u) * W + j * p + v];
    }
}

// ------------------------------------------------------------------------------------
// DeiT stem glue: cls-token concatenation + position-embedding residual QuantAct
//   vit_quant.py:259-265:  x = cat(cls, patches) ; x = qact1(x, sf, qact_pos(pos_embed), sf_pos)
//   out[b,t,c] = clamp(RNE(z*m/2^e) + RNE(pos[t,c]*m1/2^e1)),  z = t == 0 ? cls[c] : pe[b,t-1,c]
// ------------------------------------------------------------------------------------
__global__ void embed_tokens_kernel(const int16_t* __restrict__ pe, const int32_t* __restrict__ cls,
                                    const int16_t* __restrict__ pos, int B, int n_tok, int C,
                                    ivit_dyadic_t me, ivit_dyadic_t me_res, int bits, int16_t* __restrict__ out) {
    const int64_t n = (int64_t)B * n_tok * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int64_t r = i / C;
        const int t = (int)(r % n_tok);
        const int64_t b = r / n_tok;
        const long long z = (t == 0) ? (long long)cls[c] : (long long)pe[(b * (n_tok - 1) + (t - 1)) * (int64_t)C + c];
        const long long o = requant64(z, me.m, me.e) + requant64((long long)pos[(int64_t)t * C + c], me_res.m, me_res.e);
        out[i] = (int16_t)clamp_i64_bits(o, bits);
    }
}

}  // namespace ivit

using namespace ivit;

// ====================================================================================
// C ABI
// ====================================================================================
extern "C" {

int ivit_dyadic(ivit_ctx* ctx, const float* s_in, int n, const float* s_out, ivit_dyadic_t* out,
                ivit_stream stream) {
    IVIT_REQUIRE(ctx && s_in && s_out && out && n > 0, "ivit_dyadic: null pointer or n <= 0");
    dyadic_kernel<<<(n + 127) / 128, 128, 0, st(stream)>>>(s_in, n, s_out, out);
    IVIT_LAUNCH_OK("dyadic_kernel");
    return IVIT_OK;
}

int ivit_quantize_f32(ivit_ctx* ctx, const float* x, int64_t n, const float* scale, int64_t ns,
                      int64_t inner, int bits, int out_dtype, void* out, ivit_stream stream) {
    IVIT_REQUIRE(ctx && x && scale && out && n > 0, "ivit_quantize_f32: null pointer or n <= 0");
    IVIT_REQUIRE(bits >= 2 && bits <= 32 && ns >= 1 && inner >= 1, "ivit_quantize_f32: bad bits/ns/inner");
    IVIT_REQUIRE((out_dtype == IVIT_I8 && bits <= 8) || (out_dtype == IVIT_I16 && bits <= 16) || out_dtype == IVIT_I32,
                 "ivit_quantize_f32: out_dtype cannot hold %d bits", bits);
    if (ns == 1 && bits == 8 && out_dtype == IVIT_I8 && (n % 4) == 0 &&
        ((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 4) == 0) {
        const int64_t n4 = n / 4;
        quantize_f32_i8_vec4<<<grid_for(n4, 256, ctx->num_sms), 256, 0, st(stream)>>>(
            (const float4*)x, n4, scale, (char4*)out);
    } else {
        quantize_f32_kernel<<<grid_for(n, 256, ctx->num_sms), 256, 0, st(stream)>>>(
            x, n, scale, ns, inner, bits, out_dtype, out);
    }
    IVIT_LAUNCH_OK("quantize_f32_kernel");
    return IVIT_OK;
}

int ivit_carrier_to_int(ivit_ctx* ctx, const void* x, int x_dtype, int64_t rows, int cols, const float* s,
                        int s_len, int out_dtype, void* out, ivit_stream stream) {
    IVIT_REQUIRE(x_dtype == IVIT_F32 || x_dtype == IVIT_F64, "ivit_carrier_to_int: x_dtype must be F32 or F64");
    IVIT_REQUIRE(ctx && x && s && out && rows > 0 && cols > 0, "ivit_carrier_to_int: bad arguments");
    IVIT_REQUIRE(s_len == 1 || s_len == cols, "ivit_carrier_to_int: s_len must be 1 or cols");
    IVIT_REQUIRE(out_dtype == IVIT_I8 || out_dtype == IVIT_I16 || out_dtype == IVIT_I32, "ivit_carrier_to_int: bad out_dtype");
    const int64_t n = rows * cols;
    if (x_dtype == IVIT_F32)
        carrier_to_int_kernel<<<grid_for(n, 256, ctx->num_sms), 256, 0, st(stream)>>>((const float*)x, n, cols, s, s_len, out_dtype, out);
    else
        carrier64_to_int_kernel<<<grid_for(n, 256, ctx->num_sms), 256, 0, st(stream)>>>((const double*)x, n, cols, s, s_len, out_dtype, out);
    IVIT_LAUNCH_OK("carrier_to_int_kernel");
    return IVIT_OK;
}

int ivit_int_to_carrier(ivit_ctx* ctx, const void* q, int q_dtype, int64_t rows, int cols,
                        const float* s, int s_len, int out_dtype, void* out, ivit_stream stream) {
    IVIT_REQUIRE(out_dtype == IVIT_F32 || out_dtype == IVIT_F64, "ivit_int_to_carrier: out_dtype must be F32 or F64");
    IVIT_REQUIRE(ctx && q && s && out && rows > 0 && cols > 0, "ivit_int_to_carrier: bad arguments");
    IVIT_REQUIRE(s_len == 1 || s_len == cols, "ivit_int_to_carrier: s_len must be 1 or cols");
    IVIT_REQUIRE(dtype_size(q_dtype) > 0 && q_dtype != IVIT_F32 && q_dtype != IVIT_F64, "ivit_int_to_carrier: bad q_dtype");
    const int64_t n = rows * cols;
    if (out_dtype == IVIT_F32)
        int_to_carrier_kernel<<<grid_for(n, 256, ctx->num_sms), 256, 0, st(stream)>>>(q, q_dtype, n, cols, s, s_len, (float*)out);
    else
        int_to_carrier64_kernel<<<grid_for(n, 256, ctx->num_sms), 256, 0, st(stream)>>>(q, q_dtype, n, cols, s, s_len, (double*)out);
    IVIT_LAUNCH_OK("int_to_carrier_kernel");
    return IVIT_OK;
}

int ivit_requant(ivit_ctx* ctx, const void* z, int z_dtype, int64_t rows, int cols,
                 const ivit_dyadic_t* me, int me_len, const void* w, int w_dtype, int64_t w_rows,
                 const ivit_dyadic_t* me1, int me1_len, int bits, int out_dtype, void* out,
                 ivit_stream stream) {
    IVIT_REQUIRE(ctx && z && me && out && rows > 0 && cols > 0, "ivit_requant: bad arguments");
    IVIT_REQUIRE(me_len == 1 || me_len == cols, "ivit_requant: me_len must be 1 or cols");
    IVIT_REQUIRE(bits == 4 || bits == 8 || bits == 16 || bits == 32,
                 "ivit_requant: bits must be 4, 8, 16 or 32 (quant_utils.py:247)");
    IVIT_REQUIRE((out_dtype == IVIT_I8 && bits <= 8) || (out_dtype == IVIT_I16 && bits <= 16) || out_dtype == IVIT_I32,
                 "ivit_requant: out_dtype cannot hold %d bits", bits);
    if (w != nullptr) {
        IVIT_REQUIRE(me1 && (me1_len == 1 || me1_len == cols) && w_rows >= 1 && rows % w_rows == 0,
                     "ivit_requant: residual needs me1 (len 1|cols) and w_rows dividing rows");
    }
    const int64_t n = rows * cols;
    requant_kernel<<<grid_for(n, 256, ctx->num_sms), 256, 0, st(stream)>>>(
        z, z_dtype, rows, cols, me, me_len, w, w_dtype, w_rows, me1, me1_len, bits, out_dtype, out);
    IVIT_LAUNCH_OK("requant_kernel");
    return IVIT_OK;
}

int ivit_layernorm(ivit_ctx* ctx, const void* x, int x_dtype, int64_t rows, int C,
                   const int32_t* bias_int, const ivit_dyadic_t* me, int bits, int out_dtype,
                   void* out, ivit_stream stream) {
    IVIT_REQUIRE(ctx && x && out && rows > 0 && C > 0, "ivit_layernorm: bad arguments");
    IVIT_REQUIRE(C <= 32 * 64, "ivit_layernorm: C=%d exceeds the supported 2048", C);
    IVIT_REQUIRE(x_dtype == IVIT_I8 || x_dtype == IVIT_I16 || x_dtype == IVIT_I32, "ivit_layernorm: bad x_dtype");
    if (me != nullptr) {
        IVIT_REQUIRE(bits == 8 || bits == 16 || bits == 32, "ivit_layernorm: fused QuantAct bits must be 8/16/32");
        IVIT_REQUIRE((out_dtype == IVIT_I8 && bits <= 8) || (out_dtype == IVIT_I16 && bits <= 16) || out_dtype == IVIT_I32,
                     "ivit_layernorm: out_dtype cannot hold %d bits", bits);
    } else {
        IVIT_REQUIRE(out_dtype == IVIT_I32, "ivit_layernorm: un-fused output is int32");
    }
    const int wpb = 8;
    const int grid = grid_for(rows, wpb, ctx->num_sms, 16);
#define LN_LAUNCH(MAXV) layernorm_kernel<MAXV><<<grid, wpb * 32, 0, st(stream)>>>(x, x_dtype, rows, C, bias_int, me, bits, out_dtype, out)
    if (C <= 32 * 4) LN_LAUNCH(4);
    else if (C <= 32 * 8) LN_LAUNCH(8);
    else if (C <= 32 * 12) LN_LAUNCH(12);
    else if (C <= 32 * 24) LN_LAUNCH(24);
    else if (C <= 32 * 32) LN_LAUNCH(32);
    else LN_LAUNCH(64);
#undef LN_LAUNCH
    IVIT_LAUNCH_OK("layernorm_kernel");
    return IVIT_OK;
}

// Supported x0 domain of the exact 64-bit shift-exponential (see shiftexp()): 1 <= -x0 <= 2^24, i.e. every scale the
// reference can produce (its scales are clamped at fp32 eps = 1.19e-7 -> x0 >= -8.4e6, quant_utils.py:63-67).
static int check_x0(const char* who, int32_t x0) {
    if (!(x0 <= -1 && x0 >= -(1 << 24)))
        return fail(IVIT_ENOTSUP, "%s: x0=%d outside [-2^24, -1]", who, x0);
    return IVIT_OK;
}

int ivit_shiftmax(ivit_ctx* ctx, const void* q, int q_dtype, int64_t rows, int cols, int32_t x0,
                  int n, int out_bits, int out_dtype, void* out, ivit_stream stream) {
    IVIT_REQUIRE(ctx && q && out && rows > 0 && cols > 0, "ivit_shiftmax: bad arguments");
    IVIT_REQUIRE(cols <= 32 * 32, "ivit_shiftmax: cols=%d exceeds the supported 1024", cols);
    IVIT_REQUIRE(q_dtype == IVIT_I8 || q_dtype == IVIT_I32, "ivit_shiftmax: q_dtype must be I8 or I32");
    IVIT_REQUIRE((out_bits == 16 && (out_dtype == IVIT_I16 || out_dtype == IVIT_I32)) ||
                 (out_bits == 8 && (out_dtype == IVIT_I8 || out_dtype == IVIT_I16 || out_dtype == IVIT_I32)),
                 "ivit_shiftmax: out_bits must be 8 or 16 with a dtype that holds [0, 2^(bits-1)]");
    IVIT_REQUIRE(n >= 1 && n <= 30, "ivit_shiftmax: bad n");
    int rc = check_x0("ivit_shiftmax", x0);
    if (rc) return rc;
    // E*F <= S*F <= 2^31-1 while the row sum S is below the 2^31-1 clamp, so P <= 127 (8 bit) / 32767 (16 bit) and the
    // narrow storage is safe.  Only for |x0| > 65536 (E(0) = |x0| << n can exceed 2^31: the clamp bites, F = 1) the
    // reference's own result leaves the nominal range (quant_modules.py:491-493): int32 storage is required there.
    if (x0 < -65536)
        IVIT_REQUIRE(out_dtype == IVIT_I32, "ivit_shiftmax: x0=%d (scale below 2^-16) needs an int32 output", x0);
    const float inv_x0 = 1.0f / (float)x0;
    const int wpb = 8;
    const int grid = grid_for(rows, wpb, ctx->num_sms, 16);
#define SM_LAUNCH(MAXV) shiftmax_kernel<MAXV><<<grid, wpb * 32, 0, st(stream)>>>(q, q_dtype, rows, cols, x0, inv_x0, n, out_bits, out_dtype, out)
    if (cols <= 64) SM_LAUNCH(2);
    else if (cols <= 224) SM_LAUNCH(7);
    else if (cols <= 512) SM_LAUNCH(16);
    else SM_LAUNCH(32);
#undef SM_LAUNCH
    IVIT_LAUNCH_OK("shiftmax_kernel");
    return IVIT_OK;
}

int ivit_shiftgelu(ivit_ctx* ctx, const void* q, int q_dtype, int64_t rows, int cols, int32_t x0,
                   int n, const ivit_dyadic_t* me, int bits, int out_dtype, void* out,
                   ivit_stream stream) {
    IVIT_REQUIRE(ctx && q && out && rows > 0 && cols > 0, "ivit_shiftgelu: bad arguments");
    IVIT_REQUIRE(q_dtype == IVIT_I8 || q_dtype == IVIT_I32, "ivit_shiftgelu: q_dtype must be I8 or I32");
    IVIT_REQUIRE(n >= 1 && n <= 30, "ivit_shiftgelu: bad n");
    if (me != nullptr) {
        IVIT_REQUIRE(bits == 8 || bits == 16, "ivit_shiftgelu: fused QuantAct bits must be 8 or 16");
        IVIT_REQUIRE((out_dtype == IVIT_I8 && bits <= 8) || out_dtype == IVIT_I16 || out_dtype == IVIT_I32,
                     "ivit_shiftgelu: out_dtype cannot hold %d bits", bits);
    } else {
        IVIT_REQUIRE(out_dtype == IVIT_I16 || out_dtype == IVIT_I32, "ivit_shiftgelu: un-fused output is int16/int32");
        // |q * sigmoid_int| <= 128 * 127 only while the 2^31-1 clamp of the exponential sum does not bite (|x0| < 256);
        // below that scale the reference's sigmoid_int exceeds 127 (quant_modules.py:437-439) and int32 is needed
        if (out_dtype == IVIT_I16) IVIT_REQUIRE(x0 >= -255, "ivit_shiftgelu: x0=%d needs an int32 output", x0);
    }
    int rc = check_x0("ivit_shiftgelu", x0);
    if (rc) return rc;
    const float inv_x0 = 1.0f / (float)x0;
    const int wpb = 8;
    const int grid = grid_for(rows, wpb, ctx->num_sms, 16);
    shiftgelu_kernel<<<grid, wpb * 32, 0, st(stream)>>>(q, q_dtype, rows, cols, x0, inv_x0, n, me, bits, out_dtype, out);
    IVIT_LAUNCH_OK("shiftgelu_kernel");
    return IVIT_OK;
}

int ivit_embed_tokens(ivit_ctx* ctx, const int16_t* pe, const int32_t* cls, const int16_t* pos, int B,
                      int n_tok, int C, ivit_dyadic_t me, ivit_dyadic_t me_res, int bits, int16_t* out,
                      ivit_stream stream) {
    IVIT_REQUIRE(ctx && pe && cls && pos && out && B > 0 && n_tok > 1 && C > 0, "ivit_embed_tokens: bad arguments");
    IVIT_REQUIRE(bits == 16 || bits == 8, "ivit_embed_tokens: bits must be 8 or 16");
    const int64_t n = (int64_t)B * n_tok * C;
    embed_tokens_kernel<<<grid_for(n, 256, ctx->num_sms), 256, 0, st(stream)>>>(pe, cls, pos, B, n_tok, C, me, me_res, bits, out);
    IVIT_LAUNCH_OK("embed_tokens_kernel");
    return IVIT_OK;
}

int ivit_patchify_i8(ivit_ctx* ctx, const int8_t* x, int B, int Cin, int H, int W, int p,
                     int8_t* out, ivit_stream stream) {
    IVIT_REQUIRE(ctx && x && out && B > 0 && Cin > 0 && p > 0 && H % p == 0 && W % p == 0,
                 "ivit_patchify_i8: bad arguments (H, W must be multiples of the patch size)");
    const int64_t n = (int64_t)B * Cin * H * W;
    patchify_kernel<<<grid_for(n, 256, ctx->num_sms), 256, 0, st(stream)>>>(x, B, Cin, H, W, p, out);
    IVIT_LAUNCH_OK("patchify_kernel");
    return IVIT_OK;
}

}  // extern "C"
