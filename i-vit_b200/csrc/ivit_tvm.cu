// TVM-semantics compatibility mode: the three integer row operators as the reference's
// deployment tree states them in Relay -- TVM_benchmark/models/layers.py:329-350 (quantized_layernorm), :353-369
// (shift_exp), :372-386 (quantized_softmax), :389-404 (quantized_gelu).  They differ numerically from the PyTorch
// operators (quant_modules.py) that the rest of this library reproduces: int32 arithmetic throughout (wrapping),
// truncating divisions, n = 16 / 23 without the half-bit (`(r >> 1) - x0`), no clamp of the exponential sums, an
// unsigned 32-bit variance, 8-bit softmax output by a wrapping cast.  For cross-checking against the authors' deployed
// numerics only; the engines never call these.
//
// Conventions where Relay leaves the result to the backend (stated in oracle/tvm_semantics.py, which these kernels
// equal bit for bit): a division by zero yields 0; a left shift by >= 32 yields 0, a shift by a negative count is
// taken as a shift by 0 ... both only reachable outside the value ranges the deployed model produces.
#include "ivit_common.cuh"
#include "ivit_internal.h"

namespace ivit {

__device__ __forceinline__ int32_t tvm_wrap_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
__device__ __forceinline__ int32_t tvm_wrap_sub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
__device__ __forceinline__ int32_t tvm_wrap_mul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
// relay divide on int32: truncation toward zero (C semantics); x / 0 := 0, INT_MIN / -1 := INT_MIN (wrap)
__device__ __forceinline__ int32_t tvm_div(int32_t a, int32_t b) {
    if (b == 0) return 0;
    if (b == -1) return (int32_t)(0u - (uint32_t)a);
    return a / b;
}
__device__ __forceinline__ int32_t tvm_shl(int32_t a, int32_t s) {
    if (s >= 32) return 0;
    if (s < 0) s = 0;
    return (int32_t)((uint32_t)a << s);
}

// layers.py:353-369
__device__ __forceinline__ int32_t tvm_shift_exp(int32_t d, int32_t x0, int32_t n) {
    d = tvm_wrap_sub(tvm_wrap_add(d, d >> 1), d >> 4);
    const int32_t lo = tvm_wrap_mul(n, x0);
    d = d > lo ? d : lo;
    const int32_t q = tvm_div(d, x0);
    const int32_t r = tvm_wrap_sub(d, tvm_wrap_mul(q, x0));
    return tvm_shl(tvm_wrap_sub(r >> 1, x0), tvm_wrap_sub(n, q));
}

__device__ __forceinline__ int32_t warp_max_s32(int32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ uint32_t warp_sum_u32(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One warp per row; the row is re-read from global memory (L1 / L2 resident) in each pass.
// layers.py:372-386
__global__ void tvm_softmax_kernel(const int32_t* __restrict__ x, int64_t rows, int cols, int32_t x0, int32_t n,
                                   int8_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t row = warp0; row < rows; row += nw) {
        const int32_t* xr = x + row * (int64_t)cols;
        int32_t mx = INT32_MIN;
        for (int c = lane; c < cols; c += 32) mx = max(mx, xr[c]);
        mx = warp_max_s32(mx);
        uint32_t s = 0;
        for (int c = lane; c < cols; c += 32) s += (uint32_t)tvm_shift_exp(tvm_wrap_sub(xr[c], mx), x0, n);
        const int32_t f = tvm_div(2147483647, (int32_t)warp_sum_u32(s));
        for (int c = lane; c < cols; c += 32) {
            const int32_t e = tvm_shift_exp(tvm_wrap_sub(xr[c], mx), x0, n);
            out[row * (int64_t)cols + c] = (int8_t)(uint8_t)((uint32_t)(tvm_wrap_mul(f, e) >> 24) & 0xffu);   // wrapping cast
        }
    }
}

// layers.py:389-404
__global__ void tvm_gelu_kernel(const int32_t* __restrict__ x, int64_t rows, int cols, int32_t x0, int32_t n,
                                int32_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t row = warp0; row < rows; row += nw) {
        const int32_t* xr = x + row * (int64_t)cols;
        int32_t mx = INT32_MIN;
        for (int c = lane; c < cols; c += 32) mx = max(mx, xr[c]);
        mx = warp_max_s32(mx);
        const int32_t e_max = tvm_shift_exp((int32_t)(0u - (uint32_t)mx), x0, n);
        for (int c = lane; c < cols; c += 32) {
            const int32_t pre = xr[c];
            const int32_t e = tvm_shift_exp(tvm_wrap_sub(pre, mx), x0, n);
            const int32_t sig = tvm_wrap_mul(tvm_div(2147483647, tvm_wrap_add(e, e_max)), e) >> 24;
            out[row * (int64_t)cols + c] = tvm_wrap_mul(pre, sig);
        }
    }
}

// layers.py:329-350
__global__ void tvm_layernorm_kernel(const int32_t* __restrict__ x, int64_t rows, int C, const int32_t* __restrict__ bias_int,
                                     int32_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t row = warp0; row < rows; row += nw) {
        const int32_t* xr = x + row * (int64_t)C;
        uint32_t s = 0;
        for (int c = lane; c < C; c += 32) s += (uint32_t)xr[c];
        const int32_t mean = tvm_div((int32_t)warp_sum_u32(s), C);
        uint32_t v = 0;
        for (int c = lane; c < C; c += 32) {
            const int32_t d = tvm_wrap_sub(xr[c], mean);
            v += (uint32_t)tvm_wrap_mul(d, d);
        }
        const uint32_t var = warp_sum_u32(v);
        uint32_t sd = 1u << 16;
#pragma unroll 1
        for (int i = 0; i < 10; ++i) sd = sd ? (sd + var / sd) / 2u : 0u;      // unsigned; sd == 0 cannot occur (>= 64)
        const int32_t factor = tvm_div(2147483647, (int32_t)sd);
        for (int c = lane; c < C; c += 32) {
            const int32_t d = tvm_wrap_sub(xr[c], mean);
            out[row * (int64_t)C + c] = tvm_wrap_add(tvm_div(tvm_wrap_mul(factor, d), 2), bias_int[c]);
        }
    }
}

static int tvm_grid(const ivit_ctx* ctx, int64_t rows) {
    const int64_t want = (rows + 7) / 8;
    const int64_t cap = (int64_t)ctx->num_sms * 8;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace ivit

using namespace ivit;

int ivit_tvm_softmax(ivit_ctx* ctx, const int32_t* x, int64_t rows, int cols, int32_t x0, int n, int8_t* out,
                     ivit_stream stream) {
    IVIT_REQUIRE(ctx && x && out && rows > 0 && cols > 0, "ivit_tvm_softmax: bad arguments");
    IVIT_REQUIRE(x0 < 0 && n >= 0 && n <= 31, "ivit_tvm_softmax: x0 must be negative, n in [0, 31]");
    tvm_softmax_kernel<<<tvm_grid(ctx, rows), 256, 0, st(stream)>>>(x, rows, cols, x0, n, out);
    IVIT_LAUNCH_OK("tvm_softmax_kernel");
    return IVIT_OK;
}

int ivit_tvm_gelu(ivit_ctx* ctx, const int32_t* x, int64_t rows, int cols, int32_t x0, int n, int32_t* out,
                  ivit_stream stream) {
    IVIT_REQUIRE(ctx && x && out && rows > 0 && cols > 0, "ivit_tvm_gelu: bad arguments");
    IVIT_REQUIRE(x0 < 0 && n >= 0 && n <= 31, "ivit_tvm_gelu: x0 must be negative, n in [0, 31]");
    tvm_gelu_kernel<<<tvm_grid(ctx, rows), 256, 0, st(stream)>>>(x, rows, cols, x0, n, out);
    IVIT_LAUNCH_OK("tvm_gelu_kernel");
    return IVIT_OK;
}

int ivit_tvm_layernorm(ivit_ctx* ctx, const int32_t* x, int64_t rows, int C, const int32_t* bias_int, int32_t* out,
                       ivit_stream stream) {
    IVIT_REQUIRE(ctx && x && bias_int && out && rows > 0 && C > 0, "ivit_tvm_layernorm: bad arguments");
    tvm_layernorm_kernel<<<tvm_grid(ctx, rows), 256, 0, st(stream)>>>(x, rows, C, bias_int, out);
    IVIT_LAUNCH_OK("tvm_layernorm_kernel");
    return IVIT_OK;
}
