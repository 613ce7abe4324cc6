// Host-side internals shared by the translation units of libivit_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/ivit_b200.h"

struct ivit_ctx {
    int device = 0;
    int num_sms = 0;
    int cc_major = 0, cc_minor = 0;
    size_t smem_optin = 0;
    void* encode_tiled = nullptr;   // cuTensorMapEncodeTiled (driver entry point)
};

namespace ivit {
// Per-device one-time state of a kernel instantiation (cudaFuncSetAttribute and the occupancy queries are per device:
// one process may hold contexts on several GPUs).  Indexed by ivit_ctx::device.
constexpr int kMaxDevices = 64;
struct PerDevice {
    int v[kMaxDevices] = {};
    int& operator[](int dev) { return v[(unsigned)dev % kMaxDevices]; }
};
// Launch with (IVIT_PDL=1) or without the programmatic-stream-serialization attribute: the grid may be scheduled while
// its predecessor in the stream drains; every kernel launched this way starts with ptx::grid_dep_wait().
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
int fail(int code, const char* fmt, ...);
int fail_cuda(cudaError_t e, const char* what);
inline cudaStream_t st(ivit_stream s) { return reinterpret_cast<cudaStream_t>(s); }
// tcgen05 attention (ivit_attn_tc.cu); preconditions are checked by ivit_attention_i8
int launch_attention_tc(ivit_ctx* ctx, const int8_t* qkv, const ivit_attn_params* ap, long long half_s, long long half_o,
                        int8_t* out, cudaStream_t s);
// pipelined tcgen05 attention, one persistent CTA per SM (ivit_attn_pipe.cu); same preconditions, 49 <= n_tok <= 224
int launch_attention_pipe(ivit_ctx* ctx, const int8_t* qkv, const ivit_attn_params* ap, long long half_s, long long half_o,
                          int8_t* out, cudaStream_t s);
inline int dtype_size(int dt) {
    switch (dt) {
        case IVIT_I8: case IVIT_U8: return 1;
        case IVIT_I16: return 2;
        case IVIT_I32: case IVIT_F32: return 4;
        case IVIT_F64: return 8;
        default: return 0;
    }
}
}  // namespace ivit

#define IVIT_REQUIRE(cond, ...)                                   \
    do {                                                          \
        if (!(cond)) return ivit::fail(IVIT_EINVAL, __VA_ARGS__); \
    } while (0)
#define IVIT_CUDA_OK(expr)                                        \
    do {                                                          \
        cudaError_t _e = (expr);                                  \
        if (_e != cudaSuccess) return ivit::fail_cuda(_e, #expr); \
    } while (0)
#define IVIT_LAUNCH_OK(name)                                       \
    do {                                                           \
        cudaError_t _e = cudaPeekAtLastError();                    \
        if (_e != cudaSuccess) return ivit::fail_cuda(_e, name);   \
    } while (0)
