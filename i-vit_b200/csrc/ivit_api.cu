// Context, error reporting and driver entry points of libivit_b200.so.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include "ivit_internal.h"

namespace ivit {

static thread_local char g_err[1024] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

bool pdl_enabled() {
    static const int on = [] { const char* e = getenv("IVIT_PDL"); return (e && e[0] == '1') ? 1 : 0; }();
    return on != 0;
}

int fail_cuda(cudaError_t e, const char* what) {
    snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return IVIT_ECUDA;
}

}  // namespace ivit

using namespace ivit;

extern "C" {

int ivit_version(void) { return 100; }

const char* ivit_last_error(void) { return g_err; }

int ivit_create(int device, ivit_ctx** out) {
    IVIT_REQUIRE(out != nullptr, "ivit_create: out is null");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(IVIT_ENODEV, "ivit_create: no CUDA device (%s); this library has no CPU path",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    IVIT_REQUIRE(device >= 0 && device < ndev, "ivit_create: device %d out of range [0, %d)", device, ndev);
    cudaDeviceProp prop;
    IVIT_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(IVIT_ENODEV, "ivit_create: device %d is sm_%d%d; this library is built for sm_100a only",
                    device, prop.major, prop.minor);
    int prev = -1;
    IVIT_CUDA_OK(cudaGetDevice(&prev));
    IVIT_CUDA_OK(cudaSetDevice(device));
    e = cudaFree(0);                                   // make sure the primary context exists
    if (prev != device) cudaSetDevice(prev);           // the caller's current device is not ours to change
    IVIT_CUDA_OK(e);
    ivit_ctx* c = new ivit_ctx();
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    c->cc_major = prop.major;
    c->cc_minor = prop.minor;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
        cudaGetLastError();
        delete c;
        return fail(IVIT_ECUDA, "ivit_create: cuTensorMapEncodeTiled driver entry point not found");
    }
    c->encode_tiled = fn;
    *out = c;
    return IVIT_OK;
}

int ivit_destroy(ivit_ctx* ctx) {
    delete ctx;
    return IVIT_OK;
}

int ivit_num_sms(ivit_ctx* ctx) { return ctx ? ctx->num_sms : 0; }

}  // extern "C"
