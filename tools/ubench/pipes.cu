// Issue-rate micro-benchmark of the integer instructions the attention softmax warps are made of (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/pipes tools/ubench/pipes.cu && tools/ubench/pipes
// One CTA; W warps per scheduler (CTA of 128*W threads); every warp runs ITER iterations of 8 independent chains of one
// instruction; prints cycles per warp-instruction per scheduler (1.0 = full rate, 2.0 = half rate, ...).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 2048

template <int OP>
__device__ __forceinline__ void step(uint32_t (&a)[8], uint32_t b, uint32_t c, unsigned long long h, const uint32_t* sm) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if constexpr (OP == 0) a[i] = __umulhi(a[i], b);                                   // IMAD.HI.U32
        if constexpr (OP == 1) a[i] = (uint32_t)(((long long)(int32_t)a[i] * (long long)(int32_t)b + (long long)h) >> 32);   // IMAD.HI + 64-bit addend
        if constexpr (OP == 2) a[i] = a[i] * b + c;                                        // IMAD
        if constexpr (OP == 3) a[i] = (uint32_t)((int32_t)a[i] >> (b & 31));               // SHF.R.S32.HI
        if constexpr (OP == 4) a[i] = __byte_perm(a[i], b, 0x6420);                        // PRMT
        if constexpr (OP == 5) asm volatile("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));   // I2IP
        if constexpr (OP == 6) asm volatile("dp4a.s32.s32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));            // IDP.4A
        if constexpr (OP == 7) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a[i]) : "r"((a[i] & 0x7c) + c));          // LDS (address from the previous load)
        if constexpr (OP == 8) a[i] = (a[i] & b) ^ c;                                      // LOP3
        if constexpr (OP == 9) a[i] = a[i] + b + c;                                        // IADD3
        if constexpr (OP == 10) a[i] = __vmaxs2(a[i], b);                                  // VIMNMX-type SIMD max
        if constexpr (OP == 11) a[i] = __umulhi(a[i], b) * 65536u + c;                     // IMAD.HI.U32 + IMAD (pass 3 pair)
        if constexpr (OP == 12) a[i] = (uint32_t)((int32_t)(uint32_t)(((long long)(int32_t)a[i] * (long long)(int32_t)b + (long long)h) >> 32) >> (c & 31));  // IMAD.HI + SHF (pass 1 pair)
        if constexpr (OP == 13) a[i] = (uint32_t)(((unsigned long long)a[i] * b) >> 16);   // 64-bit product, funnel shift
        if constexpr (OP == 14) a[i] = __umul24(a[i], b) + c;                              // 24-bit multiply
        if constexpr (OP == 15) a[i] = (uint32_t)__float2uint_rz(__uint2float_rz(a[i]) * __uint_as_float(b));   // I2F, FMUL, F2I
        if constexpr (OP == 16) asm volatile("dp2a.lo.s32.s32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));          // IDP.2A
        if constexpr (OP == 17) a[i] = (uint32_t)(((long long)(int32_t)a[i] * (long long)(int32_t)b) >> 1);            // IMAD.WIDE + SHF.R.U64 (LayerNorm y F / 2)
        if constexpr (OP == 18) a[i] = __float_as_uint(__fmaf_rz(__uint_as_float(a[i]), __uint_as_float(b), __uint_as_float(c)));   // FFMA
        if constexpr (OP == 19) a[i] = (uint32_t)max(min((int32_t)a[i] + (int32_t)b, 127), -128);                      // IADD + clamp (VIMNMX3)
    }
}

template <int OP>
__global__ void bench(uint32_t* out, long long* cyc, uint32_t b, uint32_t c, unsigned long long h) {
    __shared__ uint32_t sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (i * 4) & 0x7c;
    __syncthreads();
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 2654435761u + i * 40503u + 12345u;
    uint32_t cc = c;
    if constexpr (OP == 7) cc = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 128;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) step<OP>(a, b, cc, h, sm);
    const long long t1 = clock64();
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) x ^= a[i];
    out[threadIdx.x] = x;
    if ((threadIdx.x & 31) == 0) cyc[threadIdx.x >> 5] = t1 - t0;
}

// TMEM read rate: W warps per lane quarter read [32 lanes x 8 columns] ITER times
__global__ void bench_ldtm(uint32_t* out, long long* cyc, int x16) {
    __shared__ uint32_t tm;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tm)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tm + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 32);
    uint32_t acc = 0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
        if (x16) {
            uint32_t r[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                           "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                         : "r"(base + (uint32_t)((it & 1) * 16)) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 16; ++i) acc ^= r[i];
        } else {
            uint32_t r[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                         : "r"(base + (uint32_t)((it & 3) * 8)) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 8; ++i) acc ^= r[i];
        }
    }
    const long long t1 = clock64();
    out[threadIdx.x] = acc;
    if ((threadIdx.x & 31) == 0) cyc[warp] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

static const char* NAMES[] = {"IMAD.HI.U32 (umulhi)", "IMAD.HI + 64-bit addend", "IMAD", "SHF.R.S32", "PRMT", "I2IP (cvt.pack.sat)", "IDP.4A",
                              "LDS (dependent, conflict-free)", "LOP3", "IADD3", "vmaxs2", "umulhi + IMAD pair", "IMAD.HI + SHF pair",
                              "mul.wide >> 16", "umul24 + add", "I2F + FMUL + F2I", "IDP.2A", "IMAD.WIDE + SHF.R.U64 pair", "FFMA", "IADD + clamp"};

template <int OP>
void run(uint32_t* out, long long* cyc, int per_instr) {
    for (int W : {1, 2, 4}) {
        bench<OP><<<1, 128 * W>>>(out, cyc, 0x9e3779b1u, 7u, 0x80000000ull);
        bench<OP><<<1, 128 * W>>>(out, cyc, 0x9e3779b1u, 7u, 0x80000000ull);
        cudaDeviceSynchronize();
        long long h[32];
        cudaMemcpy(h, cyc, sizeof(long long) * 4 * W, cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < 4 * W; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("%-34s W=%d  %.2f cycles per warp-instruction per scheduler (%d SASS instr per step assumed)\n", NAMES[OP], W,
               (double)mx / ((double)ITER * 8 * W * per_instr), per_instr);
    }
}

int main() {
    uint32_t* out;
    long long* cyc;
    cudaMalloc(&out, 4096 * 4);
    cudaMalloc(&cyc, 64 * 8);
    run<0>(out, cyc, 1); run<1>(out, cyc, 1); run<2>(out, cyc, 1); run<3>(out, cyc, 1); run<4>(out, cyc, 1); run<5>(out, cyc, 1);
    run<6>(out, cyc, 1); run<7>(out, cyc, 1); run<8>(out, cyc, 1); run<9>(out, cyc, 1); run<10>(out, cyc, 1); run<11>(out, cyc, 2);
    run<12>(out, cyc, 2); run<13>(out, cyc, 1); run<14>(out, cyc, 1); run<15>(out, cyc, 3);
    run<16>(out, cyc, 1); run<17>(out, cyc, 2); run<18>(out, cyc, 1); run<19>(out, cyc, 2);
    for (int x16 = 0; x16 < 2; ++x16)
        for (int W : {1, 2, 4}) {
            bench_ldtm<<<1, 128 * W>>>(out, cyc, x16);
            cudaDeviceSynchronize();
            long long h[32];
            cudaMemcpy(h, cyc, sizeof(long long) * 4 * W, cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int i = 0; i < 4 * W; ++i) mx = h[i] > mx ? h[i] : mx;
            const double bytes = (double)ITER * 4 * W * 32 * (x16 ? 16 : 8) * 4;
            printf("LDTM.x%d + wait, W=%d warps per lane quarter: %.1f cycles per load per warp, %.1f B/clk per SM\n", x16 ? 16 : 8, W,
                   (double)mx / ITER, bytes / (double)mx);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
