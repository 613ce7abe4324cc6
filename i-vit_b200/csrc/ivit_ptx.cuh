// Thin inline-PTX wrappers for the Blackwell (sm_100a) async machinery used by the GEMM and
// attention kernels: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld),
// UMMA shared-memory and instruction descriptors.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define IVIT_PTX __device__ __forceinline__

namespace ivit {
namespace ptx {

IVIT_PTX uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- programmatic dependent launch -----------------------------------------------------
// Wait until every grid this one depends on has completed and its memory is visible.  A no-op when the kernel was
// launched without the programmatic-serialization attribute (ivit::launch_k, IVIT_PDL=1).
IVIT_PTX void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- mbarrier ----------------------------------------------------------------------
IVIT_PTX void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
IVIT_PTX void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
IVIT_PTX void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
IVIT_PTX void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
IVIT_PTX void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
IVIT_PTX bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
IVIT_PTX void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// ---- TMA -----------------------------------------------------------------------------
IVIT_PTX void prefetch_tensormap(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// 2D tile load global -> shared, completion on an mbarrier (c0 = innermost coordinate).
IVIT_PTX void tma_load_2d(uint32_t dst_smem, const void* tmap, uint32_t bar, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst_smem), "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
IVIT_PTX void tma_load_3d(uint32_t dst_smem, const void* tmap, uint32_t bar, int32_t c0, int32_t c1, int32_t c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst_smem), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// Same, multicast: the tile lands at the same CTA-relative shared address in every CTA of `cta_mask` and signals
// the mbarrier at the same CTA-relative address in each of them.
IVIT_PTX void tma_load_2d_mc(uint32_t dst_smem, const void* tmap, uint32_t bar, int32_t c0, int32_t c1, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%4, %5}], [%2], %3;"
        ::"r"(dst_smem), "l"(tmap), "r"(bar), "h"(cta_mask), "r"(c0), "r"(c1)
        : "memory");
}
// ---- thread-block clusters ---------------------------------------------------------------
IVIT_PTX uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
IVIT_PTX void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `cta_rank` of the cluster
IVIT_PTX uint32_t mapa(uint32_t smem_addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta_rank));
    return r;
}
// arrive on an mbarrier that may live in another CTA of the cluster (address from mapa).  No cluster-scope release:
// the only thing ordered through it here are tcgen05.ld reads, which tcgen05.fence::before_thread_sync orders
// (a .release.cluster arrive costs a full MEMBAR, ~7 % of the epilogue's issue slots when measured).
IVIT_PTX void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a local mbarrier that peers arrive on.  Plain (cta-scope) acquire like the arrive above: a cluster-scope
// acquire makes ptxas put an L1 invalidation (CCTL.IVALL) into the spin loop.
IVIT_PTX bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
IVIT_PTX void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait_cluster(bar, parity)) {
    }
}
// CTA-pair load (cta_group::2): the tile lands in MY shared memory, the transaction bytes are signalled on an
// mbarrier that may live in the peer CTA (`cluster_bar` is a shared::cluster address, see mapa).
IVIT_PTX void tma_load_2d_pair(uint32_t dst_smem, const void* tmap, uint32_t cluster_bar, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst_smem), "l"(tmap), "r"(cluster_bar), "r"(c0), "r"(c1)
        : "memory");
}
// 2D tile store shared -> global (bulk async group).
IVIT_PTX void tma_store_2d(const void* tmap, uint32_t src_smem, int32_t c0, int32_t c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(tmap), "r"(src_smem), "r"(c0), "r"(c1)
                 : "memory");
}
IVIT_PTX void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
IVIT_PTX void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
IVIT_PTX void tma_store_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// ---- tcgen05 ---------------------------------------------------------------------------
IVIT_PTX void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
IVIT_PTX void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
IVIT_PTX void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
IVIT_PTX void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
IVIT_PTX void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem], int8 operands, int32 accumulate.  Single-thread issue.
IVIT_PTX void mma_i8_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed.
IVIT_PTX void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// Same, arriving on the mbarrier at the same CTA-relative address in every CTA of `cta_mask`.
IVIT_PTX void mma_commit_mc(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(cta_mask) : "memory");
}
// ---- CTA pair (cta_group::2): the two CTAs of a cluster, on the two SMs of one TPC, execute one 256-row MMA.  Each
// holds its own 128 rows of A and HALF of the W tile (the tensor cores read both halves), and each receives the
// 128 accumulator rows of its A rows in its own TMEM.  Issued by one thread of the even-ranked (leader) CTA.
IVIT_PTX void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
IVIT_PTX void tmem_relinquish_pair() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
IVIT_PTX void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
IVIT_PTX void mma_i8_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (when all previously issued pair MMAs have completed) on the mbarrier at this CTA-relative address in
// every CTA of `cta_mask`
IVIT_PTX void mma_commit_pair_mc(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(cta_mask) : "memory");
}
IVIT_PTX void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp gets lane (base_lane + i).
IVIT_PTX void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
IVIT_PTX void tmem_ld_32x32b_x1(uint32_t taddr, uint32_t& r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
}
IVIT_PTX void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr)
        : "memory");
}
IVIT_PTX void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

// ---- UMMA descriptors ------------------------------------------------------------------
// Shared-memory matrix descriptor for a K-major operand tile stored as rows of 128 bytes with
// the 128-byte swizzle (exactly what a TMA box {128 B, rows} with CU_TENSOR_MAP_SWIZZLE_128B
// writes): 8-row x 128 B swizzle atoms, stride between atoms (SBO) = 1024 B, LBO unused (=1),
// descriptor version 1 (sm_100), layout type 2 (SWIZZLE_128B).
// Field layout: cute/arch/mma_sm100_desc.hpp (SmemDescriptor) in the vendored CUTLASS headers.
IVIT_PTX uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);     // start address  [0,14)
    d |= (uint64_t)1 << 16;                           // leading byte offset (ignored for SW128 K-major)
    d |= (uint64_t)(1024u >> 4) << 32;                // stride byte offset [32,46)
    d |= (uint64_t)1 << 46;                           // version = 1 (Blackwell)
    d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
    return d;
}
// Instruction descriptor, kind::i8: D = S32, A/B = signed (1) or unsigned (0) 8-bit, both K-major.
// Field layout: cute/arch/mma_sm100_desc.hpp (InstrDescriptor).
__host__ __device__ constexpr uint32_t umma_idesc_i8(int M, int N, int a_signed, int b_signed) {
    return (2u << 4) | ((uint32_t)a_signed << 7) | ((uint32_t)b_signed << 10) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

}  // namespace ptx
}  // namespace ivit
