// Thin inline-PTX wrappers for the Blackwell (sm_100a) asynchronous machinery used by the int8 tensor-core engine (ozaki.cuh):
// mbarrier, TMA tile loads (cp.async.bulk.tensor), tcgen05 MMA / commit / TMEM alloc / TMEM load, and the shared-memory
// matrix descriptor.  Field layouts follow the PTX ISA "tcgen05" chapter (shared-memory descriptor, instruction descriptor).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace zigp {
namespace tc05 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- 256-bit global stores (sm_100: STG.256) -----------------------------------------------------------------------------
__device__ __forceinline__ void st_global_f64x4(double* p, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d));   // no memory clobber: callers never read back what they store, and loads must be free to move above
}
__device__ __forceinline__ void st_global_b32x8(void* p, uint4 a, uint4 b) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z),
                 "r"(b.w)
                 : "memory");
}

// ---- mbarrier ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU (the box is shared); after ~2^28 polls the kernel traps instead.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); spin++) {
        if (spin > (1u << 28)) __trap();
    }
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// ---- thread-block clusters ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t num_clusters_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%nclusterid.x;\n" : "=r"(r));
    return r;
}
// all threads of all CTAs of the cluster
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// ---- TMA: 3-D tiled load global -> shared, completion on an mbarrier ------------------------------------------------------
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// the same, delivered to the same CTA-relative shared address (and mbarrier) of every CTA in cta_mask
__device__ __forceinline__ void tma_load_3d_mcast(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;\n" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(cta_mask)
        : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(map) : "memory");
}

// ---- tcgen05 -------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fence_before_thread_sync() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_after_thread_sync() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

// whole warp; writes the TMEM base address of the allocation to *smem_slot
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_slot)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "n"(COLS) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, int8 x int8 -> int32, issued by ONE thread
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when they complete (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// the same, arriving on the barrier at this CTA-relative address in every CTA of cta_mask
__device__ __forceinline__ void mma_commit_mcast(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "h"(cta_mask)
                 : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// Shared-memory matrix descriptor, K-major operand, rows of SW bytes (SW = 32 / 64 / 128 = the TMA swizzle mode of the tile),
// 8-row groups packed densely (stride-byte-offset = 8 * SW).  Bits: [0,14) start >> 4, [16,30) LBO >> 4 (unused for swizzled
// K-major, 1), [32,46) SBO >> 4, [46,48) version = 1 (sm_100), [61,64) layout type (6 = 32B, 4 = 64B, 2 = 128B swizzle).
template <int SW>
__device__ __forceinline__ uint64_t smem_desc_kmajor(uint32_t saddr) {
    constexpr uint64_t layout = (SW == 32) ? 6 : (SW == 64) ? 4 : 2;
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((8 * SW) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= layout << 61;
    return d;
}

// Shared-memory matrix descriptor, MN-major operand (the M / N index is the contiguous one; k runs over rows of SW bytes): the tile is
// [k rows][SW bytes of consecutive M / N elements], 8-row groups packed densely, exactly what a TMA box {SW, rows, ..} with the SW-byte swizzle
// writes.  Canonical form (CUTLASS make_umma_desc<Major::MN>, units of 16 bytes): ((SW/16, n), (8, k)) : ((1, LBO), (SW/16, SBO)) --
// LBO = byte distance between consecutive SW-byte atoms along M / N (for operands wider than SW bytes, e.g. concatenated planes),
// SBO = byte distance between consecutive groups of 8 k-rows (= 8 * SW when the rows are dense).
template <int SW>
__device__ __forceinline__ uint64_t smem_desc_mnmajor(uint32_t saddr, uint32_t lbo_bytes) {
    constexpr uint64_t layout = (SW == 32) ? 6 : (SW == 64) ? 4 : 2;
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((8 * SW) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= layout << 61;
    return d;
}

// Instruction descriptor, kind::i8, signed 8-bit A and B (both K-major), int32 accumulator, M x N tile.
// Bits: [4,6) D format (2 = S32), [7,10) A format (1 = S8), [10,13) B format (1 = S8), 15 / 16 A / B major (0 = K),
// [17,23) N >> 3, [24,29) M >> 4.
__host__ __device__ __forceinline__ uint32_t idesc_i8(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// the same with both operands MN-major (bits 15 / 16)
__host__ __device__ __forceinline__ uint32_t idesc_i8_mn(int M, int N) { return idesc_i8(M, N) | (1u << 15) | (1u << 16); }

}  // namespace tc05
}  // namespace zigp
