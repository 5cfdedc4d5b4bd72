// k2_common.cuh -- device helpers shared by the counting kernels (k2_count.cu, k2_fused.cu).
#pragma once
#include "common.cuh"

#define EMPTY64 0xFFFFFFFFFFFFFFFFULL
#define K2_HB      256        // histogram bins kept in shared memory (larger abundances go to global atomics)
#define K2_MAXPROBE 512

// ------------------------------------------------------------------------------------------------ mbarrier / TMA
__device__ __forceinline__ uint32_t smem_u32 (const void* p) { return (uint32_t)__cvta_generic_to_shared (p); }
__device__ __forceinline__ void mbar_init (uint64_t* bar, uint32_t count)
{ asm volatile ("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32 (bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx (uint64_t* bar, uint32_t bytes)
{ asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32 (bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait (uint64_t* bar, uint32_t parity)
{
    asm volatile (
        "{\n .reg .pred p;\n WAIT_%=:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" :: "r"(smem_u32 (bar)), "r"(parity) : "memory");
}
// TMA bulk copy global -> shared (1D), completion on the mbarrier; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void tma_bulk_g2s (void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile ("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                  :: "r"(smem_u32 (dst)), "l"(src), "r"(bytes), "r"(smem_u32 (bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async () { asm volatile ("fence.proxy.async.shared::cta;" ::: "memory"); }


// slot hash of a k <= 31 key for a shared-memory table of 2^(32-shift) slots
__device__ __forceinline__ uint32_t k2_slot32 (uint32_t lo, uint32_t hi, int shift)
{ return (lo * 0x9E3779B1u + hi * 0x85EBCA77u) >> shift; }

// COUNT CONVENTION of the k <= 31 shared-memory tables (k2b_warp_bins, k2b_count_w1, k2f_count_coarse): claiming a slot
// counts ONE occurrence by itself, s_cnt holds the occurrences beyond it (count = s_cnt + 1).  The kernels are bound by the
// throughput of shared-memory atomics (ATOMS: ~1.2 cycles per lane for an add, ~2.8 for a 64-bit CAS, measured), and most
// distinct k-mers of a sequencing run are seen once (one per error and position): they now cost one CAS and no add.
// general open-addressing insert starting at 'slot'; returns the slot (bit 31 set when this call claimed it), or -1
__device__ __forceinline__ int k2_probe_loop (unsigned long long* s_klo, uint32_t* s_cnt, uint32_t slot, unsigned long long key, uint32_t tmask, uint32_t add = 1u)
{
    #pragma unroll 1
    for (int probe = 0; probe < K2_MAXPROBE; probe++)
    {
        unsigned long long cur = s_klo[slot];
        if (cur == EMPTY64) cur = atomicCAS (&s_klo[slot], EMPTY64, key);
        if (cur == EMPTY64) { if (add > 1u) atomicAdd (&s_cnt[slot], add - 1u); return (int)(slot | 0x80000000u); }
        if (cur == key)     { atomicAdd (&s_cnt[slot], add); return (int)slot; }
        slot = (slot + 1) & tmask;
    }
    return -1;
}

