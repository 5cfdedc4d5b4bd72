// k_repart.cu -- the sampling pass of GATB's RepartitorAlgorithm on the device (row f3 of SURVEY.md 8).
//
// Replaces (paths relative to /root/reference/gatb-core/src/gatb/): RepartitorAlgorithm<span>::computeRepartition
// kmer/impl/RepartitionAlgorithm.cpp:394-492 -- the serial pass over the first reads of the bank that charges kx-mers to minimizers
// (SampleRepart, :157-243) and stops after nb_seqs_to_see super-k-mers.  Thread <-> read (k_repart_core.cuh); pass 1 leaves the number
// of super-k-mers of every read, the host finds the read at which the reference's iteration is cancelled, pass 2 accumulates the
// kx-mers per minimizer over the reads up to that one.  The distribution itself (largest bin into the emptiest partition,
// kmer/impl/PartiInfo.cpp:48-106) stays on the host: api.cu.
#include "common.cuh"
#include "kernels.h"
#include "k_repart_core.cuh"

__global__ void __launch_bounds__(128) k_repart_sample (const uint64_t* __restrict__ words, const uint64_t* __restrict__ offsets, int read_len,
                                                        const uint32_t* __restrict__ nmask, uint64_t n_reads, int k, int m,
                                                        uint32_t* __restrict__ sk_count, unsigned long long* __restrict__ kx_table)
{
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const uint64_t roff = offsets ? offsets[r] : r * (uint64_t)read_len;
    const int len = offsets ? (int)(offsets[r + 1] - roff) : read_len;
    auto nuc = [&] (int i) -> uint32_t { const uint64_t p = roff + i; return (uint32_t)(words[p >> 5] >> (2 * (p & 31))) & 3u; };
    auto bad = [&] (int i) -> bool { if (!nmask) return false; const uint64_t p = roff + i; return (nmask[p >> 5] >> (p & 31)) & 1u; };
    uint32_t n;
    if (kx_table) n = krp_scan_read (nuc, bad, len, k, m, [&] (uint32_t mini, uint32_t kx) { atomicAdd (&kx_table[mini], (unsigned long long)kx); });
    else          n = krp_scan_read (nuc, bad, len, k, m, [&] (uint32_t, uint32_t) {});
    if (sk_count) sk_count[r] = n;
}

cudaError_t launch_repart_sample (const LaunchCtx& L, const uint64_t* words, const uint64_t* offsets, int read_len, const uint32_t* nmask,
                                  uint64_t n_reads, int k, int m, uint32_t* sk_count, unsigned long long* kx_table)
{
    if (n_reads == 0) return cudaSuccess;
    k_repart_sample<<<(unsigned)((n_reads + 127) / 128), 128, 0, L.stream>>> (words, offsets, read_len, nmask, n_reads, k, m, sk_count, kx_table);
    (*L.launches)++;
    return cudaGetLastError ();
}
