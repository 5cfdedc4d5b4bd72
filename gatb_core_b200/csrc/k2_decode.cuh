// k2_decode.cuh -- rebuilding the canonical k-mers of a super-k-mer record, four at a time (k <= 31, 16-byte records).
//
// Replaces the per-k-mer rolling of the reference's partition readers
//   ReadSuperKCommand::execute / PartitionsByHashCommand decode   kmer/impl/PartitionsCommand.cpp:944-1128, 420-501
// (paths relative to /root/reference/gatb-core/src/gatb/).
//
// A record holds the nucleotides of the super-k-mer in stream order (nucleotide n in bits [2n, 2n+2), kernels.h).
// Chunk c = k-mers 4c .. 4c+3.  Their union is the window X = record bits [8c, 8c+2k+6) (at most 68 bits):
//   stream bits of k-mer i     x_i  = (X >> 2i) & mask(2k)              reverse-complement VALUE = x_i ^ 0b1010..
//   forward VALUE              f_i  = pair_reverse(x_i) aligned right   = (Z >> (6-2i)) & mask(2k)
// where Z = pair_reverse of the whole window, computed once per chunk (3 BREV).  All shifts inside the chunk are
// compile-time constants; canonical = min(forward, reverse complement) as kmer/impl/Model.hpp:857-884.
// __host__ __device__ so that tests/cpp/test_k2_decode.cpp checks it on the CPU nucleotide by nucleotide.
#pragma once
#include "k1_scan.cuh"      // k1s_fshr / k1s_pair_reverse / K1S_HD

struct K2Chunk
{
    uint32_t x0, x1, x2;        // window X (stream order)
    uint32_t z0, z1, z2;        // pair-reversed window, right aligned
    uint32_t lmask, hmask;      // masks of the low / high word of a 2k-bit value
};

// r0..r3 = the record's 32-bit words (r3 already stripped of the length / fine-bin fields); c = chunk index
K1S_HD void k2_chunk_begin (K2Chunk& C, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3, int c, int k)
{
    const int s = 8 * c;                             // 0..48
    const bool hiw = s >= 32;
    const uint32_t a0 = hiw ? r1 : r0, a1 = hiw ? r2 : r1, a2 = hiw ? r3 : r2, a3 = hiw ? 0u : r3;
    const int sh = s & 31;
    C.x0 = k1s_fshr (a0, a1, sh); C.x1 = k1s_fshr (a1, a2, sh); C.x2 = k1s_fshr (a2, a3, sh);
    C.lmask = (k >= 16) ? 0xFFFFFFFFu : ((1u << (2 * k)) - 1);
    C.hmask = (k > 16) ? ((1u << (2 * k - 32)) - 1) : 0u;
    // Y = pair reversal of the 96-bit window (y2 most significant) ; Z = Y >> (96 - (2k+6))
    const uint32_t y2 = k1s_pair_reverse (C.x0), y1 = k1s_pair_reverse (C.x1), y0 = k1s_pair_reverse (C.x2);
    const int t = 90 - 2 * k;                        // 28..86, even
    if (t < 32)      { C.z0 = k1s_fshr (y0, y1, t); C.z1 = k1s_fshr (y1, y2, t); C.z2 = y2 >> t; }
    else if (t < 64) { C.z0 = k1s_fshr (y1, y2, t - 32); C.z1 = y2 >> (t - 32); C.z2 = 0; }
    else             { C.z0 = y2 >> (t - 64); C.z1 = 0; C.z2 = 0; }
}

// canonical value of k-mer I (0..3) of the chunk
template<int I> K1S_HD void k2_chunk_kmer (const K2Chunk& C, uint32_t& lo, uint32_t& hi)
{
    const uint32_t xl = I ? k1s_fshr (C.x0, C.x1, 2 * I) : C.x0;
    const uint32_t xh = I ? k1s_fshr (C.x1, C.x2, 2 * I) : C.x1;
    const uint32_t rl = (xl & C.lmask) ^ (0xAAAAAAAAu & C.lmask);
    const uint32_t rh = (xh & C.hmask) ^ (0xAAAAAAAAu & C.hmask);
    const uint32_t fl = ((I < 3) ? k1s_fshr (C.z0, C.z1, 6 - 2 * I) : C.z0) & C.lmask;
    const uint32_t fh = ((I < 3) ? k1s_fshr (C.z1, C.z2, 6 - 2 * I) : C.z1) & C.hmask;
    const bool f_lt = (fh < rh) || (fh == rh && fl < rl);
    lo = f_lt ? fl : rl; hi = f_lt ? fh : rh;
}

// ---- oriented records (k1_scan.cuh, ORI): the table key of a k-mer is the plain slice of the record -------------------
// Every k-mer of an oriented record is already in its representative orientation, so its 2k stream bits identify it:
// no reversal, no min().  GATB's canonical VALUE is rebuilt only for the distinct k-mers that are emitted (k2_raw_to_canonical).
K1S_HD void k2_chunk_begin_raw (K2Chunk& C, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3, int c, int k)
{
    const int s = 8 * c;                             // 0..48
    const bool hiw = s >= 32;
    const uint32_t a0 = hiw ? r1 : r0, a1 = hiw ? r2 : r1, a2 = hiw ? r3 : r2, a3 = hiw ? 0u : r3;
    const int sh = s & 31;
    C.x0 = k1s_fshr (a0, a1, sh); C.x1 = k1s_fshr (a1, a2, sh); C.x2 = k1s_fshr (a2, a3, sh);
    C.lmask = (k >= 16) ? 0xFFFFFFFFu : ((1u << (2 * k)) - 1);
    C.hmask = (k > 16) ? ((1u << (2 * k - 32)) - 1) : 0u;
    C.z0 = C.z1 = C.z2 = 0;
}
template<int I> K1S_HD void k2_chunk_kmer_raw (const K2Chunk& C, uint32_t& lo, uint32_t& hi)
{
    lo = (I ? k1s_fshr (C.x0, C.x1, 2 * I) : C.x0) & C.lmask;
    hi = (I ? k1s_fshr (C.x1, C.x2, 2 * I) : C.x1) & C.hmask;
}
// stream bits of a k-mer (k <= 31) -> canonical value min(forward, reverse complement), kmer/impl/Model.hpp:857-884
K1S_HD uint64_t k2_raw_to_canonical (uint64_t x, int k)
{
    const uint64_t mask = (k >= 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
    const uint64_t rc = x ^ (0xAAAAAAAAAAAAAAAAULL & mask);
    const uint64_t fwd = (((uint64_t)k1s_pair_reverse ((uint32_t)x) << 32) | k1s_pair_reverse ((uint32_t)(x >> 32))) >> (64 - 2 * k);
    return fwd < rc ? fwd : rc;
}

// ---- 32 <= k <= 63 (32-byte records, 128-bit values as four 32-bit words, word 0 least significant) ----------------------
// Same scheme: chunk c = k-mers 4c .. 4c+3, window X = record bits [8c, 8c+2k+6) (at most 132 bits, five words), one pair
// reversal of the window per chunk, every shift inside the chunk an immediate.
struct K2Chunk2
{
    uint32_t x[5];              // window X (stream order)
    uint32_t z[5];              // pair-reversed window, right aligned
    uint32_t m[4];              // masks of the four words of a 2k-bit value
};

// rw = the record's eight 32-bit words (rw[7] already stripped of the length / fine-bin fields); c = chunk index (0..14)
K1S_HD void k2_chunk2_begin (K2Chunk2& C, const uint32_t* rw, int c, int k)
{
    const int s = 8 * c, wi = s >> 5, sh = s & 31;          // wi in 0..3
    uint32_t a[6];
    #pragma unroll
    for (int i = 0; i < 6; i++) a[i] = (wi + i < 8) ? rw[wi + i] : 0u;
    #pragma unroll
    for (int i = 0; i < 5; i++) C.x[i] = k1s_fshr (a[i], a[i + 1], sh);
    #pragma unroll
    for (int w = 0; w < 4; w++)
    {
        const int bits = 2 * k - 32 * w;                    // bits of the value that fall into word w
        C.m[w] = bits >= 32 ? 0xFFFFFFFFu : (bits <= 0 ? 0u : ((1u << bits) - 1));
    }
    // Y = pair reversal of the 160-bit window (y[4] most significant = reversal of x[0]); Z = Y >> (160 - (2k+6))
    uint32_t y[8];
    #pragma unroll
    for (int i = 0; i < 5; i++) y[4 - i] = k1s_pair_reverse (C.x[i]);
    y[5] = y[6] = y[7] = 0;
    const int t = 154 - 2 * k;                              // 28..90, even
    const int wo = t >> 5, ts = t & 31;                     // wo in 0..2
    #pragma unroll
    for (int i = 0; i < 5; i++)
    {
        const uint32_t lo = wo == 0 ? y[i] : wo == 1 ? y[i + 1] : y[i + 2];
        const uint32_t hi = wo == 0 ? y[i + 1] : wo == 1 ? y[i + 2] : y[i + 3];
        C.z[i] = k1s_fshr (lo, hi, ts);
    }
}

// canonical value of k-mer I (0..3) of the chunk: v[0] least significant word
template<int I> K1S_HD void k2_chunk2_kmer (const K2Chunk2& C, uint32_t v[4])
{
    uint32_t f[4], r[4];
    #pragma unroll
    for (int w = 0; w < 4; w++)
    {
        const uint32_t xs = I ? k1s_fshr (C.x[w], C.x[w + 1], 2 * I) : C.x[w];
        r[w] = (xs & C.m[w]) ^ (0xAAAAAAAAu & C.m[w]);
        f[w] = ((I < 3) ? k1s_fshr (C.z[w], C.z[w + 1], 6 - 2 * I) : C.z[w]) & C.m[w];
    }
    bool f_lt = false, decided = false;
    #pragma unroll
    for (int w = 3; w >= 0; w--)
        if (!decided && f[w] != r[w]) { f_lt = f[w] < r[w]; decided = true; }
    #pragma unroll
    for (int w = 0; w < 4; w++) v[w] = f_lt ? f[w] : r[w];
}
