// common.cuh -- device/host helpers shared by the sm_100a kernels of the k-mer counting path.
//
// Encoding facts (see include/gatb_gpu.h): A=0 C=1 T=2 G=3; the packed read stream holds nucleotide i in bits
// [2i, 2i+2) (little-endian nucleotide order), a GATB k-mer VALUE holds its first nucleotide in the most significant
// position (kmer/impl/Model.hpp:636-657).  Hence for the 2k stream bits x of a k-mer:
//     revcomp value  = x ^ 0b1010..10            (complement is XOR 2: comp_NT = {2,3,0,1}, kmer/impl/ModelData.cpp:41)
//     forward value  = pair_reverse(x) >> (64-2k)
// so no per-nucleotide loop is ever needed to (re)build a k-mer from a record.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define FULL_MASK 0xffffffffu

struct u128 { uint64_t lo, hi; };

__host__ __device__ __forceinline__ bool lt128 (const u128& a, const u128& b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }
__host__ __device__ __forceinline__ bool eq128 (const u128& a, const u128& b) { return a.lo == b.lo && a.hi == b.hi; }

// reverse the order of the 32 two-bit groups of a 64-bit word
__device__ __forceinline__ uint64_t pair_reverse64 (uint64_t x)
{
    uint64_t r = __brevll (x);
    return ((r >> 1) & 0x5555555555555555ULL) | ((r & 0x5555555555555555ULL) << 1);
}
__device__ __forceinline__ uint32_t pair_reverse32 (uint32_t x)
{
    uint32_t r = __brev (x);
    return ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
}

__host__ __device__ __forceinline__ uint64_t mask2k64 (int k) { return k >= 32 ? ~0ULL : ((1ULL << (2*k)) - 1); }

// ---- W = 1 (k <= 31): stream bits (already masked to 2k bits) -> canonical value -------------------------------
__device__ __forceinline__ uint64_t canonical_from_stream64 (uint64_t x, int k)
{
    uint64_t rc  = x ^ (0xAAAAAAAAAAAAAAAAULL & mask2k64 (k));
    uint64_t fwd = pair_reverse64 (x) >> (64 - 2*k);
    return fwd < rc ? fwd : rc;
}
// ---- W = 2 (32 <= k <= 63): 2k stream bits in (lo,hi) ----------------------------------------------------------
__device__ __forceinline__ u128 canonical_from_stream128 (u128 x, int k)
{
    // rc = x ^ AA.. masked to 2k bits
    u128 rc; rc.lo = x.lo ^ 0xAAAAAAAAAAAAAAAAULL; rc.hi = x.hi ^ (0xAAAAAAAAAAAAAAAAULL & mask2k64 (k - 32));
    // fwd = pair_reverse over 128 bits, then >> (128-2k)
    uint64_t rl = pair_reverse64 (x.hi), rh = pair_reverse64 (x.lo);       // reversed 128-bit value = (rh:rl)
    int s = 128 - 2*k;                                                       // 2..64
    u128 fwd;
    if (s == 64) { fwd.lo = rh; fwd.hi = 0; }
    else         { fwd.lo = (rl >> s) | (rh << (64 - s)); fwd.hi = rh >> s; }
    return lt128 (fwd, rc) ? fwd : rc;
}

// ---- GATB hash functions (bit-exact restatements; used by the Bloom kernels) ------------------------------------
// tools/math/LargeInt1.pri:157-170
__host__ __device__ __forceinline__ uint64_t gatb_hash64 (uint64_t key, uint64_t seed)
{
    uint64_t hash = seed;
    hash ^= (hash <<  7) ^  key * (hash >> 3) ^ (~((hash << 11) + (key ^ (hash >> 5))));
    hash = (~hash) + (hash << 21);
    hash = hash ^ (hash >> 24);
    hash = (hash + (hash << 3)) + (hash << 8);
    hash = hash ^ (hash >> 14);
    hash = (hash + (hash << 2)) + (hash << 4);
    hash = hash ^ (hash >> 28);
    hash = hash + (hash << 31);
    return hash;
}
// tools/math/LargeInt1.pri:137-154 with the right-alignment to sizeKmer nucleotides
__device__ __forceinline__ uint64_t gatb_revcomp64 (uint64_t x, int sizeKmer)
{
    if (sizeKmer <= 0) return 0;
    uint64_t r = pair_reverse64 (x) ^ 0xAAAAAAAAAAAAAAAAULL;
    return r >> (2*(32 - sizeKmer));
}
// tools/math/LargeInt2.pri:168-197
__device__ __forceinline__ u128 gatb_revcomp128 (u128 x, int k)
{
    int nb_high = k > 32 ? k - 32 : 0, nb_low = k > 32 ? 32 : k;
    uint64_t rh = (k <= 32) ? 0 : gatb_revcomp64 (x.hi, nb_high);
    uint64_t rl = gatb_revcomp64 (x.lo, nb_low);
    u128 res;
    if (nb_high == 0)       { res.lo = rl; res.hi = 0; }
    else if (nb_high == 32) { res.lo = rh; res.hi = rl; }
    else                    { res.lo = (rl << (2*nb_high)) + rh; res.hi = rl >> (64 - 2*nb_high); }
    return res;
}

// ---- minimizer key of a canonical m-mer under GATB's rule (kmer/impl/Model.hpp:1040-1064, is_allowed :1220-1251) --
// cm = min(mmer, revcomp_m(mmer)); not allowed (an "AA" anywhere but at the prefix) -> 4^m-1
__host__ __device__ __forceinline__ uint32_t gatb_mmer_key (uint32_t cm, uint32_t mmask, uint32_t mask_ma1)
{
    uint32_t a1 = ~(cm | (cm >> 2));
    a1 = ((a1 >> 1) & a1) & mask_ma1;
    return a1 ? mmask : cm;
}
__host__ __device__ __forceinline__ uint32_t gatb_mask_ma1 (int m)
{ return (uint32_t)(0x5555555555555555ULL & ((1ULL << ((m - 2)*2)) - 1)); }

// lowbias32 mixer (public-domain constants) used to spread device bins
__host__ __device__ __forceinline__ uint32_t mix32 (uint32_t h)
{
    h ^= h >> 16; h *= 0x7feb352dU; h ^= h >> 15; h *= 0x846ca68bU; h ^= h >> 16;
    return h;
}

__device__ __forceinline__ uint64_t ldg64 (const uint64_t* p) { return __ldg ((const unsigned long long*)p); }

// splitmix64 of the synthetic generator (mirrors oracle/kmer_oracle.c orc_splitmix64)
__host__ __device__ __forceinline__ uint64_t splitmix64 (uint64_t x)
{
    uint64_t z = x + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
