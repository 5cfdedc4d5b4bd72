// k_misc.cu -- small kernels around the hot path:
//   * serialisation of GATB-mode records into the reference's super-k-mer byte format
//         Kmer<span>::SuperKmer::save                    kmer/impl/Model.hpp:1386-1471
//         CacheSuperKmerBinFiles::insertSuperkmer        tools/storage/impl/Storage.cpp:567-580  ([u8 nbK][bytes])
//   * Bloom insertion (all three kinds)                  tools/collections/impl/Bloom.hpp:394-412, 445-459, 555-588
//   * synthetic read generator and ASCII -> 2-bit packer (Data::ConvertASCII, tools/misc/api/Data.hpp:185)
#include "common.cuh"
#include "kernels.h"
#include "gatb_tables.h"

// ------------------------------------------------------------------------------------------------ serialisation
template<int W>
__device__ __forceinline__ int rec_load (const void* bins, uint64_t idx, uint64_t* r)
{
    const uint4* p = (const uint4*)bins + idx * W;
    uint4 a = p[0];
    r[0] = (uint64_t)a.x | ((uint64_t)a.y << 32); r[1] = (uint64_t)a.z | ((uint64_t)a.w << 32);
    if (W == 2) { uint4 b = p[1]; r[2] = (uint64_t)b.x | ((uint64_t)b.y << 32); r[3] = (uint64_t)b.z | ((uint64_t)b.w << 32); }
    int len;
    if (W == 1) { len = (int)((r[1] >> REC_LEN_SHIFT_W1) & 31); r[1] &= (1ULL << REC_LEN_SHIFT_W1) - 1; }
    else        { len = (int)((r[3] >> REC_LEN_SHIFT_W2) & 63); r[3] &= (1ULL << REC_LEN_SHIFT_W2) - 1; }
    return len;
}
__device__ __forceinline__ uint32_t rec_nt (const uint64_t* r, int i) { return (uint32_t)(r[i >> 5] >> (2*(i & 31))) & 3u; }

template<int W>
__global__ void __launch_bounds__(256) k_serialize_sizes (int k, const void* bins, const uint32_t* cursors, uint32_t cap, unsigned long long* key_bytes)
{
    const uint32_t key = blockIdx.x;
    const uint32_t n = min (cursors[key], cap);
    unsigned long long sum = 0;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
    {
        uint64_t r[4]; int len = rec_load<W> (bins, (uint64_t)key * cap + i, r);
        sum += 1 + (k + len - 1 + 3) / 4;
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync (FULL_MASK, sum, o);
    if ((threadIdx.x & 31) == 0 && sum) atomicAdd (&key_bytes[key], sum);
}

template<int W>
__global__ void __launch_bounds__(256) k_serialize_write (int k, const void* bins, const uint32_t* cursors, uint32_t cap,
                                                          const uint64_t* key_off, unsigned long long* key_cur, uint8_t* out)
{
    const uint32_t key = blockIdx.x;
    const uint32_t n = min (cursors[key], cap);
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
    {
        uint64_t r[4] = {0,0,0,0}; int len = rec_load<W> (bins, (uint64_t)key * cap + i, r);
        const int nbytes = 1 + (k + len - 1 + 3) / 4;
        uint8_t* dst = out + key_off[key] + atomicAdd (&key_cur[key], (unsigned long long)nbytes);
        *dst++ = (uint8_t)len;
        // first k-mer, forward VALUE = first nucleotide most significant; bytes leave from the low end:
        // byte b holds nucleotides k-4b-4 .. k-4b-1 with the LAST of them in the low 2 bits (Model.hpp:1418-1431)
        int rem = k, pos = k;                                   // pos = one past the last nucleotide not yet written
        while (rem >= 4)
        {
            uint32_t byte = 0;
            #pragma unroll
            for (int j = 0; j < 4; j++) byte |= rec_nt (r, pos - 1 - j) << (2*j);
            *dst++ = (uint8_t)byte; rem -= 4; pos -= 4;
        }
        uint32_t newbyte = 0;
        for (int j = 0; j < rem; j++) newbyte |= rec_nt (r, pos - 1 - j) << (2*j);
        int uid = rem, skid = 1;
        for (;;)
        {
            while (uid < 4 && skid < len) { newbyte |= rec_nt (r, k + skid - 1) << (2*uid); uid++; skid++; }   // last nt of k-mer 'skid'
            if (uid > 0) *dst++ = (uint8_t)newbyte;
            if (skid >= len) break;
            newbyte = 0; uid = 0;
        }
    }
}

cudaError_t launch_serialize_sizes (const LaunchCtx& L, int W, int k, const void* bins, const uint32_t* cursors, uint32_t nb1,
                                    uint32_t cap, unsigned long long* key_bytes)
{
    if (W == 1) k_serialize_sizes<1><<<nb1, 256, 0, L.stream>>> (k, bins, cursors, cap, key_bytes);
    else        k_serialize_sizes<2><<<nb1, 256, 0, L.stream>>> (k, bins, cursors, cap, key_bytes);
    (*L.launches)++;
    return cudaGetLastError ();
}
cudaError_t launch_serialize_write (const LaunchCtx& L, int W, int k, const void* bins, const uint32_t* cursors, uint32_t nb1,
                                    uint32_t cap, const uint64_t* key_off, unsigned long long* key_cur, uint8_t* out)
{
    if (W == 1) k_serialize_write<1><<<nb1, 256, 0, L.stream>>> (k, bins, cursors, cap, key_off, key_cur, out);
    else        k_serialize_write<2><<<nb1, 256, 0, L.stream>>> (k, bins, cursors, cap, key_off, key_cur, out);
    (*L.launches)++;
    return cudaGetLastError ();
}

// ------------------------------------------------------------------------------------------------ Bloom
__constant__ uint64_t c_random_values[256] = GATB_RANDOM_VALUES_INIT;      // kmer/impl/ModelData.cpp:302
__constant__ uint64_t c_bloom_seeds[10];                                   // HashFunctors::generate_hash_seed, Bloom.hpp:82-94
__constant__ uint8_t  c_cano2[16] = {0,1,2,3,4,5,3,7,8,9,0,4,9,13,1,5};    // BloomNeighborCoherent ctor, Bloom.hpp:526-541

// LargeInt1.pri:190-211 adds random_values[key & 255]; LargeInt2.pri:248-251 -> NativeInt64.hpp:211-221 does not
template<int W> __device__ __forceinline__ uint64_t simplehash16 (uint64_t key_lo, int shift)
{
    uint64_t input = key_lo >> shift;
    uint64_t res = c_random_values[input & 255];
    input >>= 8;
    res ^= c_random_values[input & 255];
    if (W == 1) res ^= c_random_values[key_lo & 255];
    return res;
}
template<int W> __device__ __forceinline__ uint64_t hash1 (uint64_t lo, uint64_t hi, uint64_t seed)
{ return W == 1 ? gatb_hash64 (lo, seed) : (gatb_hash64 (hi, seed) ^ gatb_hash64 (lo, seed)); }

// byte[h>>3] |= 1 << (h&7)  ==  word32[h>>5] |= 1 << (h&31) on a little-endian array
__device__ __forceinline__ void bloom_set (uint32_t* words, uint64_t h) { atomicOr (&words[h >> 5], 1u << (h & 31)); }

template<int W, int KIND>
__global__ void __launch_bounds__(256) k4_bloom_insert (int k, int nb_hash, uint64_t tai, int pow2, uint64_t reduced,
                                                        const uint64_t* __restrict__ lo, const uint64_t* __restrict__ hi,
                                                        uint64_t n, uint32_t* words)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    {
        uint64_t l = lo[i], h = (W == 2) ? hi[i] : 0;
        if (KIND == 0)
        {   // BloomSynchronized::insert, Bloom.hpp:394-412
            for (int f = 0; f < nb_hash; f++)
            {
                uint64_t h1 = hash1<W> (l, h, c_bloom_seeds[f]);
                h1 = pow2 ? (h1 & tai) : (h1 % tai);
                bloom_set (words, h1);
            }
        }
        else if (KIND == 1)
        {   // BloomCacheCoherent::insert, Bloom.hpp:445-459
            uint64_t h0 = hash1<W> (l, h, c_bloom_seeds[0]) % reduced;
            bloom_set (words, h0);
            for (int f = 1; f < nb_hash; f++) bloom_set (words, h0 + (simplehash16<W> (l, f) & 4095));
        }
        else
        {   // BloomNeighborCoherent::insert, Bloom.hpp:555-588
            uint32_t suffix = (uint32_t)l & 3u;
            // prefix = (item & (3 << 2(k-1))) >> 2(k-2): the first nucleotide, placed in bits [2,4)
            uint32_t first_nt = (W == 1 || k <= 32) ? (uint32_t)(l >> (2*(k-1))) & 3u : (uint32_t)(h >> (2*(k-1) - 64)) & 3u;
            uint32_t pref = ((first_nt << 2) + suffix) & 15u;
            uint64_t pref_val = c_cano2[pref];
            uint64_t pl, ph = 0;                              // hashpart = (item >> 2) & mask(k-2)
            if (W == 1) { pl = (l >> 2) & mask2k64 (k - 2); }
            else
            {
                pl = (l >> 2) | (h << 62); ph = h >> 2;
                if (k - 2 <= 32) { pl &= mask2k64 (k - 2); ph = 0; } else ph &= mask2k64 (k - 2 - 32);
            }
            if (W == 1)
            {
                uint64_t rv = gatb_revcomp64 (pl, k - 2);
                if (rv < pl) pl = rv;
            }
            else
            {
                u128 x; x.lo = pl; x.hi = ph;
                u128 rv = gatb_revcomp128 (x, k - 2);
                if (lt128 (rv, x)) { pl = rv.lo; ph = rv.hi; }
            }
            uint64_t racine = hash1<W> (pl, ph, c_bloom_seeds[0]) % reduced;
            uint64_t h0 = racine + pref_val;
            bloom_set (words, h0);
            for (int f = 1; f < nb_hash; f++) bloom_set (words, h0 + (simplehash16<W> (pl, f) & 4095));
        }
    }
}

static bool g_seeds_ready[64] = {false};
cudaError_t launch_bloom_insert (const LaunchCtx& L, int kind, int W, int k, int nb_hash, uint64_t tai, int pow2, uint64_t reduced,
                                 const uint64_t* lo, const uint64_t* hi, uint64_t n, uint32_t* words)
{
    int dev = 0; cudaGetDevice (&dev);
    if (dev < 64 && !g_seeds_ready[dev])
    {
        static const uint64_t rbase[10] = {
            0xAAAAAAAA55555555ULL, 0x33333333CCCCCCCCULL, 0x6666666699999999ULL, 0xB5B5B5B54B4B4B4BULL,
            0xAA55AA5555335533ULL, 0x33CC33CCCC66CC66ULL, 0x6699669999B599B5ULL, 0xB54BB54B4BAA4BAAULL,
            0xAA33AA3355CC55CCULL, 0x33663366CC99CC99ULL };
        uint64_t seeds[10];
        for (int i = 0; i < 10; i++) seeds[i] = rbase[i];
        for (int i = 0; i < 10; i++) seeds[i] = seeds[i] * seeds[(i + 3) % 10];      // in place and sequential, as the reference does
        cudaError_t e = cudaMemcpyToSymbol (c_bloom_seeds, seeds, sizeof(seeds));
        if (e != cudaSuccess) return e;
        g_seeds_ready[dev] = true;
    }
    if (n == 0) return cudaSuccess;
    uint64_t blocks = (n + 255) / 256; unsigned grid = (unsigned)(blocks < (uint64_t)L.sm_count * 16 ? blocks : (uint64_t)L.sm_count * 16);
    #define BL(Wv,Kv) k4_bloom_insert<Wv,Kv><<<grid, 256, 0, L.stream>>> (k, nb_hash, tai, pow2, reduced, lo, hi, n, words)
    if (W == 1) { if (kind == 0) BL(1,0); else if (kind == 1) BL(1,1); else BL(1,2); }
    else        { if (kind == 0) BL(2,0); else if (kind == 1) BL(2,1); else BL(2,2); }
    #undef BL
    (*L.launches)++;
    return cudaGetLastError ();
}

// ------------------------------------------------------------------------------------------------ synthetic reads
// One thread builds one packed 32-bit word (16 nucleotides) of the back-to-back read stream; mirrors
// oracle/kmer_oracle.c orc_synth_reads + orc_pack_2bit.
__device__ __forceinline__ uint32_t synth_base (uint64_t seed, uint64_t sR, uint64_t sE, uint64_t genome_len, uint64_t r, int L, int j)
{
    uint64_t h0 = splitmix64 (sR + 2*r), h1 = splitmix64 (sR + 2*r + 1);
    uint64_t start = h0 % (genome_len - L + 1);
    int flip = (int)(h1 >> 63);
    int jj = flip ? (L - 1 - j) : j;                        // position in generation order
    uint32_t b = (uint32_t)(splitmix64 (seed * 0x100000001B3ULL + start + jj) >> 61) & 3u;
    uint64_t e = splitmix64 (sE + r * (uint64_t)L + jj);
    if ((e % 100) == 0) b = (b + 1 + (uint32_t)((e >> 32) % 3)) & 3u;
    return flip ? (b ^ 2u) : b;
}
__global__ void __launch_bounds__(256) k_synth_reads (uint64_t seed, uint64_t sR, uint64_t sE, uint64_t genome_len, uint64_t first_read,
                                                      uint64_t n_reads, int L, uint32_t* words, uint64_t n_words)
{
    for (uint64_t wi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; wi < n_words; wi += (uint64_t)gridDim.x * blockDim.x)
    {
        uint32_t word = 0;
        uint64_t p = wi * 16;
        uint64_t r = p / L; int j = (int)(p - r * L);
        for (int t = 0; t < 16; t++)
        {
            if (r < n_reads) word |= synth_base (seed, sR, sE, genome_len, first_read + r, L, j) << (2*t);
            if (++j == L) { j = 0; r++; }
        }
        words[wi] = word;
    }
}
cudaError_t launch_synth_reads (const LaunchCtx& L, uint64_t seed, uint64_t genome_len, uint64_t first_read, uint64_t n_reads,
                                int len, uint8_t* packed)
{
    uint64_t n_words = (n_reads * (uint64_t)len + 15) / 16;
    if (n_words == 0) return cudaSuccess;
    uint64_t sR = splitmix64 (seed ^ 0x5EEDC0DE00000001ULL), sE = splitmix64 (seed ^ 0x5EEDC0DE00000002ULL);
    uint64_t blocks = (n_words + 255) / 256; unsigned grid = (unsigned)(blocks < (uint64_t)L.sm_count * 32 ? blocks : (uint64_t)L.sm_count * 32);
    k_synth_reads<<<grid, 256, 0, L.stream>>> (seed, sR, sE, genome_len, first_read, n_reads, len, (uint32_t*)packed, n_words);
    (*L.launches)++;
    return cudaGetLastError ();
}

// metagenome-like reads (BASELINE config 5): species by Zipf through the host-built threshold table; mirrors orc_synth_reads_zipf
__device__ __forceinline__ uint32_t synth_base_zipf (uint64_t seed, uint64_t sR, uint64_t sE, uint64_t n_species, const uint64_t* __restrict__ cdf,
                                                     const uint64_t* __restrict__ goff, uint64_t r, int L, int j)
{
    const uint64_t h0 = splitmix64 (sR + 3*r), h1 = splitmix64 (sR + 3*r + 1), h2 = splitmix64 (sR + 3*r + 2);
    uint64_t lo = 0, hi = n_species - 1;
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (h2 <= cdf[mid]) hi = mid; else lo = mid + 1; }
    const uint64_t glen = goff[lo+1] - goff[lo];
    const uint64_t start = goff[lo] + h0 % (glen - L + 1);
    const int flip = (int)(h1 >> 63);
    const int jj = flip ? (L - 1 - j) : j;
    uint32_t b = (uint32_t)(splitmix64 (seed * 0x100000001B3ULL + start + jj) >> 61) & 3u;
    const uint64_t e = splitmix64 (sE + r * (uint64_t)L + jj);
    if ((e % 100) == 0) b = (b + 1 + (uint32_t)((e >> 32) % 3)) & 3u;
    return flip ? (b ^ 2u) : b;
}
__global__ void __launch_bounds__(256) k_synth_reads_zipf (uint64_t seed, uint64_t sR, uint64_t sE, uint64_t n_species, const uint64_t* __restrict__ cdf,
                                                           const uint64_t* __restrict__ goff, uint64_t first_read, uint64_t n_reads, int L,
                                                           uint32_t* words, uint64_t n_words)
{
    for (uint64_t wi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; wi < n_words; wi += (uint64_t)gridDim.x * blockDim.x)
    {
        uint32_t word = 0;
        uint64_t p = wi * 16;
        uint64_t r = p / L; int j = (int)(p - r * L);
        for (int t = 0; t < 16; t++)
        {
            if (r < n_reads) word |= synth_base_zipf (seed, sR, sE, n_species, cdf, goff, first_read + r, L, j) << (2*t);
            if (++j == L) { j = 0; r++; }
        }
        words[wi] = word;
    }
}
cudaError_t launch_synth_reads_zipf (const LaunchCtx& L, uint64_t seed, uint64_t n_species, const uint64_t* d_cdf, const uint64_t* d_goff,
                                     uint64_t first_read, uint64_t n_reads, int len, uint8_t* packed)
{
    uint64_t n_words = (n_reads * (uint64_t)len + 15) / 16;
    if (n_words == 0) return cudaSuccess;
    uint64_t sR = splitmix64 (seed ^ 0x5EEDC0DE00000001ULL), sE = splitmix64 (seed ^ 0x5EEDC0DE00000002ULL);
    uint64_t blocks = (n_words + 255) / 256; unsigned grid = (unsigned)(blocks < (uint64_t)L.sm_count * 32 ? blocks : (uint64_t)L.sm_count * 32);
    k_synth_reads_zipf<<<grid, 256, 0, L.stream>>> (seed, sR, sE, n_species, d_cdf, d_goff, first_read, n_reads, len, (uint32_t*)packed, n_words);
    (*L.launches)++;
    return cudaGetLastError ();
}

// ------------------------------------------------------------------------------------------------ ASCII packer
// one thread packs 32 characters -> one 64-bit... kept at 16 characters -> one u32 of nucleotides and half a mask word
__global__ void __launch_bounds__(256) k_pack_ascii (const char* __restrict__ ascii, uint64_t n, uint32_t* words, uint32_t* nmask, unsigned long long* n_invalid)
{
    unsigned long long bad = 0;
    const uint64_t n_groups = (n + 31) / 32;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += (uint64_t)gridDim.x * blockDim.x)
    {
        uint32_t w0 = 0, w1 = 0, mk = 0;
        for (int t = 0; t < 32; t++)
        {
            uint64_t i = g * 32 + t;
            if (i >= n) break;
            unsigned char c = (unsigned char)ascii[i];
            uint32_t code = (c >> 1) & 3u;                                     // Data::ConvertASCII
            bool ok = (c=='A'||c=='C'||c=='G'||c=='T'||c=='a'||c=='c'||c=='g'||c=='t');
            if (!ok) { mk |= 1u << t; bad++; }
            if (t < 16) w0 |= code << (2*t); else w1 |= code << (2*(t-16));
        }
        words[2*g] = w0; words[2*g+1] = w1;
        if (nmask) nmask[g] = mk;
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync (FULL_MASK, bad, o);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd (n_invalid, bad);
}
cudaError_t launch_pack_ascii (const LaunchCtx& L, const char* ascii, uint64_t n, uint32_t* packed_words, uint32_t* nmask, unsigned long long* n_invalid)
{
    uint64_t groups = (n + 31) / 32;
    if (groups == 0) return cudaSuccess;
    uint64_t blocks = (groups + 255) / 256; unsigned grid = (unsigned)(blocks < (uint64_t)L.sm_count * 32 ? blocks : (uint64_t)L.sm_count * 32);
    k_pack_ascii<<<grid, 256, 0, L.stream>>> (ascii, n, packed_words, nmask, n_invalid);
    (*L.launches)++;
    return cudaGetLastError ();
}

// ------------------------------------------------------------------------------------------------ ASCII packer, appending
// The streaming input path (gatb_gpu_reads_push_ascii): a batch of n characters is packed into the stream at nucleotide
// offset 'base' (any value).  One thread produces the 32 stream positions [32g, 32g+32): two nucleotide words and one
// mask word; the first and the last group of a batch share their words with the neighbouring batches, so they are OR-ed
// in (the buffers are zero beyond the last nucleotide written).  Same encoding as k_pack_ascii (Data::ConvertASCII).
__global__ void __launch_bounds__(256) k_pack_ascii_at (const char* __restrict__ ascii, uint64_t n, uint64_t base, uint32_t* words, uint32_t* nmask,
                                                        unsigned long long* n_invalid)
{
    unsigned long long bad = 0;
    const uint64_t g_first = base / 32, g_last = (base + n - 1) / 32;
    for (uint64_t g = g_first + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g <= g_last; g += (uint64_t)gridDim.x * blockDim.x)
    {
        uint32_t w0 = 0, w1 = 0, mk = 0;
        for (int t = 0; t < 32; t++)
        {
            const uint64_t pos = g * 32 + t;
            if (pos < base || pos >= base + n) continue;
            const unsigned char c = (unsigned char)ascii[pos - base];
            const uint32_t code = (c >> 1) & 3u;
            const bool ok = (c=='A'||c=='C'||c=='G'||c=='T'||c=='a'||c=='c'||c=='g'||c=='t');
            if (!ok) { mk |= 1u << t; bad++; }
            if (t < 16) w0 |= code << (2*t); else w1 |= code << (2*(t-16));
        }
        if (g == g_first || g == g_last) { atomicOr (&words[2*g], w0); atomicOr (&words[2*g+1], w1); atomicOr (&nmask[g], mk); }
        else { words[2*g] = w0; words[2*g+1] = w1; nmask[g] = mk; }
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync (FULL_MASK, bad, o);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd (n_invalid, bad);
}
cudaError_t launch_pack_ascii_at (const LaunchCtx& L, const char* ascii, uint64_t n, uint64_t base, uint32_t* packed_words, uint32_t* nmask,
                                  unsigned long long* n_invalid)
{
    if (n == 0) return cudaSuccess;
    const uint64_t groups = (base + n - 1) / 32 - base / 32 + 1;
    const uint64_t blocks = (groups + 255) / 256; const unsigned grid = (unsigned)(blocks < (uint64_t)L.sm_count * 32 ? blocks : (uint64_t)L.sm_count * 32);
    k_pack_ascii_at<<<grid, 256, 0, L.stream>>> (ascii, n, base, packed_words, nmask, n_invalid);
    (*L.launches)++;
    return cudaGetLastError ();
}
// offsets of a pushed batch: out[i] = in[i] - in[0] + base
__global__ void k_rebase_offsets (const uint64_t* in, uint64_t n, uint64_t base, uint64_t* out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] - in[0] + base;
}
cudaError_t launch_rebase_offsets (const LaunchCtx& L, const uint64_t* in, uint64_t n, uint64_t base, uint64_t* out)
{
    if (n == 0) return cudaSuccess;
    k_rebase_offsets<<<(unsigned)((n + 255) / 256), 256, 0, L.stream>>> (in, n, base, out);
    (*L.launches)++;
    return cudaGetLastError ();
}
