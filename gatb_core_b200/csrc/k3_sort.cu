// k3_sort.cu -- k3: partition id of every emitted k-mer + ascending order inside each partition.
//
// Replaces (paths relative to /root/reference/gatb-core/src/gatb/):
//   Repartitor::operator()                            kmer/impl/PartiInfo.hpp:323   (p = table[minimizer(kmer)])
//   ModelMinimizer::computeNewMinimizerOriginal       kmer/impl/Model.hpp:1254-1287 (minimizer of one k-mer, GATB order)
//   the ascending emission order of executeDump / Hash16::iterator(true)
//                                                     kmer/impl/PartitionsCommand.cpp:1599-1805, 544
//   CountProcessorDump::process (append to dsk/solid/<pass*nb_partitions+part>)   kmer/impl/CountProcessorDump.hpp:148
//
// bucket of an emitted k-mer = key << t | (top t bits of the k-mer value), key = pass*nb_partitions + repart[minimizer] with
// the GATB minimizer found by rolling the forward/reverse m-mer along the k-mer (no table: lut value = min(mmer, revcomp)
// with the "AA" rule applied arithmetically).
// k3s  classify + scatter in ONE pass: a bucket is a chain of 32-item blocks from one bump allocator (block directory),
//      an item is one 16-byte store; four items in flight per thread.  The default.
// scan exclusive prefix sum over the bucket populations (u32 -> u64), tile scan + block-sum scan + add: output offsets.
// k3c  one CTA per bucket: distribution sort by k-mer value in shared memory (sub-bucket histogram, scan, scatter, rank among
//      the few neighbours), bitonic network for skewed buckets; writes to the final position.
// k3a / k3b  exact two-pass classify and scatter (atomic cursor per bucket): the fallback when a bucket outgrows the block
//      directory of k3s.
// k3d  buckets larger than the shared-memory budget: bitonic sort in global memory by one CTA (rare; skewed value ranges).
#include "common.cuh"
#include "kernels.h"

#define K3_SORT_CAP 4096
#define K3_ILP 4          // independent items a thread of the classify / scatter passes keeps in flight

// minimizer (GATB lexicographic order with the AA rule) of one canonical k-mer value
__device__ __forceinline__ uint32_t gatb_minimizer_w1 (uint64_t v, int k, int m, uint32_t mmask, uint32_t mask_ma1)
{
    uint32_t mm = (uint32_t)v & mmask;
    uint32_t rc = (pair_reverse32 (mm) ^ 0xAAAAAAAAu) >> (32 - 2*m);
    uint32_t best = gatb_mmer_key (min (mm, rc), mmask, mask_ma1);
    const int top = 2*(m - 1);
    for (int i = m; i < k; i++)
    {
        uint32_t nt = (uint32_t)(v >> (2*i)) & 3u;
        mm = (mm >> 2) | (nt << top);
        rc = ((rc << 2) | (nt ^ 2u)) & mmask;
        best = min (best, gatb_mmer_key (min (mm, rc), mmask, mask_ma1));
    }
    return best;
}
__device__ __forceinline__ uint32_t gatb_minimizer_w2 (u128 v, int k, int m, uint32_t mmask, uint32_t mask_ma1)
{
    uint32_t mm = (uint32_t)v.lo & mmask;
    uint32_t rc = (pair_reverse32 (mm) ^ 0xAAAAAAAAu) >> (32 - 2*m);
    uint32_t best = gatb_mmer_key (min (mm, rc), mmask, mask_ma1);
    const int top = 2*(m - 1);
    for (int i = m; i < k; i++)
    {
        uint32_t nt = (uint32_t)((i < 32 ? v.lo >> (2*i) : v.hi >> (2*(i - 32)))) & 3u;
        mm = (mm >> 2) | (nt << top);
        rc = ((rc << 2) | (nt ^ 2u)) & mmask;
        best = min (best, gatb_mmer_key (min (mm, rc), mmask, mask_ma1));
    }
    return best;
}

// Value range of a bucket.  A canonical k-mer is the smaller of two (nearly) independent values, so as a fraction u of the value
// range its density is 2(1-u): buckets cut at equal steps of the top bits would hold up to twice the average at the low end and
// almost nothing at the high end.  Any MONOTONE map of the value keeps the concatenation of sorted buckets sorted, so the
// bucket index follows the distribution function F(u) = 1 - (1-u)^2 = 2u - u^2 instead: u = the top U = min(2k, 31) bits,
// F in 2U bits (strictly increasing in u), the top 'bits' of it.  The sorting kernel takes its sub-buckets from the SAME F with
// more bits, so buckets and sub-buckets nest.  Exact integer arithmetic, one function for every kernel that needs it.
__device__ __forceinline__ int k3_range_U (int k) { return 2 * k < 31 ? 2 * k : 31; }
__device__ __forceinline__ uint64_t k3_range_of (uint32_t u, int U, int bits)
{
    const uint64_t F = (((uint64_t)u << (U + 1)) - (uint64_t)u * u);      // < 2^(2U)
    return F >> (2 * U - bits);
}
// the top U bits of a 2k-bit value
template<int W> __device__ __forceinline__ uint32_t k3_top_bits (uint64_t lo, uint64_t hi, int k, int U)
{
    const int s = 2 * k - U;
    if (W == 1) return (uint32_t)(lo >> s);
    return (uint32_t)(s >= 64 ? (hi >> (s - 64)) : (s ? ((lo >> s) | (hi << (64 - s))) : lo)) & (uint32_t)((1ULL << U) - 1);
}
// bucket = (partition key << t_bits) | range (top bits of the k-mer value)
__device__ __forceinline__ uint32_t k3_bucket_of (const K3Params& P, uint64_t lo, uint64_t hi)
{
    const int k = P.k, t = P.t_bits;
    const int U = k3_range_U (k);
    uint32_t key = 0, topbits = 0;
    if (P.W == 1)
    {
        if (P.n_keys > 1)
        {
            const uint32_t mini = gatb_minimizer_w1 (lo, k, P.m, P.mmask, P.mask_ma1);
            key = (mini % (uint32_t)P.nb_passes) * (uint32_t)P.nb_partitions + P.repart[mini];
        }
        if (t) topbits = (uint32_t) k3_range_of (k3_top_bits<1> (lo, 0, k, U), U, t);
    }
    else
    {
        u128 v; v.lo = lo; v.hi = hi;
        if (P.n_keys > 1)
        {
            const uint32_t mini = gatb_minimizer_w2 (v, k, P.m, P.mmask, P.mask_ma1);
            key = (mini % (uint32_t)P.nb_passes) * (uint32_t)P.nb_partitions + P.repart[mini];
        }
        if (t) topbits = (uint32_t) k3_range_of (k3_top_bits<2> (v.lo, v.hi, k, U), U, t);
    }
    return (key << t) | topbits;
}

// partition key of an emitted k-mer: pass * nb_partitions + repart[GATB minimizer]
__device__ __forceinline__ uint32_t k3_key_of (const K3Params& P, uint64_t lo, uint64_t hi)
{
    uint32_t mini;
    if (P.W == 1) mini = gatb_minimizer_w1 (lo, P.k, P.m, P.mmask, P.mask_ma1);
    else { u128 v; v.lo = lo; v.hi = hi; mini = gatb_minimizer_w2 (v, P.k, P.m, P.mmask, P.mask_ma1); }
    return (mini % (uint32_t)P.nb_passes) * (uint32_t)P.nb_partitions + P.repart[mini];
}
// (i = index of the item: with P.in_key the key travels with the item -- it was computed when the item was routed to the rank
//  that owns its partition, k3r_* below -- and is not computed again)
__device__ __forceinline__ uint32_t k3_bucket_of (const K3Params& P, uint64_t lo, uint64_t hi, uint64_t i)
{
    if (P.in_key)
    {
        const int U = k3_range_U (P.k);
        const uint32_t topbits = P.t_bits ? (uint32_t) k3_range_of (P.W == 1 ? k3_top_bits<1> (lo, 0, P.k, U) : k3_top_bits<2> (lo, hi, P.k, U), U, P.t_bits) : 0u;
        return ((uint32_t)P.in_key[i] << P.t_bits) | topbits;
    }
    return k3_bucket_of (P, lo, hi);
}

// ---- k3s: classify + scatter in ONE pass over the emitted k-mers ------------------------------------------------------
// Scattering 6*10^8 items into 10^6 buckets with exact offsets keeps one open line per bucket and array, spread over
// the whole multi-GB copy: every store is an address-translation miss (measured: the same scatter folded into a
// 256 MB window costs a quarter).  Here a bucket is a chain of K3_BLK-item blocks taken from ONE bump allocator, so
// all open blocks are the most recently allocated ones -- a window of n_buckets * 512 B whatever the skew between
// buckets -- and an item is one 16-byte (k <= 31) store instead of an 8- and a 4-byte one.  No counting pass is needed
// first: the item's slot comes from its bucket's cursor, the lane that opens a block (slot % K3_BLK == 0) takes it from
// the allocator and publishes it in dir[block number][bucket]; the others read it there (they can only have been
// given their slot after the opener got its own, so the entry is on its way).
// A bucket that outgrows dir_rounds * K3_BLK items raises ovf_flag: the caller then falls back to the exact two-pass
// path (k3a + k3b), which has no such limit.
__global__ void __launch_bounds__(256) k3s_pool_scatter (const K3Params P)
{
    const uint64_t base = (uint64_t)blockIdx.x * (256 * K3_ILP) + threadIdx.x;
    uint64_t lo[K3_ILP], hi[K3_ILP]; uint32_t c[K3_ILP], b[K3_ILP], slot[K3_ILP], blk[K3_ILP];
    bool ok[K3_ILP];
    #pragma unroll
    for (int j = 0; j < K3_ILP; j++)
    {
        const uint64_t i = base + 256 * j;
        lo[j] = i < P.n ? P.in_lo[i] : 0xFFFFFFFFFFFFFFFFULL;
        hi[j] = (P.W == 2) ? (i < P.n ? P.in_hi[i] : 0xFFFFFFFFFFFFFFFFULL) : 0;
        c[j] = i < P.n ? P.in_cnt[i] : 0;
    }
    const uint32_t cap = P.dir_rounds * K3_BLK;
    #pragma unroll
    for (int j = 0; j < K3_ILP; j++)
    {
        // holes left by k2b's block-wise output reservation carry an all-ones key: skip them
        ok[j] = (P.W == 1 ? lo[j] : hi[j]) != 0xFFFFFFFFFFFFFFFFULL;
        if (ok[j])
        {
            b[j] = k3_bucket_of (P, lo[j], hi[j], base + 256 * j);
            slot[j] = atomicAdd (&P.bucket_count[b[j]], 1u);
            if (slot[j] >= cap) { ok[j] = false; atomicOr (P.ovf_flag, 1u); }
        }
    }
    #pragma unroll
    for (int j = 0; j < K3_ILP; j++)
        if (ok[j] && (slot[j] % K3_BLK) == 0)
        {
            blk[j] = atomicAdd (P.pool_ptr, 1u);
            if (blk[j] >= P.pool_blocks) { ok[j] = false; atomicOr (P.ovf_flag, 2u); blk[j] = 0; }
            *(volatile uint32_t*)&P.dir[(uint64_t)(slot[j] / K3_BLK) * P.n_buckets + b[j]] = blk[j];
        }
    __syncwarp ();
    #pragma unroll
    for (int j = 0; j < K3_ILP; j++)
        if (ok[j] && (slot[j] % K3_BLK) != 0)
        {
            const volatile uint32_t* e = (const volatile uint32_t*)&P.dir[(uint64_t)(slot[j] / K3_BLK) * P.n_buckets + b[j]];
            uint32_t v;
            while ((v = *e) == 0xFFFFFFFFu) __nanosleep (20);
            blk[j] = v;
        }
    #pragma unroll
    for (int j = 0; j < K3_ILP; j++)
        if (ok[j])
        {
            const uint64_t at = (uint64_t)blk[j] * K3_BLK + (slot[j] % K3_BLK);
            if (P.W == 1) P.pool[at] = make_uint4 ((uint32_t)lo[j], (uint32_t)(lo[j] >> 32), c[j], 0u);
            else
            {
                P.pool[2 * at]     = make_uint4 ((uint32_t)lo[j], (uint32_t)(lo[j] >> 32), (uint32_t)hi[j], (uint32_t)(hi[j] >> 32));
                P.pool[2 * at + 1] = make_uint4 (c[j], 0u, 0u, 0u);
            }
        }
}

// Both passes are chains of dependent memory operations (load -> atomic -> store) on random addresses: every thread
// keeps K3_ILP independent items in flight so that the latencies overlap.
__global__ void __launch_bounds__(256) k3a_classify (const K3Params P)
{
    const uint64_t base = (uint64_t)blockIdx.x * (256 * K3_ILP) + threadIdx.x;
    uint64_t lo[K3_ILP], hi[K3_ILP];
    #pragma unroll
    for (int j = 0; j < K3_ILP; j++)
    {
        const uint64_t i = base + 256 * j;
        lo[j] = i < P.n ? P.in_lo[i] : 0xFFFFFFFFFFFFFFFFULL;
        hi[j] = (P.W == 2) ? (i < P.n ? P.in_hi[i] : 0xFFFFFFFFFFFFFFFFULL) : 0;
    }
    #pragma unroll
    for (int j = 0; j < K3_ILP; j++)
    {
        const uint64_t i = base + 256 * j;
        if (i >= P.n) continue;
        // holes left by k2b's block-wise output reservation carry an all-ones key: skip them
        if ((P.W == 1 ? lo[j] : hi[j]) == 0xFFFFFFFFFFFFFFFFULL) { P.bucket_of[i] = 0xFFFFFFFFu; continue; }
        const uint32_t b = k3_bucket_of (P, lo[j], hi[j], i);
        P.bucket_of[i] = b;
        atomicAdd (&P.bucket_count[b], 1u);
    }
}

// exact two-pass scatter (after k3a + scan): the fallback of the pooled single-pass scatter above
__global__ void __launch_bounds__(256) k3b_scatter (const K3Params P)
{
    const uint64_t base = (uint64_t)blockIdx.x * (256 * K3_ILP) + threadIdx.x;
    uint32_t b[K3_ILP]; uint64_t lo[K3_ILP], hi[K3_ILP], pos[K3_ILP]; uint32_t c[K3_ILP];
    #pragma unroll
    for (int j = 0; j < K3_ILP; j++) { const uint64_t i = base + 256 * j; b[j] = i < P.n ? P.bucket_of[i] : 0xFFFFFFFFu; }
    #pragma unroll
    for (int j = 0; j < K3_ILP; j++)
    {
        const uint64_t i = base + 256 * j;
        if (b[j] != 0xFFFFFFFFu)
        {
            pos[j] = P.bucket_off[b[j]] + atomicAdd (&P.bucket_count[b[j]], 1u);     // bucket_count was re-zeroed: it is the cursor now
            lo[j] = P.in_lo[i]; c[j] = P.in_cnt[i];
            if (P.W == 2) hi[j] = P.in_hi[i];
        }
    }
    #pragma unroll
    for (int j = 0; j < K3_ILP; j++)
        if (b[j] != 0xFFFFFFFFu)
        {
            P.tmp_lo[pos[j]] = lo[j];
            if (P.W == 2) P.tmp_hi[pos[j]] = hi[j];
            P.tmp_cnt[pos[j]] = c[j];
        }
}

// ---- per-bucket sort in shared memory ------------------------------------------------------------------------------
// The k-mers of a bucket share the top t_bits of the distribution function of canonical values (k3_range_of); below them
// it is close to uniform.  So the bucket is sorted by DISTRIBUTION: its next B bits (2^B >= n) pick a sub-bucket (shared-memory histogram, in-place scan, scatter with the
// scanned counters as cursors), and inside a sub-bucket (about one k-mer on average) every k-mer finds its rank by
// comparing itself with its few neighbours -- k-mers are distinct, so ranks are unique.  O(n) work, two block barriers.
// A bucket with a crowded sub-bucket (skewed values) takes the bitonic network instead.
#define K3_SUB_MAX 48
template<int W>
__device__ __forceinline__ uint32_t k3_sub_bucket (uint64_t lo, uint64_t hi, int k, int U, int bits, uint32_t mask)
{
    // the B bits of the distribution function below the bucket's own t_bits (bits = t_bits + B): monotone in the value
    return (uint32_t) k3_range_of (k3_top_bits<W> (lo, hi, k, U), U, bits) & mask;
}

template<int W>
__global__ void __launch_bounds__(256) k3c_sort (const K3Params P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* s_lo = (uint64_t*)smem_raw;
    uint64_t* s_hi = (W == 2) ? s_lo + K3_SORT_CAP : 0;
    uint32_t* s_c  = (uint32_t*)(s_lo + (size_t)K3_SORT_CAP * W);
    uint32_t* s_off = s_c + K3_SORT_CAP;                      // [K3_SORT_CAP] sub-bucket counters -> end offsets
    __shared__ uint32_t s_wsum[8];
    __shared__ uint32_t s_max;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // item i of bucket b: from the pooled blocks (k3s) or from the exact bucket-ordered copy (k3b)
    auto fetch = [&] (uint32_t b, uint64_t beg, int i, uint64_t& lo, uint64_t& hi, uint32_t& c)
    {
        if (P.pool)
        {
            const uint64_t at = (uint64_t)P.dir[(uint64_t)(i / K3_BLK) * P.n_buckets + b] * K3_BLK + (i % K3_BLK);
            if (W == 1) { const uint4 q = P.pool[at]; lo = (uint64_t)q.x | ((uint64_t)q.y << 32); hi = 0; c = q.z; }
            else { const uint4 q = P.pool[2 * at]; lo = (uint64_t)q.x | ((uint64_t)q.y << 32); hi = (uint64_t)q.z | ((uint64_t)q.w << 32); c = P.pool[2 * at + 1].x; }
        }
        else { lo = P.tmp_lo[beg + i]; hi = (W == 2) ? P.tmp_hi[beg + i] : 0; c = P.tmp_cnt[beg + i]; }
    };
    for (uint32_t b = P.bucket_begin + blockIdx.x; b < P.bucket_end; b += gridDim.x)
    {
        const uint64_t beg = P.bucket_off[b], end = P.bucket_off[b+1];
        const uint64_t n64 = end - beg;
        if (n64 == 0) continue;
        if (n64 > K3_SORT_CAP)
        {
            if (tid == 0) { unsigned long long idx = atomicAdd (&P.counters[0], 1ULL); P.big_list[idx] = b; }
            continue;
        }
        const int n = (int)n64;
        int np = 1, B = 0; while (np < n) { np <<= 1; B++; }
        // bits of the distribution function below the bucket's own t_bits: 2U - t_bits of them; the sub-bucket takes the top B of those
        const int U = k3_range_U (P.k);
        const int free_bits = 2 * U - P.t_bits;
        bool ranked = false;
        if (n == 1)
        {
            if (tid == 0) { uint64_t lo, hi; uint32_t c; fetch (b, beg, 0, lo, hi, c); P.out_lo[beg] = lo; if (W == 2) P.out_hi[beg] = hi; P.out_cnt[beg] = (int32_t)c; }
            continue;
        }
        if (free_bits >= B)
        {
            const int sbits = P.t_bits + B; const uint32_t mask = (uint32_t)np - 1;
            for (int i = tid; i < np; i += 256) s_off[i] = 0;
            if (tid == 0) s_max = 0;
            __syncthreads ();
            for (int i = tid; i < n; i += 256)
            {
                uint64_t lo, hi; uint32_t c; fetch (b, beg, i, lo, hi, c);
                atomicAdd (&s_off[k3_sub_bucket<W> (lo, hi, P.k, U, sbits, mask)], 1u);
            }
            __syncthreads ();
            // in-place exclusive scan of np counters: every thread owns np/256 consecutive ones (np >= 256) or one (np < 256)
            {
                const int per = np >= 256 ? np / 256 : 1;
                const int i0 = tid * per;
                uint32_t sum = 0, mx = 0;
                if (i0 < np) for (int u = 0; u < per; u++) { const uint32_t v = s_off[i0 + u]; sum += v; mx = v > mx ? v : mx; }
                uint32_t incl = sum;
                #pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync (FULL_MASK, incl, o); if (lane >= o) incl += y; }
                #pragma unroll
                for (int o = 16; o > 0; o >>= 1) { const uint32_t y = __shfl_xor_sync (FULL_MASK, mx, o); mx = y > mx ? y : mx; }
                if (lane == 31) s_wsum[wid] = incl;
                if (lane == 0 && mx) atomicMax (&s_max, mx);
                __syncthreads ();
                uint32_t run = incl - sum;
                for (int u = 0; u < wid; u++) run += s_wsum[u];
                if (i0 < np) for (int u = 0; u < per; u++) { const uint32_t v = s_off[i0 + u]; s_off[i0 + u] = run; run += v; }
                __syncthreads ();
            }
            if (s_max <= K3_SUB_MAX)
            {
                ranked = true;
                // scatter by sub-bucket (arrival order inside); s_off[sb] ends up as the END of sub-bucket sb
                for (int i = tid; i < n; i += 256)
                {
                    uint64_t lo, hi; uint32_t c; fetch (b, beg, i, lo, hi, c);
                    const uint32_t pos = atomicAdd (&s_off[k3_sub_bucket<W> (lo, hi, P.k, U, sbits, mask)], 1u);
                    s_lo[pos] = lo; if (W == 2) s_hi[pos] = hi; s_c[pos] = c;
                }
                __syncthreads ();
                for (int i = tid; i < n; i += 256)
                {
                    const uint64_t lo = s_lo[i], hi = (W == 2) ? s_hi[i] : 0;
                    const uint32_t sb = k3_sub_bucket<W> (lo, hi, P.k, U, sbits, mask);
                    const uint32_t e = s_off[sb], st0 = sb ? s_off[sb - 1] : 0;
                    uint32_t rank = 0;
                    for (uint32_t j = st0; j < e; j++)
                    {
                        const uint64_t ol = s_lo[j];
                        if (W == 1) rank += (ol < lo);
                        else { const uint64_t oh = s_hi[j]; rank += (oh < hi || (oh == hi && ol < lo)); }
                    }
                    const uint64_t dst = beg + st0 + rank;
                    P.out_lo[dst] = lo; if (W == 2) P.out_hi[dst] = hi; P.out_cnt[dst] = (int32_t)s_c[i];
                }
                __syncthreads ();
            }
        }
        if (ranked) continue;
        for (int i = threadIdx.x; i < np; i += blockDim.x)
        {
            if (i < n) { uint64_t lo, hi; uint32_t c; fetch (b, beg, i, lo, hi, c); s_lo[i] = lo; if (W == 2) s_hi[i] = hi; s_c[i] = c; }
            else       { s_lo[i] = ~0ULL; if (W == 2) s_hi[i] = ~0ULL; s_c[i] = 0; }
        }
        __syncthreads ();
        // all-ascending bitonic network: each merge starts with a "flip" step (partner = mirror inside the block),
        // then half-cleaners; +inf padding therefore never moves below index n
        for (int size = 2; size <= np; size <<= 1)
            for (int stride = size >> 1; stride > 0; stride >>= 1)
            {
                const bool flip = (stride == (size >> 1));
                for (int t = threadIdx.x; t < (np >> 1); t += blockDim.x)
                {
                    int lo_i, hi_i;
                    if (flip) { int blk = t / stride, r = t - blk * stride; lo_i = blk * size + r; hi_i = blk * size + size - 1 - r; }
                    else      { lo_i = 2*t - (t & (stride - 1)); hi_i = lo_i + stride; }
                    uint64_t al = s_lo[lo_i], bl = s_lo[hi_i];
                    uint64_t ah = (W == 2) ? s_hi[lo_i] : 0, bh = (W == 2) ? s_hi[hi_i] : 0;
                    bool gt = (W == 2) ? (ah > bh || (ah == bh && al > bl)) : (al > bl);
                    if (gt)
                    {
                        s_lo[lo_i] = bl; s_lo[hi_i] = al;
                        if (W == 2) { s_hi[lo_i] = bh; s_hi[hi_i] = ah; }
                        uint32_t c = s_c[lo_i]; s_c[lo_i] = s_c[hi_i]; s_c[hi_i] = c;
                    }
                }
                __syncthreads ();
            }
        for (int i = threadIdx.x; i < n; i += blockDim.x)
        { P.out_lo[beg + i] = s_lo[i]; if (W == 2) P.out_hi[beg + i] = s_hi[i]; P.out_cnt[beg + i] = (int32_t)s_c[i]; }
        __syncthreads ();
    }
}

// ---- oversized buckets: bitonic network directly in global memory (virtual padding with +inf) -------------------
template<int W>
__global__ void __launch_bounds__(1024) k3d_sort_big (const K3Params P, uint32_t n_big)
{
    for (uint32_t q = blockIdx.x; q < n_big; q += gridDim.x)
    {
        const uint32_t b = (uint32_t)P.big_list[q];
        const uint64_t beg = P.bucket_off[b], n = P.bucket_off[b+1] - beg;
        uint64_t np = 1; while (np < n) np <<= 1;
        uint64_t* lo = P.tmp_lo + beg; uint64_t* hi = (W == 2) ? P.tmp_hi + beg : 0; uint32_t* cn = P.tmp_cnt + beg;
        if (P.pool)
        {   // pooled scatter: the bucket's items sit in 32-item blocks; gather them into the output range and sort there
            lo = P.out_lo + beg; hi = (W == 2) ? P.out_hi + beg : 0; cn = (uint32_t*)P.out_cnt + beg;
            for (uint64_t i = threadIdx.x; i < n; i += blockDim.x)
            {
                const uint64_t at = (uint64_t)P.dir[(uint64_t)(i / K3_BLK) * P.n_buckets + b] * K3_BLK + (i % K3_BLK);
                if (W == 1) { const uint4 q4 = P.pool[at]; lo[i] = (uint64_t)q4.x | ((uint64_t)q4.y << 32); cn[i] = q4.z; }
                else { const uint4 q4 = P.pool[2 * at]; lo[i] = (uint64_t)q4.x | ((uint64_t)q4.y << 32); hi[i] = (uint64_t)q4.z | ((uint64_t)q4.w << 32); cn[i] = P.pool[2 * at + 1].x; }
            }
            __syncthreads ();
        }
        for (uint64_t size = 2; size <= np; size <<= 1)
            for (uint64_t stride = size >> 1; stride > 0; stride >>= 1)
            {
                const bool flip = (stride == (size >> 1));
                for (uint64_t t = threadIdx.x; t < (np >> 1); t += blockDim.x)
                {
                    uint64_t lo_i, hi_i;
                    if (flip) { uint64_t blk = t / stride, r = t - blk * stride; lo_i = blk * size + r; hi_i = blk * size + size - 1 - r; }
                    else      { lo_i = 2*t - (t & (stride - 1)); hi_i = lo_i + stride; }
                    if (hi_i >= n) continue;                    // partner is virtual +inf: ascending order already holds
                    uint64_t al = lo[lo_i], bl = lo[hi_i];
                    uint64_t ah = (W == 2) ? hi[lo_i] : 0, bh = (W == 2) ? hi[hi_i] : 0;
                    bool gt = (W == 2) ? (ah > bh || (ah == bh && al > bl)) : (al > bl);
                    if (gt)
                    {
                        lo[lo_i] = bl; lo[hi_i] = al;
                        if (W == 2) { hi[lo_i] = bh; hi[hi_i] = ah; }
                        uint32_t c = cn[lo_i]; cn[lo_i] = cn[hi_i]; cn[hi_i] = c;
                    }
                }
                __syncthreads ();
            }
        if (!P.pool)
            for (uint64_t i = threadIdx.x; i < n; i += blockDim.x)
            { P.out_lo[beg + i] = lo[i]; if (W == 2) P.out_hi[beg + i] = hi[i]; P.out_cnt[beg + i] = (int32_t)cn[i]; }
        __syncthreads ();
    }
}

cudaError_t launch_k3a_classify (const LaunchCtx& L, const K3Params& P)
{
    if (P.n == 0) return cudaSuccess;
    k3a_classify<<<(unsigned)((P.n + 256 * K3_ILP - 1) / (256 * K3_ILP)), 256, 0, L.stream>>> (P);
    (*L.launches)++;
    return cudaGetLastError ();
}
cudaError_t launch_k3b_scatter (const LaunchCtx& L, const K3Params& P)
{
    if (P.n == 0) return cudaSuccess;
    k3b_scatter<<<(unsigned)((P.n + 256 * K3_ILP - 1) / (256 * K3_ILP)), 256, 0, L.stream>>> (P);
    (*L.launches)++;
    return cudaGetLastError ();
}
// ---- k3r: routing of the emitted k-mers to the rank that owns their partition (several GPUs) ---------------------------------
// A k-mer lives in one device bin, hence on one rank; its GATB partition (key) is unrelated to that bin.  For the result of a
// partition to be ONE ascending sequence (what ICountProcessor::process sees in the reference) the k-mers are moved once more:
// key -> owner rank key % n_ranks.  k3r_route computes the key of every item (once: it travels with the item, 16 bits) and writes the
// items grouped by destination (SoA: value, count, key) into regions of dest_cap items each -- a block reserves its share of every
// destination with one atomic; the Repartitor balances the partitions, so the destinations fill evenly and a quarter of head-room
// suffices (a full region raises ovf_flag: the caller retries with regions that hold everything).  The caller exchanges the groups
// (all-to-all) and runs the sort stage on what it received.
__global__ void __launch_bounds__(256) k3r_route (const K3Params P, uint32_t n_ranks, uint64_t dest_cap, unsigned long long* __restrict__ dest_cursor,
                                                  uint64_t* __restrict__ o_lo, uint64_t* __restrict__ o_hi, uint32_t* __restrict__ o_cnt, uint16_t* __restrict__ o_key,
                                                  uint32_t* __restrict__ ovf_flag)
{
    __shared__ uint32_t s_cnt[8];
    __shared__ unsigned long long s_base[8];
    if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
    __syncthreads ();
    const uint64_t base = (uint64_t)blockIdx.x * (256 * K3_ILP) + threadIdx.x;
    uint64_t lo[K3_ILP], hi[K3_ILP];
    uint32_t key[K3_ILP], at[K3_ILP];
    #pragma unroll
    for (int j = 0; j < K3_ILP; j++)
    {
        const uint64_t i = base + 256 * j;
        lo[j] = i < P.n ? P.in_lo[i] : 0xFFFFFFFFFFFFFFFFULL;
        hi[j] = (P.W == 2) ? (i < P.n ? P.in_hi[i] : 0xFFFFFFFFFFFFFFFFULL) : 0;
    }
    #pragma unroll
    for (int j = 0; j < K3_ILP; j++)
    {
        key[j] = 0xFFFFFFFFu;                                  // holes left by k2b's block-wise output reservation: skipped
        if ((P.W == 1 ? lo[j] : hi[j]) != 0xFFFFFFFFFFFFFFFFULL)
        {
            key[j] = P.n_keys > 1 ? k3_key_of (P, lo[j], hi[j]) : 0u;
        }
        // slot inside the block's share of the destination: one shared atomic per (warp, destination), not per item
        const uint32_t d = key[j] == 0xFFFFFFFFu ? 0xFFu : key[j] % n_ranks;
        const unsigned peers = __match_any_sync (0xFFFFFFFFu, d);
        if (d != 0xFFu)
        {
            const int leader = __ffs (peers) - 1;
            uint32_t b0 = 0;
            if ((int)(threadIdx.x & 31) == leader) b0 = atomicAdd (&s_cnt[d], (uint32_t)__popc (peers));
            at[j] = __shfl_sync (peers, b0, leader) + __popc (peers & ((1u << (threadIdx.x & 31)) - 1));
        }
    }
    __syncthreads ();
    // (the cursors of the destinations sit in separate 128-byte lines: 10^6 blocks x n_ranks atomics on ONE line took 70 ms on 8 GPUs)
    if (threadIdx.x < n_ranks) s_base[threadIdx.x] = s_cnt[threadIdx.x] ? atomicAdd (&dest_cursor[threadIdx.x * 16], (unsigned long long)s_cnt[threadIdx.x]) : 0ULL;
    __syncthreads ();
    #pragma unroll
    for (int j = 0; j < K3_ILP; j++)
    {
        if (key[j] == 0xFFFFFFFFu) continue;
        const uint32_t d = key[j] % n_ranks;
        const unsigned long long slot = s_base[d] + at[j];
        if (slot >= dest_cap) { atomicOr (ovf_flag, 1u); continue; }           // the region of this destination is full: the caller retries with exact sizes
        const unsigned long long pos = (unsigned long long)d * dest_cap + slot;
        o_lo[pos] = lo[j]; if (P.W == 2) o_hi[pos] = hi[j];
        o_cnt[pos] = P.in_cnt[base + 256 * j]; o_key[pos] = (uint16_t)(key[j] / n_ranks);       // index among the keys its owner holds
    }
}
cudaError_t launch_k3r_route (const LaunchCtx& L, const K3Params& P, uint32_t n_ranks, uint64_t dest_cap, unsigned long long* dest_cursor,
                              uint64_t* o_lo, uint64_t* o_hi, uint32_t* o_cnt, uint16_t* o_key, uint32_t* ovf_flag)
{
    if (P.n == 0) return cudaSuccess;
    const uint64_t grid = (P.n + 256 * K3_ILP - 1) / (256 * K3_ILP);
    k3r_route<<<(unsigned)grid, 256, 0, L.stream>>> (P, n_ranks, dest_cap, dest_cursor, o_lo, o_hi, o_cnt, o_key, ovf_flag);
    (*L.launches)++;
    return cudaGetLastError ();
}
cudaError_t launch_k3s_pool_scatter (const LaunchCtx& L, const K3Params& P)
{
    if (P.n == 0) return cudaSuccess;
    k3s_pool_scatter<<<(unsigned)((P.n + 256 * K3_ILP - 1) / (256 * K3_ILP)), 256, 0, L.stream>>> (P);
    (*L.launches)++;
    return cudaGetLastError ();
}
uint32_t k3_sort_cap () { return K3_SORT_CAP; }
cudaError_t launch_k3c_sort (const LaunchCtx& L, const K3Params& P)
{
    if (P.n == 0) return cudaSuccess;
    size_t smem = (size_t)K3_SORT_CAP * (8 * P.W + 4 + 4);
    const void* fn = P.W == 1 ? (const void*)k3c_sort<1> : (const void*)k3c_sort<2>;
    cudaError_t e = cudaFuncSetAttribute (fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const unsigned nb = P.bucket_end - P.bucket_begin;
    if (nb == 0) return cudaSuccess;
    unsigned grid = nb < (unsigned)(L.sm_count * 32) ? nb : (unsigned)(L.sm_count * 32);
    if (P.W == 1) k3c_sort<1><<<grid, 256, smem, L.stream>>> (P); else k3c_sort<2><<<grid, 256, smem, L.stream>>> (P);
    (*L.launches)++;
    return cudaGetLastError ();
}
cudaError_t launch_k3d_sort_big (const LaunchCtx& L, const K3Params& P, uint32_t n_big)
{
    if (n_big == 0) return cudaSuccess;
    unsigned grid = n_big < (unsigned)L.sm_count ? n_big : (unsigned)L.sm_count;
    if (P.W == 1) k3d_sort_big<1><<<grid, 1024, 0, L.stream>>> (P, n_big); else k3d_sort_big<2><<<grid, 1024, 0, L.stream>>> (P, n_big);
    (*L.launches)++;
    return cudaGetLastError ();
}

// ---- exclusive scan u32 -> u64 ----------------------------------------------------------------------------------
#define SCAN_TILE 2048     // elements per block (256 threads x 8)
__global__ void __launch_bounds__(256) scan_tile_sums (const uint32_t* __restrict__ in, uint64_t n, uint64_t* __restrict__ sums)
{
    __shared__ unsigned long long s[8];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
    unsigned long long v = 0;
    for (int i = 0; i < 8; i++) { uint64_t idx = base + threadIdx.x + 256*i; if (idx < n) v += in[idx]; }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync (FULL_MASK, v, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads ();
    if (threadIdx.x == 0) { unsigned long long t = 0; for (int i = 0; i < 8; i++) t += s[i]; sums[blockIdx.x] = t; }
}
// single block: exclusive scan of m values in place (m arbitrary; sequential over chunks of 1024)
__global__ void __launch_bounds__(1024) scan_sums_inplace (uint64_t* sums, uint64_t m)
{
    __shared__ unsigned long long s_w[32]; __shared__ unsigned long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads ();
    for (uint64_t c0 = 0; c0 < m; c0 += 1024)
    {
        uint64_t idx = c0 + threadIdx.x;
        unsigned long long v = idx < m ? sums[idx] : 0, incl = v;
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned long long y = __shfl_up_sync (FULL_MASK, incl, o); if (lane >= o) incl += y; }
        if (lane == 31) s_w[wid] = incl;
        __syncthreads ();
        if (wid == 0) { unsigned long long w = s_w[lane], wi = w; for (int o = 1; o < 32; o <<= 1) { unsigned long long y = __shfl_up_sync (FULL_MASK, wi, o); if (lane >= o) wi += y; } s_w[lane] = wi - w; }
        __syncthreads ();
        unsigned long long excl = s_carry + s_w[wid] + incl - v;
        if (idx < m) sums[idx] = excl;
        __syncthreads ();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads ();
    }
}
__global__ void __launch_bounds__(256) scan_tile_apply (const uint32_t* __restrict__ in, uint64_t n, const uint64_t* __restrict__ sums, uint64_t* __restrict__ out)
{
    // each thread owns 8 consecutive elements of the tile
    __shared__ unsigned long long s_w[8];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * 8;
    uint32_t v[8]; unsigned long long sum = 0;
    for (int i = 0; i < 8; i++) { v[i] = base + i < n ? in[base + i] : 0; sum += v[i]; }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned long long incl = sum;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned long long y = __shfl_up_sync (FULL_MASK, incl, o); if (lane >= o) incl += y; }
    if (lane == 31) s_w[wid] = incl;
    __syncthreads ();
    unsigned long long woff = 0; for (int i = 0; i < wid; i++) woff += s_w[i];
    unsigned long long run = sums[blockIdx.x] + woff + incl - sum;
    for (int i = 0; i < 8; i++) { if (base + i < n) out[base + i] = run; run += v[i]; }
    // the element one past the end receives the grand total
    if (blockIdx.x == gridDim.x - 1)
    {
        __syncthreads ();
        if (threadIdx.x == 255) out[n] = run;      // thread 255 of the last tile: run == total (elements past n are zero)
    }
}
uint64_t scan_scratch_elems (uint64_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE + 1; }
// out has n+1 entries (out[n] = total)
cudaError_t launch_scan_u32_to_u64 (const LaunchCtx& L, const uint32_t* in, uint64_t* out, uint64_t n, uint64_t* scratch)
{
    if (n == 0) return cudaMemsetAsync (out, 0, sizeof(uint64_t), L.stream);
    unsigned tiles = (unsigned)((n + SCAN_TILE - 1) / SCAN_TILE);
    scan_tile_sums<<<tiles, 256, 0, L.stream>>> (in, n, scratch);
    scan_sums_inplace<<<1, 1024, 0, L.stream>>> (scratch, tiles);
    scan_tile_apply<<<tiles, 256, 0, L.stream>>> (in, n, scratch, out);
    (*L.launches) += 3;
    return cudaGetLastError ();
}
