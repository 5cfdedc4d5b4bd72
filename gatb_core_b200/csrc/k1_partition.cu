// k1_partition.cu -- k1_superkmer_fast / k1_superkmer_partition: reads -> super-k-mer records scattered into HBM bins.
//
// Replaces (paths relative to /root/reference/gatb-core/src/gatb/):
//   ModelAbstract::iterate + ModelMinimizer::next         kmer/impl/Model.hpp:725-765, 1106-1139
//   Sequence2SuperKmer::operator() / KmerFunctor          kmer/impl/Sequence2SuperKmer.hpp:81-159
//   FillPartitions<span,true>::processSuperkmer           kmer/impl/SortingCountAlgorithm.cpp:1081-1151
//   SuperKmer::save (+ CacheSuperKmerBinFiles)            kmer/impl/Model.hpp:1386-1471, tools/storage/impl/Storage.cpp:567-580
//
// Two kernels (integer only, no tensor cores), both thread <-> read with a per-warp shared-memory event queue that the
// warp empties cooperatively (lane <-> closed super-k-mer: bin, one 32-bit global atomic for the slot, record built with
// funnel shifts, one 16-byte store), so that the divergent part runs at full lane occupancy:
//   * k1_superkmer_fast<WIN,W,HAS_N> -- the counting path (K1_MODE_DEVICE, k >= 15): the scan is K1Scanner (k1_scan.cuh),
//     compile-time indexed so that it lives in registers; records are cut out of a shared-memory ring of the words the
//     scanner has already seen.  Second half of this file.
//   * k1_superkmer_partition<W,HAS_N,MODE> -- the general kernel (GATB order, k < 15): 64-bit nucleotide shift register,
//     rolling forward / reverse-complement m-mers, sliding-window minimum over w = k-m+1 keys WITHOUT data-dependent
//     rescans (blocks of w keys: suffix minima of the previous block, prefix minima of the current one; the ring of w
//     keys per thread lives in shared memory laid out [slot][thread]); records are rebuilt from the packed stream.
//
// Two rank functions:
//   K1_MODE_GATB   : key = GATB's mmer_lut value (min(mmer, revcomp) or 4^m-1 when an "AA" sits anywhere but at the
//                    prefix; Model.hpp:1040-1064, 1220-1251) under integer '<' -- bit-exact super-k-mers and partitions
//                    p = repart[minimizer] (PartiInfo.hpp:323), pass = minimizer % nb_passes.
//   K1_MODE_DEVICE : key = canonical m-mer * odd + odd (m up to 16, k1s_key) -- a *random* minimizer order, whose
//                    buckets are balanced enough for one bin's distinct k-mers to fit a shared-memory hash table in k2b.
//                    Any order works for counting because every occurrence of a canonical k-mer has the same
//                    window minimum; GATB's own partition id is recomputed exactly for each emitted k-mer in k3.
#include "common.cuh"
#include "kernels.h"
#include "k1_scan.cuh"
#include "k2_common.cuh"      // mbarrier / TMA bulk copy helpers

#define K1_THREADS 128
#define K1_QCAP    384        // events per warp queue: flush threshold 96 + 8 steps x 32 lanes + slack
#define K1_QFLUSH  96

__device__ __forceinline__ uint32_t k1_key_device (uint32_t cm) { return k1s_key (cm); }

template<int W>
__device__ __forceinline__ void k1_store_record (const K1Params& P, uint32_t key, uint32_t meta, const uint64_t* roffs,
                                                 unsigned long long& stored, unsigned long long& dropped)
{
    const int      olane = meta & 31;
    const int      len   = (meta >> 5) & 63;
    const uint32_t start = meta >> 11;
    // ---- bin ----
    uint32_t bin, fine = 0;
    if (P.mode == K1_MODE_GATB)
    {
        uint32_t p = P.repart ? P.repart[key] : 0;
        bin = (key % (uint32_t)P.nb_passes) * (uint32_t)P.nb_partitions + p;
    }
    else
    {
        uint32_t h = mix32 (key);
        bin  = __umulhi (h, P.nb1);
        fine = (h * 0x9E3779B1u) >> (32 - P.fine_bits);
    }
    uint32_t slot = atomicAdd (&P.cursors[bin], 1u);
    if (P.count_only) return;
    uint64_t ridx = (uint64_t)bin * P.cap + slot;                         // GATB mode: bin-major
    if (P.bin_off)
    {
        const uint64_t o = P.bin_off[bin];
        if (slot >= P.bin_off[bin + 1] - o) { dropped++; return; }
        ridx = o + slot;
    }
    else
    {
        if (slot >= P.cap) { dropped++; return; }
        if (P.mode == K1_MODE_DEVICE)
        {
            const uint32_t region = bin / P.bins_per_region;
            ridx = (uint64_t)region * P.bins_per_region * P.cap + coarse_index (bin - region * P.bins_per_region, slot, P.bins_per_region);
        }
    }
    stored++;
    // ---- record: nn = k+len-1 nucleotides starting at stream position roff+start ----
    const uint64_t bp = 2 * (roffs[olane] + start);
    const uint64_t* wp = P.words + (bp >> 6);
    const int sh = (int)(bp & 63);
    const int nn = P.k + len - 1;
    if (W == 1)
    {
        uint64_t w0 = ldg64 (wp), w1 = ldg64 (wp + 1), w2 = ldg64 (wp + 2);
        uint64_t lo = sh ? ((w0 >> sh) | (w1 << (64 - sh))) : w0;
        uint64_t hi = sh ? ((w1 >> sh) | (w2 << (64 - sh))) : w1;
        if (nn <= 32) { lo &= mask2k64 (nn); hi = 0; } else { hi &= mask2k64 (nn - 32); }
        if (P.mode == K1_MODE_DEVICE) hi |= ((uint64_t)len << DEV_LEN_SHIFT_W1) | ((uint64_t)fine << DEV_FINE_SHIFT_W1);
        else                          hi |= ((uint64_t)len << REC_LEN_SHIFT_W1) | ((uint64_t)fine << REC_FINE_SHIFT_W1);
        uint4 rec = make_uint4 ((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
        ((uint4*)P.bins)[ridx] = rec;
    }
    else
    {
        uint64_t w[5];
        #pragma unroll
        for (int i=0; i<5; i++) w[i] = ldg64 (wp + i);
        uint64_t r[4];
        #pragma unroll
        for (int i=0; i<4; i++) r[i] = sh ? ((w[i] >> sh) | (w[i+1] << (64 - sh))) : w[i];
        #pragma unroll
        for (int i=0; i<4; i++)
        {
            int lo_nt = 32*i;                                   // first nucleotide held by word i
            if (nn <= lo_nt) r[i] = 0; else if (nn < lo_nt + 32) r[i] &= mask2k64 (nn - lo_nt);
        }
        r[3] |= ((uint64_t)len << REC_LEN_SHIFT_W2) | ((uint64_t)fine << REC_FINE_SHIFT_W2);
        uint4* dst = (uint4*)P.bins + 2 * ridx;
        dst[0] = make_uint4 ((uint32_t)r[0], (uint32_t)(r[0] >> 32), (uint32_t)r[1], (uint32_t)(r[1] >> 32));
        dst[1] = make_uint4 ((uint32_t)r[2], (uint32_t)(r[2] >> 32), (uint32_t)r[3], (uint32_t)(r[3] >> 32));
    }
}

template<int W, bool HAS_N, int MODE>
__global__ void __launch_bounds__(K1_THREADS) k1_superkmer_partition (const K1Params P)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int w = P.w, k = P.k, m = P.m;
    uint32_t* arr   = smem;                                                   // [w][K1_THREADS]
    uint32_t* q     = smem + w * K1_THREADS + wid * (K1_QCAP * 2);            // [K1_QCAP][2] per warp
    uint64_t* roffs = (uint64_t*)(smem + w * K1_THREADS + 4 * (K1_QCAP * 2)) + wid * 32;
    uint32_t* qcnt  = (uint32_t*)((uint64_t*)(smem + w * K1_THREADS + 4 * (K1_QCAP * 2)) + 4 * 32) + wid;

    const uint32_t mmask = P.mmask;
    const int shift_m = 2 * (m - 1);
    unsigned long long nvalid = 0, ninvalid = 0, stored = 0, dropped = 0;

    const uint64_t n_tiles = (P.n_reads + K1_THREADS - 1) / K1_THREADS;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
    {
        const uint64_t ri = tile * K1_THREADS + tid;
        const uint64_t r = P.first_read + ri;
        uint64_t roff = 0; int len = 0;
        if (ri < P.n_reads)
        {
            if (P.offsets) { roff = P.offsets[r]; len = (int)(P.offsets[r+1] - roff); }
            else           { roff = r * (uint64_t)P.read_len; len = P.read_len; }
        }
        roffs[lane] = roff;
        if (lane == 0) *qcnt = 0;
        int nm = (len >= k) ? (len - m + 1) : 0;            // reads shorter than k are skipped (Sequence2SuperKmer.hpp:144)
        const int nm_max = __reduce_max_sync (FULL_MASK, nm);
        __syncwarp ();

        // ---- nucleotide (and invalid-flag) shift registers ----
        const uint64_t* wp = P.words + (roff >> 5);
        uint64_t cur = 0; int avail = 0;
        const uint32_t* np = 0; uint32_t ncur = 0; int navail = 0;
        if (nm > 0)
        {
            int sh = (int)((2 * roff) & 63);
            cur = ldg64 (wp) >> sh; avail = (64 - sh) >> 1; wp++;
            if (HAS_N) { np = P.nmask + (roff >> 5); int s2 = (int)(roff & 31); ncur = __ldg (np) >> s2; navail = 32 - s2; np++; }
        }
        uint32_t f = 0, rr = 0;
        int since_bad = k;                                   // nucleotides since the last invalid one (saturating at "enough")
        #define NEXT_NT(c)                                                                      \
            { c = (uint32_t)cur & 3u; cur >>= 2; if (--avail == 0) { cur = ldg64 (wp); wp++; avail = 32; }          \
              if (HAS_N) { bool bad = ncur & 1u; ncur >>= 1; if (--navail == 0) { ncur = __ldg (np); np++; navail = 32; } \
                           since_bad = bad ? 0 : since_bad + 1; } }
        if (nm > 0)
            for (int i = 0; i < m - 1; i++)
            {
                uint32_t c; NEXT_NT (c);
                f  = ((f << 2) | c) & mmask;
                rr = (rr >> 2) | ((c ^ 2u) << shift_m);
            }

        // ---- main scan over m-mer index j; k-mer index i = j-(w-1) ----
        uint32_t prefix = 0xFFFFFFFFu; int t = 0;
        bool open = false; uint32_t sk_key = 0; int sk_start = 0, sk_len = 0;
        #define PUSH_EVENT()                                                                    \
            { uint32_t e = atomicAdd (qcnt, 1u); q[2*e] = sk_key; q[2*e+1] = ((uint32_t)sk_start << 11) | ((uint32_t)sk_len << 5) | (uint32_t)lane; }

        for (int j = 0; j < nm_max; j++)
        {
            const bool act = j < nm;
            uint32_t key = 0xFFFFFFFFu;
            if (act)
            {
                uint32_t c; NEXT_NT (c);
                f  = ((f << 2) | c) & mmask;
                rr = (rr >> 2) | ((c ^ 2u) << shift_m);
                uint32_t cm = min (f, rr);
                key = (MODE == K1_MODE_GATB) ? gatb_mmer_key (cm, mmask, P.mask_ma1) : k1_key_device (cm);
            }
            const uint32_t s = (t + 1 < w) ? arr[(t + 1) * K1_THREADS + tid] : 0xFFFFFFFFu;
            arr[t * K1_THREADS + tid] = key;
            prefix = min (prefix, key);
            const uint32_t wmin = min (prefix, s);
            if (act && j >= w - 1)
            {
                const bool valid = !HAS_N || since_bad >= k;
                if (valid)
                {
                    nvalid++;
                    if (!open || wmin != sk_key || sk_len == P.maxlen)
                    {
                        if (open) PUSH_EVENT ();
                        sk_start = j - (w - 1); sk_len = 0; sk_key = wmin; open = true;
                    }
                    sk_len++;
                }
                else
                {
                    ninvalid++;
                    if (open) { PUSH_EVENT (); open = false; }
                }
                if (j == nm - 1 && open) { PUSH_EVENT (); open = false; }
            }
            if (++t == w)
            {   // turn the finished block into suffix minima, in place
                uint32_t run = arr[(w - 1) * K1_THREADS + tid];
                for (int u = w - 2; u >= 0; u--) { run = min (run, arr[u * K1_THREADS + tid]); arr[u * K1_THREADS + tid] = run; }
                prefix = 0xFFFFFFFFu; t = 0;
            }
            if ((j & 7) == 7 || j == nm_max - 1)
            {
                __syncwarp ();
                const uint32_t nq = *qcnt;
                if (nq >= K1_QFLUSH || (j == nm_max - 1 && nq > 0))
                {
                    for (uint32_t e = lane; e < nq; e += 32)
                        k1_store_record<W> (P, q[2*e], q[2*e+1], roffs, stored, dropped);
                    __syncwarp ();
                    if (lane == 0) *qcnt = 0;
                }
                __syncwarp ();
            }
        }
        #undef NEXT_NT
        #undef PUSH_EVENT
        __syncwarp ();
    }
    // ---- statistics: one atomic per warp ----
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        nvalid   += __shfl_xor_sync (FULL_MASK, nvalid, o);
        ninvalid += __shfl_xor_sync (FULL_MASK, ninvalid, o);
        stored   += __shfl_xor_sync (FULL_MASK, stored, o);
        dropped  += __shfl_xor_sync (FULL_MASK, dropped, o);
    }
    if (lane == 0)
    {
        if (nvalid)   atomicAdd (&P.stats[0], nvalid);
        if (ninvalid) atomicAdd (&P.stats[1], ninvalid);
        if (stored)   atomicAdd (&P.stats[2], stored);
        if (dropped)  atomicAdd (&P.stats[3], dropped);
    }
}


// =====================================================================================================================
//  k1_superkmer_fast<WIN,W,HAS_N>: the device-order partition kernel (HAS_N: the reads may hold invalid nucleotides).
//  thread <-> read; the scan itself is K1Scanner (k1_scan.cuh: registers only, ~20 instructions per nucleotide);
//  closing lanes push 8-byte events into their warp's shared-memory queue; after every 16 positions the warp empties
//  the queue cooperatively (lane <-> event): bin from the key, one 32-bit global atomic for the slot, the record cut
//  out of the owner's shared-memory word ring with funnel shifts, one 16-byte store.  Per record the memory system
//  sees exactly two scattered operations (the atomic and the store); everything else is shared memory.
// =====================================================================================================================
#define K1F_QCAP 640           // events per warp queue: 18 per lane per step at the very worst

// ORI (k <= 31 only): 'key' is the strand-tagged window minimum t of k1_scan.cuh (bin from its rank t >> 1, class R when t is
// odd: the record holds the reverse complement of the span), meta = start << 12 | ambiguous << 11 | len << 5 | lane.  A
// super-k-mer flagged ambiguous is taken apart: every k-mer is classified exactly (k1s_classify_kmer) and leaves as a
// record of its own, ambiguous k-mers in their canonical orientation min(K, revcomp K) like kmer/impl/Model.hpp:857-884.
struct K1fBin { uint32_t bin, fine, lbin; uint64_t rbase; };

// one record: l k-mers from k-mer 'st' of the owner's read on; cls 0 = as read, 1 = reverse-complemented
template<int W, bool ORI, bool DENSE>
__device__ __forceinline__ void k1f_put (const K1Params& P, const K1fBin& B, const uint32_t* col, uint32_t st, int l, int cls,
                                         unsigned long long& stored, unsigned long long& dropped)
{
    const uint32_t slot = atomicAdd (&P.cursors[B.bin], 1u);
    uint64_t ridx;
    if (DENSE)
    {   // dense layout (second run after a counting run): exact room for every bin
        const uint64_t o = P.bin_off[B.bin];
        if (slot >= P.bin_off[B.bin + 1] - o) { dropped++; return; }
        ridx = o + slot;
    }
    else
    {
        if (slot >= P.cap) { dropped++; return; }
        ridx = B.rbase + coarse_index (B.lbin, slot, P.bins_per_region);
    }
    stored++;
    const uint32_t w0 = st >> 4;
    const int sh = 2 * (int)(st & 15);
    const int nn = P.k + l - 1;
    constexpr int NW = (W == 1) ? 5 : 9;
    uint32_t x[NW];
    #pragma unroll
    for (int i = 0; i < NW; i++) x[i] = col[((w0 + i) & (K1S_RING - 1)) * K1_THREADS];
    uint32_t r[NW - 1];
    #pragma unroll
    for (int i = 0; i < NW - 1; i++) r[i] = __funnelshift_r (x[i], x[i + 1], sh);
    if (W == 1)
    {
        uint64_t lo = (uint64_t)r[0] | ((uint64_t)r[1] << 32), hi = (uint64_t)r[2] | ((uint64_t)r[3] << 32);
        if (nn <= 32) { lo &= mask2k64 (nn); hi = 0; } else { hi &= mask2k64 (nn - 32); }
        if (ORI && cls) k1s_revcomp_span (lo, hi, nn);
        hi |= ((uint64_t)l << DEV_LEN_SHIFT_W1) | ((uint64_t)B.fine << DEV_FINE_SHIFT_W1);
        ((uint4*)P.bins)[ridx] = make_uint4 ((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
    }
    else
    {
        uint64_t q[4];
        #pragma unroll
        for (int i = 0; i < 4; i++)
        {
            q[i] = (uint64_t)r[2 * i] | ((uint64_t)r[2 * i + 1] << 32);
            const int lo_nt = 32 * i;
            if (nn <= lo_nt) q[i] = 0; else if (nn < lo_nt + 32) q[i] &= mask2k64 (nn - lo_nt);
        }
        q[3] |= ((uint64_t)l << REC_LEN_SHIFT_W2) | ((uint64_t)B.fine << REC_FINE_SHIFT_W2);
        uint4* dst = (uint4*)P.bins + 2 * ridx;
        dst[0] = make_uint4 ((uint32_t)q[0], (uint32_t)(q[0] >> 32), (uint32_t)q[1], (uint32_t)(q[1] >> 32));
        dst[1] = make_uint4 ((uint32_t)q[2], (uint32_t)(q[2] >> 32), (uint32_t)q[3], (uint32_t)(q[3] >> 32));
    }
}

// A super-k-mer flagged ambiguous (rare: hairpins, palindromic minimizers) is taken apart: every k-mer is classified
// exactly (k1s_classify_kmer) and leaves as a record of its own, ambiguous k-mers in their canonical orientation
// min(K, revcomp K) like kmer/impl/Model.hpp:857-884.  Kept out of line so that it costs the scan no registers.
template<int WIN, bool DENSE>
__device__ __noinline__ void k1f_store_ambiguous (const K1Params& P, const K1fBin& B, const uint32_t* col, uint32_t start, int len,
                                                  unsigned long long& stored, unsigned long long& dropped)
{
    auto word = [&] (int t) { return col[(t & (K1S_RING - 1)) * K1_THREADS]; };
    for (int i = 0; i < len; i++)
    {
        const uint32_t st = start + i;
        int cls = k1s_classify_kmer (word, (int)st, WIN, P.m);
        if (cls == 2)
        {   // canonical orientation of this one k-mer: the strand with the smaller VALUE
            const uint32_t w0 = st >> 4; const int sh = 2 * (int)(st & 15);
            const uint32_t a0 = word (w0), a1 = word (w0 + 1), a2 = word (w0 + 2);
            const uint64_t xs = ((uint64_t)__funnelshift_r (a0, a1, sh) | ((uint64_t)__funnelshift_r (a1, a2, sh) << 32)) & mask2k64 (P.k);
            const uint64_t rcv = xs ^ (0xAAAAAAAAAAAAAAAAULL & mask2k64 (P.k));
            const uint64_t fwv = pair_reverse64 (xs) >> (64 - 2 * P.k);
            cls = (rcv < fwv) ? 1 : 0;
        }
        k1f_put<1, true, DENSE> (P, B, col, st, 1, cls, stored, dropped);
    }
}

// ORI (k <= 31 only): 'key' is the strand-tagged window minimum t of k1_scan.cuh (bin from its rank t >> 1, class R when t is
// odd: the record holds the reverse complement of the span), meta = start << 12 | ambiguous << 11 | len << 5 | lane.
template<int W, bool ORI, int WIN, bool DENSE>
__device__ __forceinline__ void k1f_store (const K1Params& P, uint32_t key, uint32_t meta, const uint32_t* ring_w,
                                           unsigned long long& stored, unsigned long long& dropped)
{
    const int olane = meta & 31;
    int       len   = (meta >> 5) & 63;
    uint32_t  start = ORI ? (meta >> 12) : (meta >> 11);
    const uint32_t h    = mix32 (ORI ? (key >> 1) : key);
    K1fBin B;
    B.bin  = __umulhi (h, P.nb1);
    B.fine = (h * 0x9E3779B1u) >> (32 - P.fine_bits);
    const uint32_t region = __umulhi (h, P.n_regions);      // = bin / bins_per_region (nested floors)
    B.rbase = (uint64_t)region * P.bins_per_region * P.cap;
    B.lbin = B.bin - region * P.bins_per_region;
    const uint32_t* col = ring_w + olane;                   // the owner's column: word t at col[(t % K1S_RING) * K1_THREADS]
    if constexpr (ORI)
        if ((meta >> 11) & 1u) { k1f_store_ambiguous<WIN, DENSE> (P, B, col, start, len, stored, dropped); return; }
    while (len > 0)
    {
        const int l = len < P.maxlen ? len : P.maxlen;
        k1f_put<W, ORI, DENSE> (P, B, col, start, l, ORI ? (int)(key & 1u) : 0, stored, dropped);
        start += l; len -= l;
    }
}

template<int N> struct K1Int { static constexpr int value = N; };
template<int N, int I, class F> __device__ __forceinline__ void k1_static_for (F& f)
{
    if constexpr (I < N) { f (K1Int<I> ()); k1_static_for<N, I + 1> (f); }
}

struct K1Emit
{
    uint32_t* q; uint32_t* tail; uint32_t lane;
    __device__ __forceinline__ void operator() (uint32_t key, int start, int len)
    {
        const uint32_t e = atomicAdd (tail, 1u);
        *(uint2*)(q + 2 * e) = make_uint2 (key, ((uint32_t)start << 11) | ((uint32_t)len << 5) | lane);
    }
    __device__ __forceinline__ void operator() (uint32_t key, int start, int len, bool amb)      // oriented scan
    {
        const uint32_t e = atomicAdd (tail, 1u);
        *(uint2*)(q + 2 * e) = make_uint2 (key, ((uint32_t)start << 12) | ((uint32_t)amb << 11) | ((uint32_t)len << 5) | lane);
    }
};

// SM: the packed reads of a warp (32 consecutive reads = one contiguous byte range of the stream) are staged in shared memory
// by ONE TMA bulk copy (cp.async.bulk + mbarrier, SASS UBLKCP) issued by lane 0 one tile ahead (two buffers per warp); the scan
// reads its words with ld.shared.  Every input byte crosses HBM -> SM once, in 16-byte-aligned bursts, instead of being
// fetched word by word at a 37.5-byte stride per thread.  tile_bytes = dynamic shared memory per warp and buffer.
template<int WIN, int W, bool HAS_N, bool ORI, bool SM, bool DENSE>
__global__ void __launch_bounds__(K1_THREADS) k1_superkmer_fast (const K1Params P, const uint32_t tile_bytes)
{
    __shared__ __align__(16) uint32_t s_q[K1_THREADS / 32][K1F_QCAP * 2];
    __shared__ uint32_t s_ring[K1S_RING * K1_THREADS];
    __shared__ uint32_t s_tail[K1_THREADS / 32];
    __shared__ __align__(8) uint64_t s_bar[K1_THREADS / 32][2];
    __shared__ unsigned long long s_lo[K1_THREADS / 32][2];              // first stream byte held by the buffer
    extern __shared__ __align__(16) unsigned char k1_tiles[];            // [warp][2][tile_bytes]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int k = P.k, m = P.m;
    uint32_t* q = s_q[wid];
    const uint32_t* ring_w = s_ring + wid * 32;
    K1Emit emit; emit.q = q; emit.tail = &s_tail[wid]; emit.lane = (uint32_t)lane;
    unsigned long long nvalid = 0, ninvalid = 0, stored = 0, dropped = 0;
    if (lane == 0) s_tail[wid] = 0;
    __syncwarp ();

    // empties the warp's event queue (lane <-> event)
    auto drain = [&] ()
    {
        __syncwarp ();
        const uint32_t n = *(volatile uint32_t*)&s_tail[wid];
        for (uint32_t e0 = 0; e0 < n; e0 += 32)
        {
            const uint32_t e = e0 + lane;
            if (e < n)
            {
                const uint2 ev = *(const uint2*)(q + 2 * e);
                k1f_store<W, ORI, WIN, DENSE> (P, ev.x, ev.y, ring_w, stored, dropped);
            }
        }
        __syncwarp ();
        if (lane == 0) s_tail[wid] = 0;
        __syncwarp ();
    };

    const uint64_t n_tiles = (P.n_reads + K1_THREADS - 1) / K1_THREADS;
    // lane 0: one bulk copy for the 32 reads this warp scans in 'tile' (nothing when the slice is empty)
    auto stage = [&] (uint64_t tile, int buf)
    {
        const uint64_t ra = tile * K1_THREADS + (uint64_t)wid * 32;
        if (ra >= P.n_reads) { s_lo[wid][buf] = ~0ULL; return; }
        const uint64_t rb = (ra + 32 < P.n_reads) ? ra + 32 : P.n_reads;
        const uint64_t oa = P.offsets ? P.offsets[P.first_read + ra] : (P.first_read + ra) * (uint64_t)P.read_len;
        const uint64_t ob = P.offsets ? P.offsets[P.first_read + rb] : (P.first_read + rb) * (uint64_t)P.read_len;
        const uint64_t lo = (oa >> 2) & ~15ULL;
        uint64_t hi = (((ob + 3) >> 2) + 9 + 15) & ~15ULL;              // the scan looks two words past the last nucleotide
        if (hi - lo > tile_bytes) hi = lo + tile_bytes;                  // (cannot happen: the launcher sized tile_bytes for the longest read)
        s_lo[wid][buf] = lo;
        fence_proxy_async ();
        mbar_expect_tx (&s_bar[wid][buf], (uint32_t)(hi - lo));
        tma_bulk_g2s (k1_tiles + ((size_t)wid * 2 + buf) * tile_bytes, (const unsigned char*)P.words + lo, (uint32_t)(hi - lo), &s_bar[wid][buf]);
    };
    if (SM)
    {
        if (lane == 0) { mbar_init (&s_bar[wid][0], 1); mbar_init (&s_bar[wid][1], 1); }
        asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp ();
        if (lane == 0 && blockIdx.x < n_tiles) stage (blockIdx.x, 0);
        __syncwarp ();
    }
    uint32_t it = 0;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++)
    {
        const uint64_t ri = tile * K1_THREADS + tid;
        const uint64_t r = P.first_read + ri;
        uint64_t roff = 0; int len = 0;
        if (ri < P.n_reads)
        {
            if (P.offsets) { roff = P.offsets[r]; len = (int)(P.offsets[r+1] - roff); }
            else           { roff = r * (uint64_t)P.read_len; len = P.read_len; }
        }
        const int nm = (len >= k) ? (len - m + 1) : 0;      // reads shorter than k are skipped (Sequence2SuperKmer.hpp:144)
        const int nm_max = __reduce_max_sync (FULL_MASK, nm);
        K1Scanner<WIN, K1_THREADS, HAS_N, ORI, SM> sc;
        uint32_t tile_saddr = 0; unsigned long long tile_lo = 0;
        if (SM)
        {
            const int buf = (int)(it & 1);
            __syncwarp ();                                   // everybody is done with the other buffer (previous tile)
            if (lane == 0 && tile + gridDim.x < n_tiles) stage (tile + gridDim.x, buf ^ 1);
            __syncwarp ();
            tile_lo = s_lo[wid][buf];
            if (tile_lo != ~0ULL) mbar_wait (&s_bar[wid][buf], (it >> 1) & 1u);
            tile_saddr = smem_u32 (k1_tiles + ((size_t)wid * 2 + buf) * tile_bytes);
        }
        if (nm > 0)
        {
            nvalid += (unsigned long long)(len - k + 1);
            if (SM) sc.begin_shared (tile_saddr + (uint32_t)((((2 * roff) >> 5) << 2) - tile_lo), roff, len, m, s_ring + tid, P.nmask);
            else    sc.begin ((const uint32_t*)P.words, roff, len, m, s_ring + tid, P.nmask);
            if (sc.j >= nm) sc.finish (emit);
        }
        int j0 = WIN;
        auto body = [&] (auto ph)
        {
            constexpr int PH = decltype (ph)::value;
            if (j0 < nm_max)
            {
                if (j0 < nm)
                {
                    if (j0 + 16 <= nm) sc.template step16<PH, false> (emit);
                    else               sc.template step16<PH, true>  (emit);
                    if (j0 + 16 >= nm) sc.finish (emit);
                }
                j0 += 16;
                drain ();
            }
        };
        if (nm_max > 0 && nm_max <= WIN) drain ();          // reads of exactly k nucleotides: one k-mer each
        while (j0 < nm_max) k1_static_for<K1Scanner<WIN>::PHASES, 0> (body);
        if (HAS_N && nm > 0) { nvalid -= sc.ninv; ninvalid += sc.ninv; }
    }
    // ---- statistics: one atomic per warp ----
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        nvalid   += __shfl_xor_sync (FULL_MASK, nvalid, o);
        ninvalid += __shfl_xor_sync (FULL_MASK, ninvalid, o);
        stored   += __shfl_xor_sync (FULL_MASK, stored, o);
        dropped  += __shfl_xor_sync (FULL_MASK, dropped, o);
    }
    if (lane == 0)
    {
        if (nvalid)   atomicAdd (&P.stats[0], nvalid);
        if (ninvalid) atomicAdd (&P.stats[1], ninvalid);
        if (stored)   atomicAdd (&P.stats[2], stored);
        if (dropped)  atomicAdd (&P.stats[3], dropped);
    }
}

template<int WIN, int W, bool HAS_N, bool ORI, bool SM, bool DENSE>
static cudaError_t k1_fast_launch_s (const LaunchCtx& L, const K1Params& P, uint32_t tile_bytes)
{
    const size_t smem = SM ? (size_t)tile_bytes * 2 * (K1_THREADS / 32) : 0;
    cudaError_t e = cudaFuncSetAttribute (k1_superkmer_fast<WIN,W,HAS_N,ORI,SM,DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor (&per_sm, k1_superkmer_fast<WIN,W,HAS_N,ORI,SM,DENSE>, K1_THREADS, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const uint64_t n_tiles = (P.n_reads + K1_THREADS - 1) / K1_THREADS;
    uint64_t grid = (uint64_t)L.sm_count * per_sm;                 // persistent: a multiple of the SM count
    if (grid > n_tiles) grid = n_tiles;
    if (grid == 0) return cudaSuccess;
    k1_superkmer_fast<WIN,W,HAS_N,ORI,SM,DENSE><<<(unsigned)grid, K1_THREADS, smem, L.stream>>> (P, tile_bytes);
    (*L.launches)++;
    return cudaGetLastError ();
}
// staged in shared memory when the longest read is known and 32 of them fit 12 KB (reads of up to ~1500 nucleotides)
template<int WIN, int W, bool HAS_N, bool ORI>
static cudaError_t k1_fast_launch_n (const LaunchCtx& L, const K1Params& P)
{
    const uint64_t longest = P.offsets ? (uint64_t)P.max_len : (uint64_t)P.read_len;
    const uint64_t tile_bytes = (32 * ((longest + 3) / 4) + 64 + 15) & ~15ULL;
    if (P.bin_off) return k1_fast_launch_s<WIN, W, HAS_N, ORI, false, true> (L, P, 0);          // dense layout (second run of a skewed input)
    if (longest > 0 && tile_bytes <= 12288 && !P.no_staging) return k1_fast_launch_s<WIN, W, HAS_N, ORI, true, false> (L, P, (uint32_t)tile_bytes);
    return k1_fast_launch_s<WIN, W, HAS_N, ORI, false, false> (L, P, 0);
}
template<int WIN, int W>
static cudaError_t k1_fast_launch_t (const LaunchCtx& L, const K1Params& P)
{
    if constexpr (W == 1)
        if (P.oriented) return P.nmask ? k1_fast_launch_n<WIN, W, true, true> (L, P) : k1_fast_launch_n<WIN, W, false, true> (L, P);
    return P.nmask ? k1_fast_launch_n<WIN, W, true, false> (L, P) : k1_fast_launch_n<WIN, W, false, false> (L, P);
}

// window sizes the register scanner is compiled for: k in [WIN+7, WIN+15] with m = k-WIN+1 in [8,16]
int k1_fast_window (int k)
{
    if (k < 15 || k > 63) return 0;
    int win = ((k - 15) + 7) / 8 * 8;
    return win < 8 ? 8 : win;
}
static bool k1_fast_geometry_ok (int k, int m, int w)
{
    const int W = (k < 32) ? 1 : 2;
    if (w != k1_fast_window (k) || m != k - w + 1 || m < 8 || m > 16) return false;
    return (W == 1) ? (w == 8 || w == 16) : (w >= 24 && w <= 48);
}
static bool k1_fast_ok (const K1Params& P)
{
    if (P.mode != K1_MODE_DEVICE || P.count_only || P.force_general) return false;
    return k1_fast_geometry_ok (P.k, P.m, P.w);
}
// oriented records: the register scanner for k <= 31 unless a test asks for the canonical variant (path_flags of gatb_gpu.h)
bool k1_oriented (int k, int m, int w, int path_flags)
{ return k < 32 && !(path_flags & 1) && !(path_flags & 64) && k1_fast_geometry_ok (k, m, w); }

static size_t k1_smem_bytes (int w) { return (size_t)w * K1_THREADS * 4 + 4 * (K1_QCAP * 2) * 4 + 4 * 32 * 8 + 4 * 4 + 16; }

template<int W, bool HAS_N, int MODE>
static cudaError_t k1_launch_t (const LaunchCtx& L, const K1Params& P)
{
    size_t smem = k1_smem_bytes (P.w);
    cudaError_t e = cudaFuncSetAttribute (k1_superkmer_partition<W,HAS_N,MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor (&per_sm, k1_superkmer_partition<W,HAS_N,MODE>, K1_THREADS, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    uint64_t n_tiles = (P.n_reads + K1_THREADS - 1) / K1_THREADS;
    uint64_t grid = (uint64_t)L.sm_count * per_sm;                 // persistent: a multiple of the SM count
    if (grid > n_tiles) grid = n_tiles;
    if (grid == 0) return cudaSuccess;
    k1_superkmer_partition<W,HAS_N,MODE><<<(unsigned)grid, K1_THREADS, smem, L.stream>>> (P);
    (*L.launches)++;
    return cudaGetLastError ();
}

cudaError_t launch_k1 (const LaunchCtx& L, const K1Params& P)
{
    const int W = (P.k < 32) ? 1 : 2;
    if (k1_fast_ok (P))
        switch (P.w)
        {
            case 8:  return k1_fast_launch_t<8, 1>  (L, P);
            case 16: return k1_fast_launch_t<16, 1> (L, P);
            case 24: return k1_fast_launch_t<24, 2> (L, P);
            case 32: return k1_fast_launch_t<32, 2> (L, P);
            case 40: return k1_fast_launch_t<40, 2> (L, P);
            case 48: return k1_fast_launch_t<48, 2> (L, P);
        }
    const bool n = P.nmask != 0;
    if (P.mode == K1_MODE_GATB)
    {
        if (W == 1) return n ? k1_launch_t<1,true,K1_MODE_GATB> (L, P) : k1_launch_t<1,false,K1_MODE_GATB> (L, P);
        else        return n ? k1_launch_t<2,true,K1_MODE_GATB> (L, P) : k1_launch_t<2,false,K1_MODE_GATB> (L, P);
    }
    if (W == 1) return n ? k1_launch_t<1,true,K1_MODE_DEVICE> (L, P) : k1_launch_t<1,false,K1_MODE_DEVICE> (L, P);
    else        return n ? k1_launch_t<2,true,K1_MODE_DEVICE> (L, P) : k1_launch_t<2,false,K1_MODE_DEVICE> (L, P);
}
