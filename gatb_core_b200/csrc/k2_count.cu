// k2_count.cu -- fine split (k2a) and per-bin counting (k2b tiers, global-memory fallback k2c).
//
// Replaces (paths relative to /root/reference/gatb-core/src/gatb/):
//   ReadSuperKCommand::execute / hash-mode decode    kmer/impl/PartitionsCommand.cpp:944-1128, 420-501
//   SortCommand + executeDump (sort, merge, count)   kmer/impl/PartitionsCommand.cpp:1400-1445, 1599-1805
//   Hash16::insert (over-memory partitions)          tools/collections/impl/Hash16.hpp:198-230
//   CountProcessorHistogram::process                 kmer/impl/CountProcessorHistogram.hpp:173 (Histogram::inc, Histogram.hpp:92)
//   CountProcessorSoliditySum::check                 kmer/impl/CountProcessorSolidity.hpp:186
//
// k2a_fine_split      one CTA per coarse bin (gathered from up to 16 source pieces, walked as one record range): pass 1
//                     counts the records per fine bin in shared memory, pass 2 scatters them (16-byte loads / stores).
// k2b_warp_bins       k <= 31, the default: one WARP owns a fine bin (private 512-slot table, no block barrier); lane <->
//                     chunk of four k-mers (k2_decode.cuh), converged probe / 64-bit CAS claim / count steps.
// k2b_count_w1        k <= 31, CTA per bin with TMA-staged records (cp.async.bulk + mbarrier) and the same chunked
//                     insert: the second and third TIER for bins that overflow a warp's table (2048, then 8192 slots,
//                     bins taken from a list), and a test variant of the first tier (GATB_GPU_K2B=128|256).
// k2b_bucket_hash_count<W>  CTA per bin, one k-mer per lane: the path of 32 <= k <= 63 (128-bit keys claimed through a LOCK
//                     value in the high half), test variant for k <= 31 (GATB_GPU_K2B=0).
// k2b_warp_bins_w2    warp per bin for 32 <= k <= 63 (opt-in, see k2b_default_table_log2).
// k2c_*               what overflows every tier shares one global-memory table.
// All of them feed the abundance histogram (shared-memory bins, flushed once per CTA), the solidity statistics, and
// append the k-mers inside [emit_min, emit_max] to per-warp blocks of the output.
#include "common.cuh"
#include "kernels.h"
#include "k2_decode.cuh"
#include "k2_common.cuh"
#include <stdlib.h>

// ------------------------------------------------------------------------------------------------ k2a
// After the fine split the fine-bin id of a 16-byte record is implied by its position: the field carries the MULTIPLICITY of
// the record instead (1 here; k2a_dedup_split collapses identical records).  Every k <= 31 counting kernel adds it.
// PRE-SPLIT mode (PS.sub_off != NULL, k <= 31, several ranks): 'fine_bits' is the number of LEADING id bits to split by (the id
// shifted right by PS.shift); the records are copied verbatim, and instead of descriptors the kernel writes the first record
// (absolute) and the record count of every sub-bin -- the dense source k2a_dedup_split then takes its bins from.
// desc_abs == 0: descriptor of fine bin f of coarse bin b at bin_desc[b << fine_bits | f] = {offset inside the coarse bin, records}
// desc_abs != 0: (bins listed in bin_list, after k2a_dedup_split gave up on them) descriptors at bin_desc[desc_base +
//                (blockIdx.x << fine_bits | f)] = {ABSOLUTE record offset, records}, like the ones k2a_dedup_split writes itself.
template<int W>
__global__ void __launch_bounds__(256) k2a_fine_split (const K2aSrc S, uint4* __restrict__ dst, const uint64_t* __restrict__ coarse_off,
                                                        uint32_t nb, uint32_t cap, int fine_bits, uint2* __restrict__ bin_desc, const uint32_t* __restrict__ bin_list,
                                                        int desc_abs, uint64_t desc_base, K2aPresplit PS)
{
    // nb = bins of a source region (laid out by coarse_index, kernels.h)
    extern __shared__ uint32_t k2a_fs_smem[];                             // 3 << fine_bits counters
    const int nf = 1 << fine_bits;
    uint32_t* s_off = k2a_fs_smem; uint32_t* s_cur = s_off + nf; uint32_t* s_tmp = s_cur + nf;
    __shared__ uint32_t s_first[K2A_MAXSRC + 1];                        // record range of every source inside the gathered bin
    const uint32_t b = bin_list ? bin_list[blockIdx.x] : blockIdx.x;
    const int tid = threadIdx.x;
    for (int i = tid; i < nf; i += blockDim.x) { s_tmp[i] = 0; s_cur[i] = 0; }
    if (tid == 0)
    {
        uint32_t run = 0;
        for (int s = 0; s < S.n; s++) { s_first[s] = run; run += min (S.cursors[s][b], cap); }
        for (int s = S.n; s <= K2A_MAXSRC; s++) s_first[s] = run;
    }
    __syncthreads ();
    const uint32_t n_all = s_first[K2A_MAXSRC];
    // record g of the gathered bin: the pieces of all sources laid end to end, so that the threads stay busy however
    // small a single piece is (32 sources of a few hundred records each on 8 GPUs)
    auto locate = [&] (uint32_t g, const uint4*& src, uint64_t& ci)
    {
        int s = 0;
        #pragma unroll
        for (int u = K2A_MAXSRC / 2; u > 0; u >>= 1) if (s + u < K2A_MAXSRC && g >= s_first[s + u]) s += u;
        src = S.bins[s];
        ci = k2a_record_index (S, s, b, g - s_first[s], nb);
    };
    // pass 1: records per fine bin (the second pass finds the same lines in L2)
    for (uint32_t g = tid; g < n_all; g += blockDim.x)
    {
        const uint4* src; uint64_t ci; locate (g, src, ci);
        const uint32_t top = __ldg (&src[ci * W + (W - 1)]).w;
        atomicAdd (&s_tmp[W == 1 ? (((top >> (DEV_FINE_SHIFT_W1 - 32)) >> PS.shift) & (nf - 1)) : (top >> (32 - FINE_BITS_W2))], 1u);
    }
    __syncthreads ();
    const uint64_t dbase = coarse_off[b];
    if (tid < 32)
    {   // exclusive scan of the nf counters by one warp (nf/32 consecutive ones per lane, at least one)
        const int per = nf > 32 ? nf / 32 : 1;
        uint32_t sum = 0;
        for (int i = 0; i < per; i++) { const int idx = tid * per + i; if (idx < nf) sum += s_tmp[idx]; }
        uint32_t incl = sum;
        #pragma unroll
        for (int o=1; o<32; o<<=1) { uint32_t y = __shfl_up_sync (FULL_MASK, incl, o); if (tid >= o) incl += y; }
        uint32_t run = incl - sum;
        for (int i = 0; i < per; i++)
        {
            const int idx = tid * per + i;
            if (idx < nf)
            {
                const uint32_t v = s_tmp[idx]; s_off[idx] = run;
                if (PS.sub_off) { PS.sub_off[((uint64_t)b << fine_bits) + idx] = dbase + run; PS.sub_cnt[((uint64_t)b << fine_bits) + idx] = v; }
                else if (desc_abs) bin_desc[desc_base + ((uint64_t)blockIdx.x << fine_bits) + idx] = make_uint2 ((uint32_t)(dbase + run), v);
                else          bin_desc[((uint64_t)b << fine_bits) + idx] = make_uint2 (run, v);
                run += v;
            }
        }
    }
    __syncthreads ();
    for (uint32_t g = tid; g < n_all; g += blockDim.x)
    {
        const uint4* src; uint64_t ci; locate (g, src, ci);
        if (W == 1)
        {
            uint4 rec = __ldg (&src[ci]);
            uint32_t f = ((rec.w >> (DEV_FINE_SHIFT_W1 - 32)) >> PS.shift) & (nf - 1);
            uint32_t p = s_off[f] + atomicAdd (&s_cur[f], 1u);
            if (!PS.sub_off) rec.w = (rec.w & ((1u << (DEV_FINE_SHIFT_W1 - 32)) - 1)) | (1u << (DEV_FINE_SHIFT_W1 - 32));     // (pre-split: the id stays)
            dst[dbase + p] = rec;
        }
        else
        {
            uint4 r0 = __ldg (&src[2*ci]), r1 = __ldg (&src[2*ci + 1]);
            uint32_t f = r1.w >> (32 - FINE_BITS_W2);
            uint32_t p = s_off[f] + atomicAdd (&s_cur[f], 1u);
            dst[2*(dbase + p)] = r0; dst[2*(dbase + p) + 1] = r1;
        }
    }
}

// ------------------------------------------------------------------------------------------------ k2a with deduplication
// k <= 31.  One CTA stages a whole gathered coarse bin in shared memory (read ONCE from HBM), collapses identical records
// (oriented records of the reads that cover one locus without an error in the span are bit-identical, whatever the
// strand: about half of all records), and writes every distinct record once, ordered by fine id, with its multiplicity
// in place of the fine id.  The counting kernels then insert a record's k-mers once, adding the multiplicity.
//   dedup table: TS slots of {record index : 16, count : 16}; a slot is claimed with a 32-bit CAS, duplicates add 1 << 16.
// ADAPTIVE BINS: the fine ids written by the partition kernel are several times finer than a counting bin.  Knowing the k-mers
// of the surviving records per fine id, the CTA cuts the sequence of ids wherever the running sum passes a multiple of
// 'target': consecutive ids are merged into bins of <= target k-mer occurrences plus one id's worth, so the counting kernel
// sees bins of even load (hashed minimizers alone give bins whose load is a compound Poisson sum of a few loci: one in a
// hundred used to overflow a warp's table on one GPU, one in eleven on multi-Gb inputs) and the tables can be planned fuller.
// A bin that still ends up with more than 'big_load' k-mers (one fine id can hold several loci of a multi-Gb genome) is flagged
// (K2_DESC_BIG in the count): the first-tier counting kernel hands it to the next tier without trying.
// Descriptors {absolute record offset, records} are appended to bin_desc through counters[2] (the order of the bins is free).
// Shared memory: records rmax * 16 | table ts * 4 | per-id arrays 4 * nf * 4.  A bin larger than the staging area is taken in
// several passes over ranges of fine ids; bins the kernel gives up on are listed for the plain two-pass kernel above.
template<int NT>
__global__ void __launch_bounds__(NT, 1024 / NT) k2a_dedup_split (const K2aSrc S, uint4* __restrict__ dst, const uint64_t* __restrict__ coarse_off,
                                                       uint32_t nb, uint32_t cap, int fine_bits, uint2* __restrict__ bin_desc,
                                                       uint32_t rmax, uint32_t ts, uint32_t* __restrict__ big_list, unsigned long long* __restrict__ counters,
                                                       uint32_t target, uint32_t big_load)
{
    extern __shared__ __align__(16) unsigned char k2a_smem[];
    uint4*    recs  = (uint4*)k2a_smem;
    uint32_t* tbl   = (uint32_t*)(recs + rmax);
    const int nf = 1 << fine_bits;
    const uint32_t idmask = (uint32_t)nf - 1;          // (after a pre-split the leading id bits are implied by the bin)
    uint32_t* s_off = tbl + ts;          // first record of the id inside the coarse bin
    uint32_t* s_cur = s_off + nf;        // k-mers of the surviving records of the id -> their exclusive prefix inside a chunk of 32 ids
    uint32_t* s_tmp = s_cur + nf;        // surviving records of the id -> write cursor of the id
    uint32_t* s_q   = s_tmp + nf;        // 1: the id starts a counting bin
    __shared__ uint32_t s_first[K2A_MAXSRC + 1];
    __shared__ uint32_t s_staged, s_written, s_nb, s_nb2;
    __shared__ uint32_t s_ct[128], s_ckt[128];                  // per chunk of 32 fine ids: surviving records, their k-mers
    __shared__ unsigned long long s_gbase;
    __shared__ unsigned long long s_pass_base[16]; __shared__ uint32_t s_pass_nb[16];      // descriptors appended by every pass of the current bin
    const int tid = threadIdx.x;
    unsigned long long n_unique = 0;
    for (uint32_t b = blockIdx.x; b < nb; b += gridDim.x)
    {
        __syncthreads ();
        if (tid == 0)
        {
            uint32_t run = 0;
            for (int s = 0; s < S.n; s++) { s_first[s] = run; run += min (S.cursors[s][b], cap); }
            for (int s = S.n; s <= K2A_MAXSRC; s++) s_first[s] = run;
            s_written = 0;
        }
        __syncthreads ();
        const uint32_t n_all = s_first[K2A_MAXSRC];
        // A bin that fits the staging area is done in one pass.  A larger one (several ranks gather into one bin; dense bin loads)
        // takes P passes over ranges of fine ids: pass j stages only the records of its range -- the passes after the first
        // find the bin in L2 -- so that identical records still collapse whatever the size of the bin.  (fine ids are hashed:
        // a range holds n_all / P records give or take a few per cent; 75 % of the staging area is planned.)
        const uint32_t P = n_all <= rmax ? 1u : (uint32_t)(((uint64_t)n_all * 4 + 3 * rmax - 1) / (3 * (uint64_t)rmax));
        if (P > (uint32_t)nf || P > 16u)
        {   // more passes than fine ids (or than this CTA keeps track of): handed to the two-pass kernel
            if (tid == 0) { const uint32_t idx = (uint32_t) atomicAdd (&counters[0], 1ULL); big_list[idx] = b; }
            continue;
        }
        auto src_of = [&] (uint32_t g) -> const uint4*
        {
            if (S.n == 1) return S.bins[0] + k2a_record_index (S, 0, b, g, nb);          // one GPU: no search among the sources
            int s = 0;
            #pragma unroll
            for (int u = K2A_MAXSRC / 2; u > 0; u >>= 1) if (s + u < K2A_MAXSRC && g >= s_first[s + u]) s += u;
            return S.bins[s] + k2a_record_index (S, s, b, g - s_first[s], nb);
        };
        const uint64_t dbase = coarse_off[b];
        bool failed = false;
        uint32_t n_done = 0;                                                   // passes completed (CTA-uniform)
        for (uint32_t pass = 0; pass < P && !failed; pass++)
        {
            const uint32_t f_lo = (uint32_t)((uint64_t)nf * pass / P), f_hi = (uint32_t)((uint64_t)nf * (pass + 1) / P);
            for (uint32_t i = tid; i < ts; i += NT) tbl[i] = 0xFFFFFFFFu;
            for (uint32_t i = f_lo + tid; i < f_hi; i += NT) { s_tmp[i] = 0; s_cur[i] = 0; }
            if (tid == 0) { s_staged = 0; s_nb = 0; s_nb2 = 0; }
            __syncthreads ();
            // ---- stage the records of this pass: four independent 16-byte loads in flight per thread ----
            uint32_t n_st;
            if (P == 1)
            {
                for (uint32_t g0 = tid; g0 < n_all; g0 += 4 * NT)
                {
                    uint4 r[4];
                    #pragma unroll
                    for (int u = 0; u < 4; u++) { const uint32_t g = g0 + u * NT; if (g < n_all) r[u] = __ldg (src_of (g)); }
                    #pragma unroll
                    for (int u = 0; u < 4; u++) { const uint32_t g = g0 + u * NT; if (g < n_all) recs[g] = r[u]; }
                }
                __syncthreads ();
                n_st = n_all;
            }
            else
            {
                for (uint32_t g00 = 0; g00 < n_all; g00 += 4 * NT)            // warp-uniform bounds: the ballots below need whole warps
                {
                    const uint32_t g0 = g00 + tid;
                    uint4 r[4];
                    #pragma unroll
                    for (int u = 0; u < 4; u++) { const uint32_t g = g0 + u * NT; if (g < n_all) r[u] = __ldg (src_of (g)); }
                    #pragma unroll
                    for (int u = 0; u < 4; u++)
                    {
                        const uint32_t g = g0 + u * NT;
                        const uint32_t f = (g < n_all) ? ((r[u].w >> (DEV_FINE_SHIFT_W1 - 32)) & idmask) : 0xFFFFFFFFu;
                        const bool keep = f >= f_lo && f < f_hi;
                        const unsigned km = __ballot_sync (FULL_MASK, keep);            // one shared atomic per warp, not per record
                        if (km)
                        {
                            uint32_t at = 0;
                            if ((tid & 31) == 0) at = atomicAdd (&s_staged, (uint32_t)__popc (km));
                            at = __shfl_sync (FULL_MASK, at, 0) + __popc (km & ((1u << (tid & 31)) - 1));
                            if (keep && at < rmax) recs[at] = r[u];
                        }
                    }
                }
                __syncthreads ();
                n_st = s_staged;
                if (n_st > rmax) { failed = true; break; }                  // CTA-uniform
            }
            // ---- collapse identical records ----
            for (uint32_t g = tid; g < n_st; g += NT)
            {
                const uint4 r = recs[g];
                uint32_t h = (r.x * 0x9E3779B1u) ^ (r.y * 0x85EBCA77u) ^ (r.z * 0xC2B2AE3Du) ^ (r.w * 0x27D4EB2Fu);
                h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 13;
                h = __umulhi (h, ts);
                for (;;)
                {
                    uint32_t e = *(volatile uint32_t*)&tbl[h];
                    if (e == 0xFFFFFFFFu) { e = atomicCAS (&tbl[h], 0xFFFFFFFFu, g | (1u << 16)); if (e == 0xFFFFFFFFu) break; }
                    const uint4 o = recs[e & 0xFFFFu];
                    if (o.x == r.x && o.y == r.y && o.z == r.z && o.w == r.w) { atomicAdd (&tbl[h], 1u << 16); break; }
                    h = (h + 1 == ts) ? 0u : h + 1;
                }
            }
            __syncthreads ();
            // ---- surviving records and their k-mers per fine id ----
            for (uint32_t i = tid; i < ts; i += NT)
            {
                const uint32_t e = tbl[i];
                if (e != 0xFFFFFFFFu)
                {
                    const uint32_t w = recs[e & 0xFFFFu].w;
                    const uint32_t f = (w >> (DEV_FINE_SHIFT_W1 - 32)) & idmask;
                    // one atomic for both counters: records of the id in the low 13 bits (a pass stages at most 8191), their k-mers above
                    atomicAdd (&s_tmp[f], 1u | (((w >> (DEV_LEN_SHIFT_W1 - 32)) & 31u) << 13));
                }
            }
            __syncthreads ();
            // ---- exclusive scans (records, k-mers of the survivors) over the ids of this pass: every warp scans chunks of 32 ids ... ----
            const uint32_t nr = f_hi - f_lo, nch = (nr + 31) >> 5;                       // nch <= 128
            const uint32_t lane = tid & 31;
            for (uint32_t c = tid >> 5; c < nch; c += NT / 32)
            {
                const uint32_t idx = f_lo + 32 * c + lane;
                const uint32_t both = idx < f_hi ? s_tmp[idx] : 0u;
                const uint32_t v = both & 0x1FFFu, kv = both >> 13;
                uint32_t incl = v, kincl = kv;
                #pragma unroll
                for (int o = 1; o < 32; o <<= 1)
                {
                    const uint32_t y = __shfl_up_sync (FULL_MASK, incl, o), ky = __shfl_up_sync (FULL_MASK, kincl, o);
                    if (lane >= (uint32_t)o) { incl += y; kincl += ky; }
                }
                if (idx < f_hi) { s_off[idx] = incl - v; s_cur[idx] = kincl - kv; }
                if (lane == 31) { s_ct[c] = incl; s_ckt[c] = kincl; }
            }
            __syncthreads ();
            // ---- ... and every thread adds the totals of the chunks before its own (a handful: summed on the spot).  Record offsets
            //      continue after the previous passes; the k-mer prefix divided by 'target' is the id's bin quotient: ids with equal
            //      quotients form one bin, an id whose quotient differs from its predecessor's starts one (s_q: quotient | start flag).
            //      s_tmp becomes the write cursor of the id. ----
            const uint32_t written = s_written;
            uint32_t starts = 0, pass_total = 0, pass_kmers = 0;
            {
                uint32_t c_done = 0, roff = 0, koff = 0;                                     // totals of the chunks [0, c_done)
                for (uint32_t idx = f_lo + tid; idx < f_hi; idx += NT)
                {
                    const uint32_t c = (idx - f_lo) >> 5;
                    for (; c_done < c; c_done++) { roff += s_ct[c_done]; koff += s_ckt[c_done]; }
                    const uint32_t q = target ? (s_cur[idx] + koff) / target : idx;
                    bool start = (idx == f_lo);
                    if (!start)
                    {   // the predecessor's prefix: same chunk, or the last id of the previous chunk (lane 0; NT is a multiple of 32)
                        const uint32_t kp = s_cur[idx - 1] + (lane == 0 ? koff - s_ckt[c - 1] : koff);
                        start = (target ? kp / target : idx - 1) != q;
                    }
                    s_q[idx - f_lo] = ((s_cur[idx] + koff) << 1) | (start ? 1u : 0u);       // k-mers before the id | start flag
                    s_off[idx] += written + roff;
                    s_tmp[idx] = 0;
                    starts += start ? 1u : 0u;
                }
                for (; c_done < nch; c_done++) { roff += s_ct[c_done]; koff += s_ckt[c_done]; }
                pass_total = roff; pass_kmers = koff;
            }
            if (starts) atomicAdd (&s_nb, starts);
            __syncthreads ();
            const uint32_t end_all = written + pass_total;
            if (tid == 0)
            {   // (the global reservation travels while the CTA writes its survivors)
                s_gbase = atomicAdd (&counters[2], (unsigned long long)s_nb); s_pass_base[pass] = s_gbase; s_pass_nb[pass] = s_nb;
                s_written = end_all; n_unique += pass_total;
            }
            // ---- write the survivors ordered by fine id, multiplicity in place of the id ----
            for (uint32_t i = tid; i < ts; i += NT)
            {
                const uint32_t e = tbl[i];
                if (e == 0xFFFFFFFFu) continue;
                uint4 r = recs[e & 0xFFFFu];
                const uint32_t f = (r.w >> (DEV_FINE_SHIFT_W1 - 32)) & idmask;
                const uint32_t p = s_off[f] + atomicAdd (&s_tmp[f], 1u);
                r.w = (r.w & ((1u << (DEV_FINE_SHIFT_W1 - 32)) - 1)) | ((e >> 16) << (DEV_FINE_SHIFT_W1 - 32));
                dst[dbase + p] = r;
            }
            __syncthreads ();
            // ---- descriptors: a bin ends where the next one starts (a bin without records -- possible at the start of a pass -- is
            //      written with a zero count: the counting kernels skip it); slots from a shared counter, the order of the bins is free ----
            {
                const unsigned long long gbase = s_gbase;
                for (uint32_t idx = f_lo + tid; idx < f_hi; idx += NT)
                {
                    const uint32_t qs = s_q[idx - f_lo];
                    if (!(qs & 1u)) continue;
                    uint32_t e = idx + 1; while (e < f_hi && !(s_q[e - f_lo] & 1u)) e++;
                    const uint32_t first = s_off[idx], cnt = (e < f_hi ? s_off[e] : end_all) - first;
                    const uint32_t load = (e < f_hi ? (s_q[e - f_lo] >> 1) : pass_kmers) - (qs >> 1);       // k-mers of the bin's records
                    bin_desc[gbase + atomicAdd (&s_nb2, 1u)] = make_uint2 ((uint32_t)(dbase + first), cnt | ((big_load && load > big_load) ? K2_DESC_BIG : 0u));
                }
            }
            __syncthreads ();
            n_done = pass + 1;
        }
        if (failed)
        {   // a range overfilled the staging area (skewed fine ids): the two-pass kernel redoes the whole bin; the descriptors the
            // earlier passes appended are emptied (their records are about to be rewritten: they must not be counted twice)
            if (tid == 0) { const uint32_t idx = (uint32_t) atomicAdd (&counters[0], 1ULL); big_list[idx] = b; }
            // ('failed' was raised in pass index = number of completed passes; recompute it: passes done = those with a base recorded)
            for (uint32_t pp = 0; pp < n_done; pp++)
                for (uint32_t i = tid; i < s_pass_nb[pp]; i += NT) bin_desc[s_pass_base[pp] + i].y = 0;
        }
    }
    if (tid == 0 && n_unique) atomicAdd (&counters[1], n_unique);
}

// W: records of 16*W bytes.  bin_list == NULL: all nb1 coarse bins, descriptors indexed by (coarse bin, fine id); else the listed bins
// only (bins k2a_dedup_split gave up on), descriptors in its format appended at desc_base
cudaError_t launch_k2a_split (const LaunchCtx& L, int W, const K2aSrc& src, void* dst, const uint64_t* coarse_off,
                              uint32_t nb1, uint32_t cap, int fine_bits, uint2* bin_desc, const uint32_t* bin_list, uint32_t n_list,
                              int desc_abs, uint64_t desc_base, const K2aPresplit* presplit)
{
    K2aPresplit PS; PS.sub_off = 0; PS.sub_cnt = 0; PS.shift = 0;
    if (presplit) PS = *presplit;
    const uint32_t grid = bin_list ? n_list : nb1;
    if (grid == 0) return cudaSuccess;
    const size_t smem = (size_t)12 << fine_bits;
    if (smem > 40 * 1024)
    {
        cudaError_t e = cudaFuncSetAttribute (W == 1 ? (const void*)k2a_fine_split<1> : (const void*)k2a_fine_split<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    if (W == 1) k2a_fine_split<1><<<grid, 256, smem, L.stream>>> (src, (uint4*)dst, coarse_off, nb1, cap, fine_bits, bin_desc, bin_list, desc_abs, desc_base, PS);
    else        k2a_fine_split<2><<<grid, 256, smem, L.stream>>> (src, (uint4*)dst, coarse_off, nb1, cap, fine_bits, bin_desc, bin_list, desc_abs, desc_base, PS);
    (*L.launches)++;
    return cudaGetLastError ();
}

// k <= 31: dedup + split of all bins; counters[0] = bins listed in big_list (handed to the two-pass kernel), counters[1] = distinct
// records written, counters[2] = bin descriptors appended
// dedup table: 1.25 slots per staged record (not a power of two: the slot is umulhi (hash, ts)); about half of the staged records
// are duplicates, so the table runs at 35..45 % load, 80 % at the very worst
static uint32_t k2a_table_slots (uint32_t rmax) { return rmax + rmax / 4 + 32; }
static size_t k2a_dedup_smem (uint32_t rmax, int fine_bits) { return (size_t)rmax * 16 + (size_t)k2a_table_slots (rmax) * 4 + 4 * ((size_t)4 << fine_bits); }
// records a CTA can stage with two CTAs per SM
uint32_t k2a_two_cta_capacity (int fine_bits) { uint32_t r2 = 256; while (k2a_dedup_smem (r2 + 64, fine_bits) <= 111 * 1024) r2 += 64; return r2; }
uint32_t k2a_dedup_rmax (uint32_t max_bin_records, uint32_t mean_bin_records, int fine_bits)
{
    // Two CTAs of 512 threads per SM (111 KB each) when the typical bin fits with a quarter of head-room: the few larger bins take two
    // passes.  Otherwise one CTA of 1024 threads with up to 200 KB; at most 8191 records (multiplicities keep 15 bits, indices 16).
    uint32_t r2 = 256; while (k2a_dedup_smem (r2 + 64, fine_bits) <= 111 * 1024) r2 += 64;
    if (max_bin_records <= r2) return max_bin_records < 256 ? 256 : max_bin_records;
    if (mean_bin_records + mean_bin_records / 4 <= r2) return r2;
    uint32_t r1 = r2; while (r1 + 64 <= 8191 && k2a_dedup_smem (r1 + 64, fine_bits) <= 200 * 1024) r1 += 64;
    return max_bin_records < r1 ? max_bin_records : r1;
}
template<int NT>
static cudaError_t k2a_dedup_launch (const LaunchCtx& L, const K2aSrc& src, void* dst, const uint64_t* coarse_off, uint32_t nb1, uint32_t cap,
                                     int fine_bits, uint2* bin_desc, uint32_t rmax, uint32_t ts, size_t smem, uint32_t* big_list, unsigned long long* counters,
                                     uint32_t target, uint32_t big_load)
{
    cudaError_t e = cudaFuncSetAttribute (k2a_dedup_split<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor (&per_sm, k2a_dedup_split<NT>, NT, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    uint64_t grid = (uint64_t)L.sm_count * per_sm;
    if (grid > nb1) grid = nb1;
    k2a_dedup_split<NT><<<(unsigned)grid, NT, smem, L.stream>>> (src, (uint4*)dst, coarse_off, nb1, cap, fine_bits, bin_desc, rmax, ts, big_list, counters, target, big_load);
    (*L.launches)++;
    return cudaGetLastError ();
}
cudaError_t launch_k2a_dedup_split (const LaunchCtx& L, const K2aSrc& src, void* dst, const uint64_t* coarse_off, uint32_t nb1, uint32_t cap,
                                    int fine_bits, uint2* bin_desc, uint32_t rmax, uint32_t* big_list, unsigned long long* counters, uint32_t target, uint32_t big_load)
{
    if (nb1 == 0) return cudaSuccess;
    const uint32_t ts = k2a_table_slots (rmax);
    const size_t smem = k2a_dedup_smem (rmax, fine_bits);
    // a staging area so large that only one CTA fits an SM (bins gathered from several ranks): 1024 threads keep the SM busy
    if (smem > 113 * 1024) return k2a_dedup_launch<1024> (L, src, dst, coarse_off, nb1, cap, fine_bits, bin_desc, rmax, ts, smem, big_list, counters, target, big_load);
    return k2a_dedup_launch<512> (L, src, dst, coarse_off, nb1, cap, fine_bits, bin_desc, rmax, ts, smem, big_list, counters, target, big_load);
}

// ------------------------------------------------------------------------------------------------ record decoding
// k-mer j of a record -> canonical value.  W=1: record = (lo,hi) 128 bits, stream bits at [2j, 2j+2k)
__device__ __forceinline__ uint64_t rec_kmer_w1 (uint64_t lo, uint64_t hi, int j, int k)
{
    int s = 2*j;                                            // 0..54
    uint64_t x = s ? ((lo >> s) | (hi << (64 - s))) : lo;
    return canonical_from_stream64 (x & mask2k64 (k), k);
}
__device__ __forceinline__ u128 rec_kmer_w2 (const uint64_t* r, int j, int k)
{
    int s = 2*j, wi = s >> 6, sh = s & 63;                  // s in 0..118
    uint64_t a = r[wi], b = wi+1 < 4 ? r[wi+1] : 0, c = wi+2 < 4 ? r[wi+2] : 0;
    u128 x;
    x.lo = sh ? ((a >> sh) | (b << (64 - sh))) : a;
    x.hi = sh ? ((b >> sh) | (c << (64 - sh))) : b;
    x.hi &= mask2k64 (k - 32);
    return canonical_from_stream128 (x, k);
}

// multiply-shift slot hash: top 'log2' bits of key * odd constant
__device__ __forceinline__ uint32_t slot_hash64 (uint64_t key, int log2)
{ return (uint32_t)((key * 0x9E3779B97F4A7C15ULL) >> (64 - log2)); }
__device__ __forceinline__ uint32_t slot_hash128 (u128 key, int log2)
{ return slot_hash64 (key.lo ^ (key.hi * 0xC2B2AE3D27D4EB4FULL), log2); }

#define K2_THREADS 256
#define K2_ROUNDS  6          // table-scan rounds: occupied-slot list capacity (3T/4) <= K2_ROUNDS * K2_THREADS  => T <= 4096
#define K2_OUT_BLOCK 8192     // output slots a CTA reserves with one global atomic
template<int W> struct K2Cfg;
template<> struct K2Cfg<1> { enum { CH = 128, MAXLEN = 28, JBITS = 5 }; };   // records per staged chunk, k-mers per record
template<> struct K2Cfg<2> { enum { CH = 64,  MAXLEN = 60, JBITS = 6 }; };

// cheap slot hash for the shared-memory table: two 32-bit multiplies
__device__ __forceinline__ uint32_t smem_slot64 (uint64_t key, int log2)
{ uint32_t h = ((uint32_t)key * 0x9E3779B1u) ^ ((uint32_t)(key >> 32) * 0x85EBCA77u); h ^= h >> 15; return (h * 0x2C1B3C6Du) >> (32 - log2); }

// ---- insertion into an open-addressed table in GLOBAL memory (fallback path) -----------------------------------
// W=1: the key word itself is claimed with a 64-bit CAS.
__device__ __forceinline__ bool table_insert_w1 (unsigned long long* keys, uint32_t* cnts, int log2, uint64_t key, int maxprobe, uint32_t add = 1u)
{
    const uint32_t tmask = (1u << log2) - 1;
    uint32_t slot = slot_hash64 (key, log2);
    for (int probe = 0; probe < maxprobe; probe++)
    {
        unsigned long long cur = keys[slot];
        if (cur == EMPTY64) cur = atomicCAS (&keys[slot], EMPTY64, (unsigned long long)key);
        if (cur == EMPTY64 || cur == key) { atomicAdd (&cnts[slot], add); return true; }
        slot = (slot + 1) & tmask;
    }
    return false;
}
// W=2: the high word is claimed with a CAS from EMPTY to a LOCK value that is never a valid high half (k <= 63 keeps
// bit 63 clear); the low word is then published and the high word released.  Readers spin while they see LOCK.
#define LOCK64 0xFFFFFFFFFFFFFFFEULL
template<bool GLOBAL>
__device__ __forceinline__ int table_insert_w2 (unsigned long long* klo, unsigned long long* khi, uint32_t* cnts, int log2, uint32_t slot, u128 key, int maxprobe)
{
    const uint32_t tmask = (1u << log2) - 1;
    for (int probe = 0; probe < maxprobe; probe++)
    {
        for (;;)
        {
            unsigned long long h = *(volatile unsigned long long*)&khi[slot];
            if (h == EMPTY64)
            {
                h = atomicCAS (&khi[slot], EMPTY64, LOCK64);
                if (h == EMPTY64)
                {   // we own the slot
                    *(volatile unsigned long long*)&klo[slot] = key.lo;
                    if (GLOBAL) __threadfence (); else __threadfence_block ();
                    *(volatile unsigned long long*)&khi[slot] = key.hi;
                    atomicAdd (&cnts[slot], 1u);
                    return (int)slot | 0x40000000;           // new key
                }
            }
            if (h == LOCK64) continue;                       // another thread is publishing this slot
            if (GLOBAL) __threadfence ();                    // order the high-word read before the low-word read
            if (h == key.hi && *(volatile unsigned long long*)&klo[slot] == key.lo) { atomicAdd (&cnts[slot], 1u); return (int)slot; }
            break;                                           // occupied by another key
        }
        slot = (slot + 1) & tmask;
    }
    return -1;
}

// ---- consuming one table slot of the GLOBAL fallback table: histogram + statistics + emission -------------------
struct EmitState { unsigned long long distinct, solid, emitted; };

__device__ __forceinline__ void consume_entry (const K2Params& P, bool occupied, uint64_t klo, uint64_t khi, uint32_t c, EmitState& st, uint32_t* s_hist)
{
    bool emit = false;
    if (occupied)
    {
        st.distinct++;
        uint32_t hb = c >= (uint32_t)P.histo_max ? (uint32_t)P.histo_max : c;
        if (hb < K2_HB) atomicAdd (&s_hist[hb], 1u); else atomicAdd (&P.histogram[hb], 1ULL);
        if (c >= P.solid_min && c <= P.solid_max) st.solid++;
        emit = (c >= P.emit_min && c <= P.emit_max);
    }
    unsigned ballot = __ballot_sync (__activemask (), emit);
    if (emit)
    {
        st.emitted++;
        int leader = __ffs (ballot) - 1, lane = threadIdx.x & 31;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd (&P.counters[0], (unsigned long long)__popc (ballot));
        base = __shfl_sync (ballot, base, leader);
        unsigned long long pos = base + __popc (ballot & ((1u << lane) - 1));
        if (pos < P.out_cap) { P.out_lo[pos] = klo; if (P.out_hi) P.out_hi[pos] = khi; P.out_cnt[pos] = c; }
    }
}

// ------------------------------------------------------------------------------------------------ k2b
// Persistent CTAs; per fine bin:
//   * the bin's records arrive in shared memory by TMA bulk copy into one of two staging buffers; thread 0 fetches the
//     NEXT bin from the work counter and starts its copy before the current bin is processed (latency hidden);
//   * insert phase, warp-autonomous (no block barrier): a warp takes 32 records (lane <-> record), prefix-sums their
//     k-mer counts with shuffles and then walks the k-mers 32 at a time (lane <-> k-mer).  The k-mer -> record map of a
//     32-wide window is one __reduce_or_sync of "head" bits plus a popc; the record comes back from shared memory with
//     one 16-byte load and the k-mer is rebuilt with bit tricks (common.cuh).  Insert = LDS.64 probe, 64-bit atomicCAS
//     only on an empty slot, shared atomicAdd on the count; a slot claimed for the first time is appended to the bin's
//     claimed-slot list;
//   * scan phase: the claimed-slot list (not the T slots) is walked: histogram in shared memory, statistics, slots
//     cleared, and k-mers in the emission range appended to a per-WARP output block (one global atomic per 2048
//     emitted k-mers; unused tails are marked EMPTY and skipped by k3a).
// Block barriers per bin: one after the inserts, one after the scan.
// Shared-memory layout (T = table slots): staging 2 x CHR records | keys T*8W | counts T*4 | claimed 3T/4 u16 | hist
template<int W>
__global__ void __launch_bounds__(K2_THREADS) k2b_bucket_hash_count (const K2Params P)
{
    constexpr int CHR = (W == 1) ? 256 : 128;              // records per staging buffer (4 KB)
    constexpr int GRP = 16;                                // records a warp expands at a time (lanes 0..GRP-1 hold one each)
    constexpr int NWARP = K2_THREADS / 32;
    constexpr unsigned WBLOCK = 2048;                      // output slots a warp reserves at a time
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int T = 1 << P.table_log2;
    const int OCC_CAP = (T * 3) / 4;
    uint4* s_recs = (uint4*)smem_raw;                                        // [2][CHR*W]
    unsigned long long* s_klo = (unsigned long long*)(smem_raw + (size_t)2 * CHR * 16 * W);
    unsigned long long* s_khi = (W == 2) ? s_klo + T : 0;
    uint32_t* s_cnt  = (uint32_t*)(s_klo + (size_t)T * W);
    uint16_t* s_occ  = (uint16_t*)(s_cnt + T);
    uint32_t* s_hist = (uint32_t*)(s_occ + OCC_CAP + (OCC_CAP & 1));
    uint64_t* s_bar  = (uint64_t*)(s_hist + K2_HB + (K2_HB & 1));             // [2]
    __shared__ uint32_t s_bin[2], s_n[2], s_nocc[2];
    __shared__ unsigned long long s_base[2];
    __shared__ int s_ovf;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int k = P.k;
    for (int i = tid; i < T; i += K2_THREADS) { if (W == 1) s_klo[i] = EMPTY64; else { s_khi[i] = EMPTY64; s_klo[i] = 0; } s_cnt[i] = 0; }
    for (int i = tid; i < K2_HB; i += K2_THREADS) s_hist[i] = 0;

    // thread 0 only: take the next bin from the work counter and start the TMA copy of its first chunk
    auto fetch = [&] (int st)
    {
        uint32_t bin; uint2 d = make_uint2 (0, 0);
        for (;;)
        {   // skip empty bins here so that the CTA never synchronises for nothing
            bin = (uint32_t) atomicAdd (&P.counters[3], 1ULL);
            if (P.bin_list) bin = bin < P.n_list ? P.bin_list[bin] : 0xFFFFFFFFu;       // tier run: the bins an earlier table could not hold
            if (bin >= P.nbins) break;
            d = P.bin_desc[bin]; d.y &= K2_DESC_COUNT;
            if (d.y) break;
        }
        s_bin[st] = bin; s_nocc[st] = 0;
        if (bin < P.nbins)
        {
            const unsigned long long base = P.coarse_off[bin >> P.fine_bits] + d.x;
            s_n[st] = d.y; s_base[st] = base;
            const uint32_t mrec = min ((uint32_t)CHR, d.y);
            fence_proxy_async ();
            mbar_expect_tx (&s_bar[st], mrec * 16 * W);
            tma_bulk_g2s (s_recs + (size_t)st * CHR * W, (const uint4*)P.recs + base * W, mrec * 16 * W, &s_bar[st]);
        }
    };
    if (tid == 0)
    {
        mbar_init (&s_bar[0], 1); mbar_init (&s_bar[1], 1);
        asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_ovf = 0;
        fetch (0);
    }
    __syncthreads ();
    uint32_t par0 = 0, par1 = 0;
    unsigned long long n_distinct = 0, n_solid = 0, n_emitted = 0;
    unsigned long long out_pos = 0, out_end = 0;           // this warp's output block (warp-uniform)
    int st = 0;

    for (;;)
    {
        const uint32_t bin = s_bin[st];
        if (bin >= P.nbins) break;
        const uint32_t n = s_n[st];
        const unsigned long long base = s_base[st];
        if (tid == 0) fetch (st ^ 1);                       // prefetch the next bin into the other staging buffer
        const uint4* recs = s_recs + (size_t)st * CHR * W;

        for (uint32_t c0 = 0; c0 < n; c0 += CHR)
        {
            const uint32_t mrec = min ((uint32_t)CHR, n - c0);
            if (c0)
            {   // bins larger than one staging buffer (rare): reload the same buffer
                __syncthreads ();
                if (tid == 0)
                {
                    fence_proxy_async ();
                    mbar_expect_tx (&s_bar[st], mrec * 16 * W);
                    tma_bulk_g2s ((void*)recs, (const uint4*)P.recs + (base + c0) * W, mrec * 16 * W, &s_bar[st]);
                }
            }
            if (st == 0) { mbar_wait (&s_bar[0], par0); par0 ^= 1; } else { mbar_wait (&s_bar[1], par1); par1 ^= 1; }

            // ---- insert phase: warp-autonomous ----
            for (uint32_t g0 = wid * GRP; g0 < mrec; g0 += NWARP * GRP)
            {
                const uint32_t ri = g0 + lane;
                int len = 0;
                if (lane < GRP && ri < mrec)
                {
                    const uint32_t top = recs[(size_t)ri * W + (W - 1)].w;
                    len = (W == 1) ? (int)((top >> (DEV_LEN_SHIFT_W1 - 32)) & 31) : (int)((top >> (REC_LEN_SHIFT_W2 - 32)) & 63);
                }
                uint32_t incl = (uint32_t)len;
                #pragma unroll
                for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync (FULL_MASK, incl, o); if (lane >= o) incl += y; }
                const uint32_t excl = incl - (uint32_t)len;
                const uint32_t total = __shfl_sync (FULL_MASK, incl, 31);
                uint32_t r0 = 0;                                            // records that start before the window
                for (uint32_t wb = 0; wb < total; wb += 32)
                {
                    const uint32_t h = excl - wb;                           // head position inside the window (unsigned wrap = outside)
                    const uint32_t M = __reduce_or_sync (FULL_MASK, (len > 0 && h < 32u) ? (1u << h) : 0u);
                    const uint32_t gk = wb + lane;
                    uint32_t r = r0 + __popc (M & (0xFFFFFFFFu >> (31 - lane))) - 1;
                    r0 += __popc (M);
                    const uint32_t ex = __shfl_sync (FULL_MASK, excl, r & 31);
                    if (gk < total)
                    {
                        const int j = (int)(gk - ex);
                        const uint32_t rr = g0 + r;
                        if (W == 1)
                        {
                            const uint4 q = recs[rr];
                            const uint64_t lo = (uint64_t)q.x | ((uint64_t)q.y << 32);
                            const uint64_t hi = ((uint64_t)q.z | ((uint64_t)q.w << 32)) & ((1ULL << DEV_LEN_SHIFT_W1) - 1);
                            const uint64_t key = rec_kmer_w1 (lo, hi, j, k);
                            const uint32_t mult = q.w >> (DEV_FINE_SHIFT_W1 - 32);          // multiplicity of the record (k2a)
                            uint32_t slot = smem_slot64 (key, P.table_log2);
                            int probe = 0;
                            for (;;)
                            {
                                unsigned long long cur = s_klo[slot];
                                if (cur == key) { atomicAdd (&s_cnt[slot], mult); break; }
                                if (cur == EMPTY64)
                                {
                                    cur = atomicCAS (&s_klo[slot], EMPTY64, (unsigned long long)key);
                                    if (cur == EMPTY64)
                                    {   // first occurrence in this bin: remember the slot
                                        const uint32_t qn = atomicAdd (&s_nocc[st], 1u);
                                        if (qn < (uint32_t)OCC_CAP) s_occ[qn] = (uint16_t)slot; else s_ovf = 1;
                                        atomicAdd (&s_cnt[slot], mult); break;
                                    }
                                    if (cur == key) { atomicAdd (&s_cnt[slot], mult); break; }
                                }
                                slot = (slot + 1) & (T - 1);
                                if (++probe >= K2_MAXPROBE) { s_ovf = 1; break; }
                            }
                        }
                        else
                        {
                            const uint4 a = recs[2*rr], b = recs[2*rr+1];
                            uint64_t rw[4] = { (uint64_t)a.x | ((uint64_t)a.y << 32), (uint64_t)a.z | ((uint64_t)a.w << 32),
                                               (uint64_t)b.x | ((uint64_t)b.y << 32), ((uint64_t)b.z | ((uint64_t)b.w << 32)) & ((1ULL << REC_LEN_SHIFT_W2) - 1) };
                            const u128 key = rec_kmer_w2 (rw, j, k);
                            const uint32_t slot = smem_slot64 (key.lo ^ (key.hi * 0xC2B2AE3D27D4EB4FULL), P.table_log2);
                            const int res = table_insert_w2<false> (s_klo, s_khi, s_cnt, P.table_log2, slot, key, K2_MAXPROBE);
                            if (res < 0) s_ovf = 1;
                            else if (res & 0x40000000)
                            {
                                const uint32_t qn = atomicAdd (&s_nocc[st], 1u);
                                if (qn < (uint32_t)OCC_CAP) s_occ[qn] = (uint16_t)(res & 0xFFFF); else s_ovf = 1;
                            }
                        }
                    }
                }
            }
        }
        __syncthreads ();                                     // all inserts of the bin are done
        const bool ovf = s_ovf != 0;
        if (ovf)
        {   // the bin goes to the global-memory fallback (k2c): wipe the whole table
            if (tid == 0) { uint32_t idx = (uint32_t) atomicAdd (&P.counters[P.ovf_counter], 1ULL); P.ovf_list[idx] = bin; }
            for (int i = tid; i < T; i += K2_THREADS) { if (W == 1) s_klo[i] = EMPTY64; else s_khi[i] = EMPTY64; s_cnt[i] = 0; }
            __syncthreads ();
            if (tid == 0) s_ovf = 0;
            __syncthreads ();
            st ^= 1;
            continue;
        }
        // ---- scan phase over the claimed slots: warp-autonomous ----
        const uint32_t nocc = s_nocc[st];
        for (uint32_t q0 = wid * 32; q0 < nocc; q0 += NWARP * 32)
        {
            const uint32_t q = q0 + lane;
            bool emit = false; uint64_t klo = 0, khi = 0; uint32_t c = 0;
            if (q < nocc)
            {
                const uint32_t slot = s_occ[q];
                c = s_cnt[slot]; klo = s_klo[slot]; if (W == 2) khi = s_khi[slot];
                if (W == 1) s_klo[slot] = EMPTY64; else s_khi[slot] = EMPTY64;
                s_cnt[slot] = 0;
                n_distinct++;
                const uint32_t hb = c >= (uint32_t)P.histo_max ? (uint32_t)P.histo_max : c;
                if (hb < K2_HB) atomicAdd (&s_hist[hb], 1u); else atomicAdd (&P.histogram[hb], 1ULL);
                if (c >= P.solid_min && c <= P.solid_max) n_solid++;
                emit = (c >= P.emit_min && c <= P.emit_max);
            }
            const unsigned ballot = __ballot_sync (FULL_MASK, emit);
            if (ballot)
            {
                const unsigned ne = __popc (ballot);
                if (out_pos + ne > out_end)
                {   // rest of the block -> holes; reserve a new block
                    for (unsigned long long hpos = out_pos + lane; hpos < out_end; hpos += 32)
                        if (hpos < P.out_cap) { if (W == 1) P.out_lo[hpos] = EMPTY64; else P.out_hi[hpos] = EMPTY64; }
                    unsigned long long b0 = 0;
                    if (lane == 0) b0 = atomicAdd (&P.counters[0], (unsigned long long)WBLOCK);
                    b0 = __shfl_sync (FULL_MASK, b0, 0);
                    out_pos = b0; out_end = b0 + WBLOCK;
                }
                if (emit)
                {
                    const unsigned long long pos = out_pos + __popc (ballot & ((1u << lane) - 1));
                    n_emitted++;
                    if (pos < P.out_cap) { P.out_lo[pos] = klo; if (W == 2) P.out_hi[pos] = khi; P.out_cnt[pos] = c; }
                }
                out_pos += ne;
            }
        }
        __syncthreads ();                                     // table clean again before the next bin's inserts
        st ^= 1;
    }
    // ---- holes at the end of each warp's last block, shared histogram, statistics ----
    for (unsigned long long hpos = out_pos + lane; hpos < out_end; hpos += 32)
        if (hpos < P.out_cap) { if (W == 1) P.out_lo[hpos] = EMPTY64; else P.out_hi[hpos] = EMPTY64; }
    __syncthreads ();
    for (int i = tid; i < K2_HB; i += K2_THREADS) { uint32_t v = s_hist[i]; if (v) atomicAdd (&P.histogram[i], (unsigned long long)v); }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        n_distinct += __shfl_xor_sync (FULL_MASK, n_distinct, o); n_solid += __shfl_xor_sync (FULL_MASK, n_solid, o);
        n_emitted += __shfl_xor_sync (FULL_MASK, n_emitted, o);
    }
    if (lane == 0)
    {
        if (n_distinct) atomicAdd (&P.counters[1], n_distinct);
        if (n_solid)    atomicAdd (&P.counters[2], n_solid);
        if (n_emitted)  atomicAdd (&P.counters[6], n_emitted);
    }
}


// ------------------------------------------------------------------------------------------------ k2b, k <= 31
// Same bin pipeline as k2b_bucket_hash_count (persistent CTAs, TMA-staged bins, per-warp output blocks) with an insert
// phase built for SIMT efficiency:
//   * lane <-> CHUNK of four consecutive k-mers of one record.  A warp takes 32 records (lane <-> record), prefix-sums
//     their chunk counts and walks the chunks 32 at a time; the chunk -> record map of a window is one __reduce_or_sync
//     of "head" bits plus a popc.  The four k-mers of a chunk share one record load and one pair-reversal of their
//     common 68-bit window (k2_decode.cuh); every shift inside the chunk is an immediate;
//   * the whole warp stays converged through the four insert steps: one probe (LDS.64 issued for all four k-mers up
//     front), a predicated 64-bit CAS on an empty slot, a predicated count increment.  No data-dependent loop;
//   * first occurrences are appended to the WARP's claimed-slot list with a ballot + popc (the cursor is a warp-uniform
//     register: no atomics); k-mers whose first slot holds another key go to the warp's retry list and are re-inserted
//     32 at a time with the general probing loop, so the rare long probe never stalls 31 other lanes.
// After the block barrier each warp walks its own claimed-slot list: histogram, statistics, emission, slot cleanup.
#define K2_RETRY_CAP 160        // per warp: 31 left over + 4 steps x 32 lanes at the very worst

template<int NT>
__global__ void __launch_bounds__(NT, 768 / NT) k2b_count_w1 (const K2Params P)
{
    constexpr int CHR = 256;                               // records per staging buffer (4 KB)
    constexpr int NWARP = NT / 32;
    constexpr unsigned WBLOCK = 2048;                      // output slots a warp reserves at a time
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int T = 1 << P.table_log2;
    const uint32_t tmask = (uint32_t)T - 1;
    const int hshift = 32 - P.table_log2;
    const uint32_t OCC_W = (uint32_t)((T * 3) / 4) / NWARP;                 // claimed-slot capacity per warp
    uint4* s_recs = (uint4*)smem_raw;                                        // [2][CHR]
    unsigned long long* s_klo = (unsigned long long*)(smem_raw + (size_t)2 * CHR * 16);
    unsigned long long* s_retry = s_klo + T;                                 // [NWARP][K2_RETRY_CAP]
    uint32_t* s_cnt  = (uint32_t*)(s_retry + NWARP * K2_RETRY_CAP);
    uint32_t* s_hist = s_cnt + T;
    uint64_t* s_bar  = (uint64_t*)(s_hist + K2_HB);                          // [2]
    uint16_t* s_occ  = (uint16_t*)(s_bar + 2);                               // [NWARP][OCC_W]
    __shared__ uint16_t s_retry_m[NWARP * K2_RETRY_CAP];                     // multiplicities of the retry entries
    __shared__ uint32_t s_bin[2], s_n[2];
    __shared__ unsigned long long s_base[2];
    __shared__ int s_ovf;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1;
    const int k = P.k;
    for (int i = tid; i < T; i += NT) { s_klo[i] = EMPTY64; s_cnt[i] = 0; }
    for (int i = tid; i < K2_HB; i += NT) s_hist[i] = 0;
    uint16_t* occ_w = s_occ + (size_t)wid * OCC_W;
    unsigned long long* retry_w = s_retry + wid * K2_RETRY_CAP;
    uint16_t* retry_m = s_retry_m + wid * K2_RETRY_CAP;

    auto fetch = [&] (int st)
    {
        uint32_t bin; uint2 d = make_uint2 (0, 0);
        for (;;)
        {
            const uint32_t idx = (uint32_t) atomicAdd (&P.counters[3], 1ULL);
            if (P.bin_list) { if (idx >= P.n_list) { bin = P.nbins; break; } bin = P.bin_list[idx]; }
            else            { bin = idx; if (bin >= P.nbins) break; }
            d = P.bin_desc[bin]; d.y &= K2_DESC_COUNT;
            if (d.y) break;
        }
        s_bin[st] = bin;
        if (bin < P.nbins)
        {
            const unsigned long long base = P.coarse_off[bin >> P.fine_bits] + d.x;
            s_n[st] = d.y; s_base[st] = base;
            const uint32_t mrec = min ((uint32_t)CHR, d.y);
            fence_proxy_async ();
            mbar_expect_tx (&s_bar[st], mrec * 16);
            tma_bulk_g2s (s_recs + (size_t)st * CHR, (const uint4*)P.recs + base, mrec * 16, &s_bar[st]);
        }
    };
    if (tid == 0)
    {
        mbar_init (&s_bar[0], 1); mbar_init (&s_bar[1], 1);
        asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_ovf = 0;
        fetch (0);
    }
    __syncthreads ();
    uint32_t par0 = 0, par1 = 0;
    unsigned long long n_distinct = 0, n_solid = 0, n_emitted = 0;
    unsigned long long out_pos = 0, out_end = 0;
    int st = 0;

    for (;;)
    {
        const uint32_t bin = s_bin[st];
        if (bin >= P.nbins) break;
        const uint32_t n = s_n[st];
        const unsigned long long base = s_base[st];
        if (tid == 0) fetch (st ^ 1);
        const uint4* recs = s_recs + (size_t)st * CHR;
        uint32_t wn = 0, rn = 0;                              // warp-uniform: claimed slots, pending retries
        bool w_ovf = false;

        // appends the slots claimed by this step (ballot order) to the warp's list
        auto append_new = [&] (bool isnew, uint32_t slot)
        {
            const unsigned m = __ballot_sync (FULL_MASK, isnew);
            if (m)
            {
                const uint32_t idx = wn + __popc (m & lt_mask);
                if (isnew && idx < OCC_W) occ_w[idx] = (uint16_t)slot;
                wn += __popc (m);
            }
        };
        // re-inserts the pending retries (general probing loop), 32 at a time
        auto drain_retries = [&] ()
        {
            __syncwarp ();
            for (uint32_t e0 = 0; e0 < rn; e0 += 32)
            {
                const uint32_t e = e0 + lane;
                int res = 0;
                if (e < rn)
                {
                    const unsigned long long key = retry_w[e];
                    res = k2_probe_loop (s_klo, s_cnt, k2_slot32 ((uint32_t)key, (uint32_t)(key >> 32), hshift), key, tmask, (uint32_t)retry_m[e]);
                    if (res == -1) w_ovf = true;
                }
                __syncwarp ();
                append_new (e < rn && res != -1 && (res & 0x80000000), (uint32_t)res & 0xFFFFu);
            }
            rn = 0;
            __syncwarp ();
        };

        for (uint32_t c0 = 0; c0 < n; c0 += CHR)
        {
            const uint32_t mrec = min ((uint32_t)CHR, n - c0);
            if (c0)
            {
                __syncthreads ();
                if (tid == 0)
                {
                    fence_proxy_async ();
                    mbar_expect_tx (&s_bar[st], mrec * 16);
                    tma_bulk_g2s ((void*)recs, (const uint4*)P.recs + (base + c0), mrec * 16, &s_bar[st]);
                }
            }
            if (st == 0) { mbar_wait (&s_bar[0], par0); par0 ^= 1; } else { mbar_wait (&s_bar[1], par1); par1 ^= 1; }

            for (uint32_t g0 = wid * 32; g0 < mrec; g0 += NWARP * 32)
            {
                const uint32_t ri = g0 + lane;
                uint32_t nch = 0;
                if (ri < mrec) nch = (((recs[ri].w >> (DEV_LEN_SHIFT_W1 - 32)) & 31u) + 3u) >> 2;
                uint32_t incl = nch;
                #pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync (FULL_MASK, incl, o); if (lane >= o) incl += y; }
                const uint32_t excl = incl - nch;
                const uint32_t total = __shfl_sync (FULL_MASK, incl, 31);
                uint32_t r0 = 0;
                for (uint32_t wb = 0; wb < total; wb += 32)
                {
                    const uint32_t h = excl - wb;
                    const uint32_t M = __reduce_or_sync (FULL_MASK, (nch > 0 && h < 32u) ? (1u << h) : 0u);
                    const uint32_t gk = wb + lane;
                    const bool act = gk < total;
                    uint32_t r = r0 + __popc (M & (0xFFFFFFFFu >> (31 - lane))) - 1;
                    r0 += __popc (M);
                    r &= 31u;
                    const uint32_t ex = __shfl_sync (FULL_MASK, excl, r);
                    const int c = act ? (int)(gk - ex) : 0;
                    const uint4 q = recs[act ? g0 + r : g0];
                    const int nkc = act ? (int)((q.w >> (DEV_LEN_SHIFT_W1 - 32)) & 31u) - 4 * c : 0;      // k-mers of this chunk
                    K2Chunk C;
                    k2_chunk_begin (C, q.x, q.y, q.z, q.w & ((1u << (DEV_LEN_SHIFT_W1 - 32)) - 1), c, k);
                    const uint32_t mult = q.w >> (DEV_FINE_SHIFT_W1 - 32);                  // multiplicity of the record (k2a)
                    uint32_t lo[4], hi[4], slot[4];
                    k2_chunk_kmer<0> (C, lo[0], hi[0]); k2_chunk_kmer<1> (C, lo[1], hi[1]);
                    k2_chunk_kmer<2> (C, lo[2], hi[2]); k2_chunk_kmer<3> (C, lo[3], hi[3]);
                    unsigned long long cur[4];
                    #pragma unroll
                    for (int i = 0; i < 4; i++) { slot[i] = k2_slot32 (lo[i], hi[i], hshift); cur[i] = (i < nkc) ? s_klo[slot[i]] : 0ULL; }
                    #pragma unroll
                    for (int i = 0; i < 4; i++)
                    {
                        const bool valid = i < nkc;
                        const unsigned long long key = ((unsigned long long)hi[i] << 32) | lo[i];
                        unsigned long long cv = cur[i];
                        const bool isE = valid && cv == EMPTY64;
                        if (isE) cv = atomicCAS (&s_klo[slot[i]], EMPTY64, key);
                        const bool isnew = isE && cv == EMPTY64;
                        const bool hit = valid && (isnew || cv == key);
                        if (hit && mult > (isnew ? 1u : 0u)) atomicAdd (&s_cnt[slot[i]], isnew ? mult - 1u : mult);     // a claim counts one by itself
                        append_new (isnew, slot[i]);
                        const bool miss = valid && !hit;
                        const unsigned mm = __ballot_sync (FULL_MASK, miss);
                        if (mm)
                        {
                            if (miss) { retry_w[rn + __popc (mm & lt_mask)] = key; retry_m[rn + __popc (mm & lt_mask)] = (uint16_t)mult; }
                            rn += __popc (mm);
                        }
                    }
                    if (rn >= 32) drain_retries ();
                }
            }
        }
        if (rn) drain_retries ();
        if (w_ovf || wn > OCC_W) s_ovf = 1;
        __syncthreads ();                                     // all inserts of the bin are done
        const bool ovf = s_ovf != 0;
        if (ovf)
        {
            if (tid == 0) { uint32_t idx = (uint32_t) atomicAdd (&P.counters[P.ovf_counter], 1ULL); P.ovf_list[idx] = bin; }
            for (int i = tid; i < T; i += NT) { s_klo[i] = EMPTY64; s_cnt[i] = 0; }
            __syncthreads ();
            if (tid == 0) s_ovf = 0;
            __syncthreads ();
            st ^= 1;
            continue;
        }
        for (uint32_t q0 = 0; q0 < wn; q0 += 32)
        {
            const uint32_t q = q0 + lane;
            bool emit = false; uint64_t klo = 0; uint32_t c = 0;
            if (q < wn)
            {
                const uint32_t slot = occ_w[q];
                c = s_cnt[slot] + 1u; klo = s_klo[slot];                    // the claim itself counts one (k2_common.cuh)
                s_klo[slot] = EMPTY64; s_cnt[slot] = 0;
                n_distinct++;
                const uint32_t hb = c >= (uint32_t)P.histo_max ? (uint32_t)P.histo_max : c;
                if (hb < K2_HB) atomicAdd (&s_hist[hb], 1u); else atomicAdd (&P.histogram[hb], 1ULL);
                if (c >= P.solid_min && c <= P.solid_max) n_solid++;
                emit = (c >= P.emit_min && c <= P.emit_max);
            }
            const unsigned ballot = __ballot_sync (FULL_MASK, emit);
            if (ballot)
            {
                const unsigned ne = __popc (ballot);
                if (out_pos + ne > out_end)
                {
                    for (unsigned long long hpos = out_pos + lane; hpos < out_end; hpos += 32)
                        if (hpos < P.out_cap) P.out_lo[hpos] = EMPTY64;
                    unsigned long long b0 = 0;
                    if (lane == 0) b0 = atomicAdd (&P.counters[0], (unsigned long long)WBLOCK);
                    b0 = __shfl_sync (FULL_MASK, b0, 0);
                    out_pos = b0; out_end = b0 + WBLOCK;
                }
                if (emit)
                {
                    const unsigned long long pos = out_pos + __popc (ballot & lt_mask);
                    n_emitted++;
                    if (pos < P.out_cap) { P.out_lo[pos] = klo; P.out_cnt[pos] = c; }
                }
                out_pos += ne;
            }
        }
        __syncthreads ();                                     // table clean again before the next bin's inserts
        st ^= 1;
    }
    for (unsigned long long hpos = out_pos + lane; hpos < out_end; hpos += 32)
        if (hpos < P.out_cap) P.out_lo[hpos] = EMPTY64;
    __syncthreads ();
    for (int i = tid; i < K2_HB; i += NT) { uint32_t v = s_hist[i]; if (v) atomicAdd (&P.histogram[i], (unsigned long long)v); }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        n_distinct += __shfl_xor_sync (FULL_MASK, n_distinct, o); n_solid += __shfl_xor_sync (FULL_MASK, n_solid, o);
        n_emitted += __shfl_xor_sync (FULL_MASK, n_emitted, o);
    }
    if (lane == 0)
    {
        if (n_distinct) atomicAdd (&P.counters[1], n_distinct);
        if (n_solid)    atomicAdd (&P.counters[2], n_solid);
        if (n_emitted)  atomicAdd (&P.counters[6], n_emitted);
    }
}
static size_t k2b_w1_smem_bytes (int table_log2, int nt)
{
    const size_t T = (size_t)1 << table_log2, nwarp = nt / 32;
    size_t occ_w = ((T * 3) / 4) / nwarp;
    return 2 * 256 * 16 + T * 8 + nwarp * K2_RETRY_CAP * 8 + T * 4 + K2_HB * 4 + 16 + nwarp * occ_w * 2 + 16;
}
template<int NT>
static cudaError_t k2b_w1_launch (const LaunchCtx& L, const K2Params& P)
{
    const size_t smem = k2b_w1_smem_bytes (P.table_log2, NT);
    cudaError_t e = cudaFuncSetAttribute (k2b_count_w1<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor (&per_sm, k2b_count_w1<NT>, NT, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    uint64_t grid = (uint64_t)L.sm_count * per_sm;
    if (grid > P.nbins) grid = P.nbins;
    k2b_count_w1<NT><<<(unsigned)grid, NT, smem, L.stream>>> (P);
    (*L.launches)++;
    return cudaGetLastError ();
}

// ------------------------------------------------------------------------------------------------ k2b, k <= 31, warp per bin
// One WARP owns a fine bin from its first record to its last emitted k-mer: private shared-memory table (T slots),
// private claimed-slot and retry lists, no block barrier and no work counter (warp g takes bins g, g+G, g+2G, ...).
// The records of a bin go straight from global memory into registers (lane <-> record, one coalesced 16-byte load per
// lane; the chunk -> record gather inside the warp is four shuffles), and the descriptor and first records of the NEXT
// bin are requested before the current bin is processed, so their latency hides behind the inserts.
// Insert and scan phases are those of k2b_count_w1 (chunks of four k-mers, converged probe/claim/count steps).
// ORI: the records are oriented (k1_scan.cuh): keys are plain slices of the record, identical records of a batch are
// collapsed first (lane <-> record, four MATCH.ANY: reads covering the same locus without an error in the span produce
// the same record whatever their strand) and inserted once with their multiplicity; the canonical VALUE is rebuilt
// only for the k-mers that are emitted.
template<int NT, bool ORI>
__global__ void __launch_bounds__(NT, 768 / NT) k2b_warp_bins (const K2Params P)
{
    constexpr int NWARP = NT / 32;
    constexpr unsigned WBLOCK = 2048;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int T = 1 << P.table_log2;
    const uint32_t tmask = (uint32_t)T - 1;
    const int hshift = 32 - P.table_log2;
    const uint32_t OCC_W = (uint32_t)(T * 3) / 4;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1;
    const int k = P.k;
    // per warp: keys T*8 | retry K2_RETRY_CAP*8 | counts T*4 | claimed OCC_W*2 | retry multiplicities ; per CTA: histogram
    const size_t per_warp = (size_t)T * 8 + K2_RETRY_CAP * 8 + (size_t)T * 4 + (((size_t)OCC_W * 2 + 15) & ~(size_t)15) + K2_RETRY_CAP * 2;
    unsigned char* wbase = smem_raw + per_warp * wid;
    unsigned long long* s_klo = (unsigned long long*)wbase;
    unsigned long long* retry_w = s_klo + T;
    uint32_t* s_cnt = (uint32_t*)(retry_w + K2_RETRY_CAP);
    uint16_t* occ_w = (uint16_t*)(s_cnt + T);
    uint16_t* retry_m = (uint16_t*)(wbase + per_warp - K2_RETRY_CAP * 2);
    uint32_t* s_hist = (uint32_t*)(smem_raw + per_warp * NWARP);

    for (int i = lane; i < T; i += 32) { s_klo[i] = EMPTY64; s_cnt[i] = 0; }
    for (int i = tid; i < K2_HB; i += NT) s_hist[i] = 0;
    __syncthreads ();

    unsigned long long n_distinct = 0, n_solid = 0, n_emitted = 0, n_once = 0;    // n_once: k-mers seen once (most of them: kept out of the shared histogram)
    unsigned long long out_pos = 0, out_end = 0;
    const uint32_t G = gridDim.x * NWARP;
    const uint4 zero4 = make_uint4 (0, 0, 0, 0);

    // software pipeline over this warp's bins: (bin0, d0, base0, rec0) is current; d1/co1 of the next bin are in flight.
    // With a bin list (the warp tier of the overflow path: the bins a smaller table could not hold) work item g is bin_list[g].
    constexpr uint32_t NOBIN = 0xFFFFFFFFu;
    const uint32_t n_work = P.bin_list ? P.n_list : P.nbins;
    auto bin_of = [&] (uint32_t g) -> uint32_t { return g < n_work ? (P.bin_list ? __ldg (P.bin_list + g) : g) : NOBIN; };
    uint32_t g1 = blockIdx.x * NWARP + wid;
    uint32_t bin0 = bin_of (g1);
    // (descriptor counts keep a flag in the top bit: the dedup split marks the bins it knows to be too large for a first-tier
    //  table; the first-tier run -- no bin list -- hands them to the next tier untouched, the runs over a list ignore the flag)
    const uint32_t ymask = P.bin_list ? K2_DESC_COUNT : 0xFFFFFFFFu;
    auto desc_of = [&] (uint32_t bin) -> uint2 { uint2 d = make_uint2 (0, 0); if (bin != NOBIN) { d = P.bin_desc[bin]; d.y &= ymask; } return d; };
    uint2 d0 = desc_of (bin0);
    unsigned long long base0 = bin0 != NOBIN ? P.coarse_off[bin0 >> P.fine_bits] + d0.x : 0;
    uint4 rec0 = ((uint32_t)lane < d0.y && !(d0.y & K2_DESC_BIG)) ? __ldg ((const uint4*)P.recs + base0 + lane) : zero4;
    g1 += G;
    uint32_t bin1 = bin_of (g1);
    uint2 d1 = desc_of (bin1);
    unsigned long long co1 = bin1 != NOBIN ? P.coarse_off[bin1 >> P.fine_bits] : 0;

    while (bin0 != NOBIN)
    {
        // ---- requests for the following bins ----
        const unsigned long long base1 = co1 + d1.x;
        const uint4 rec1 = ((uint32_t)lane < d1.y && !(d1.y & K2_DESC_BIG)) ? __ldg ((const uint4*)P.recs + base1 + lane) : zero4;
        g1 += G;
        const uint32_t bin2 = bin_of (g1);
        const uint2 d2 = desc_of (bin2);
        const unsigned long long co2 = bin2 != NOBIN ? P.coarse_off[bin2 >> P.fine_bits] : 0;

        const uint32_t n = d0.y;
        // a bin with this many records (five times the planned load) practically never fits the warp's table: hand it to the
        // next tier untouched instead of filling the table first (any bin may go there, this only saves the wasted attempt);
        // likewise the bins the dedup split flagged (n has the flag bit then)
        if (n > (uint32_t)T / 3)
        {
            if (lane == 0) { const uint32_t idx = (uint32_t) atomicAdd (&P.counters[P.ovf_counter], 1ULL); P.ovf_list[idx] = bin0; }
        }
        else if (n)
        {
            uint32_t wn = 0, rn = 0;                              // warp-uniform: claimed slots, pending retries
            bool w_ovf = false;
            auto append_new = [&] (bool isnew, uint32_t slot)
            {
                const unsigned m = __ballot_sync (FULL_MASK, isnew);
                if (m)
                {
                    const uint32_t idx = wn + __popc (m & lt_mask);
                    if (isnew && idx < OCC_W) occ_w[idx] = (uint16_t)slot;
                    wn += __popc (m);
                }
            };
            auto drain_retries = [&] ()
            {
                __syncwarp ();
                for (uint32_t e0 = 0; e0 < rn; e0 += 32)
                {
                    const uint32_t e = e0 + lane;
                    int res = 0;
                    if (e < rn)
                    {
                        const unsigned long long key = retry_w[e];
                        res = k2_probe_loop (s_klo, s_cnt, k2_slot32 ((uint32_t)key, (uint32_t)(key >> 32), hshift), key, tmask, (uint32_t)retry_m[e]);
                        if (res == -1) w_ovf = true;
                    }
                    __syncwarp ();
                    append_new (e < rn && res != -1 && (res & 0x80000000), (uint32_t)res & 0xFFFFu);
                }
                rn = 0;
                __syncwarp ();
            };

            for (uint32_t g0 = 0; g0 < n; g0 += 32)
            {
                uint4 rec = rec0;
                if (g0) rec = (g0 + lane < n) ? __ldg ((const uint4*)P.recs + base0 + g0 + lane) : zero4;
                const uint32_t nch = (((rec.w >> (DEV_LEN_SHIFT_W1 - 32)) & 31u) + 3u) >> 2;      // 0 for the zero record
                uint32_t incl = nch;
                #pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync (FULL_MASK, incl, o); if (lane >= o) incl += y; }
                const uint32_t excl = incl - nch;
                const uint32_t total = __shfl_sync (FULL_MASK, incl, 31);
                uint32_t r0 = 0;
                for (uint32_t wb = 0; wb < total; wb += 32)
                {
                    const uint32_t h = excl - wb;
                    const uint32_t M = __reduce_or_sync (FULL_MASK, (nch > 0 && h < 32u) ? (1u << h) : 0u);
                    const uint32_t gk = wb + lane;
                    const bool act = gk < total;
                    uint32_t r = r0 + __popc (M & (0xFFFFFFFFu >> (31 - lane))) - 1;
                    r0 += __popc (M);
                    r &= 31u;
                    const uint32_t ex = __shfl_sync (FULL_MASK, excl, r);
                    uint4 q;
                    q.x = __shfl_sync (FULL_MASK, rec.x, r); q.y = __shfl_sync (FULL_MASK, rec.y, r);
                    q.z = __shfl_sync (FULL_MASK, rec.z, r); q.w = __shfl_sync (FULL_MASK, rec.w, r);
                    const int c = act ? (int)(gk - ex) : 0;
                    const int nkc = act ? (int)((q.w >> (DEV_LEN_SHIFT_W1 - 32)) & 31u) - 4 * c : 0;      // k-mers of this chunk
                    K2Chunk C;
                    uint32_t lo[4], hi[4], slot[4];
                    const uint32_t mult = q.w >> (DEV_FINE_SHIFT_W1 - 32);                  // multiplicity of the record (k2a)
                    if (ORI)
                    {
                        k2_chunk_begin_raw (C, q.x, q.y, q.z, q.w & ((1u << (DEV_LEN_SHIFT_W1 - 32)) - 1), c, k);
                        k2_chunk_kmer_raw<0> (C, lo[0], hi[0]); k2_chunk_kmer_raw<1> (C, lo[1], hi[1]);
                        k2_chunk_kmer_raw<2> (C, lo[2], hi[2]); k2_chunk_kmer_raw<3> (C, lo[3], hi[3]);
                    }
                    else
                    {
                        k2_chunk_begin (C, q.x, q.y, q.z, q.w & ((1u << (DEV_LEN_SHIFT_W1 - 32)) - 1), c, k);
                        k2_chunk_kmer<0> (C, lo[0], hi[0]); k2_chunk_kmer<1> (C, lo[1], hi[1]);
                        k2_chunk_kmer<2> (C, lo[2], hi[2]); k2_chunk_kmer<3> (C, lo[3], hi[3]);
                    }
                    unsigned long long cur[4];
                    #pragma unroll
                    for (int i = 0; i < 4; i++) { slot[i] = k2_slot32 (lo[i], hi[i], hshift); cur[i] = (i < nkc) ? s_klo[slot[i]] : 0ULL; }
                    #pragma unroll
                    for (int i = 0; i < 4; i++)
                    {
                        const bool valid = i < nkc;
                        const unsigned long long key = ((unsigned long long)hi[i] << 32) | lo[i];
                        unsigned long long cv = cur[i];
                        const bool isE = valid && cv == EMPTY64;
                        if (isE) cv = atomicCAS (&s_klo[slot[i]], EMPTY64, key);
                        const bool isnew = isE && cv == EMPTY64;
                        const bool hit = valid && (isnew || cv == key);
                        if (hit && mult > (isnew ? 1u : 0u)) atomicAdd (&s_cnt[slot[i]], isnew ? mult - 1u : mult);     // a claim counts one by itself
                        append_new (isnew, slot[i]);
                        const bool miss = valid && !hit;
                        const unsigned mm = __ballot_sync (FULL_MASK, miss);
                        if (mm)
                        {
                            if (miss) { retry_w[rn + __popc (mm & lt_mask)] = key; retry_m[rn + __popc (mm & lt_mask)] = (uint16_t)mult; }
                            rn += __popc (mm);
                        }
                    }
                    if (rn >= 32) drain_retries ();
                    if (wn > OCC_W) break;                      // the bin cannot fit any more: it goes to the next tier as a whole
                }
                if (wn > OCC_W) break;
            }
            if (rn) drain_retries ();
            __syncwarp ();
            if (__any_sync (FULL_MASK, w_ovf) || wn > OCC_W)
            {   // the bin goes to the next tier: wipe the warp's table
                if (lane == 0) { const uint32_t idx = (uint32_t) atomicAdd (&P.counters[P.ovf_counter], 1ULL); P.ovf_list[idx] = bin0; }
                for (int i = lane; i < T; i += 32) { s_klo[i] = EMPTY64; s_cnt[i] = 0; }
                __syncwarp ();
            }
            else
            {
                for (uint32_t q0 = 0; q0 < wn; q0 += 32)
                {
                    const uint32_t q = q0 + lane;
                    bool emit = false; uint64_t klo = 0; uint32_t c = 0;
                    if (q < wn)
                    {
                        const uint32_t slot = occ_w[q];
                        c = s_cnt[slot] + 1u; klo = s_klo[slot];            // the claim itself counts one (k2_common.cuh)
                        s_klo[slot] = EMPTY64; s_cnt[slot] = 0;
                        n_distinct++;
                        const uint32_t hb = c >= (uint32_t)P.histo_max ? (uint32_t)P.histo_max : c;
                        if (hb == 1) n_once++; else if (hb < K2_HB) atomicAdd (&s_hist[hb], 1u); else atomicAdd (&P.histogram[hb], 1ULL);
                        if (c >= P.solid_min && c <= P.solid_max) n_solid++;
                        emit = (c >= P.emit_min && c <= P.emit_max);
                    }
                    const unsigned ballot = __ballot_sync (FULL_MASK, emit);
                    if (ballot)
                    {
                        const unsigned ne = __popc (ballot);
                        if (out_pos + ne > out_end)
                        {
                            for (unsigned long long hpos = out_pos + lane; hpos < out_end; hpos += 32)
                                if (hpos < P.out_cap) P.out_lo[hpos] = EMPTY64;
                            unsigned long long b0 = 0;
                            if (lane == 0) b0 = atomicAdd (&P.counters[0], (unsigned long long)WBLOCK);
                            b0 = __shfl_sync (FULL_MASK, b0, 0);
                            out_pos = b0; out_end = b0 + WBLOCK;
                        }
                        if (emit)
                        {
                            const unsigned long long pos = out_pos + __popc (ballot & lt_mask);
                            n_emitted++;
                            if (pos < P.out_cap) { P.out_lo[pos] = ORI ? k2_raw_to_canonical (klo, k) : klo; P.out_cnt[pos] = c; }
                        }
                        out_pos += ne;
                    }
                }
                __syncwarp ();
            }
        }
        bin0 = bin1; d0 = d1; base0 = base1; rec0 = rec1;
        bin1 = bin2; d1 = d2; co1 = co2;
    }
    for (unsigned long long hpos = out_pos + lane; hpos < out_end; hpos += 32)
        if (hpos < P.out_cap) P.out_lo[hpos] = EMPTY64;
    __syncthreads ();
    for (int i = tid; i < K2_HB; i += NT) { uint32_t v = s_hist[i]; if (v) atomicAdd (&P.histogram[i], (unsigned long long)v); }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        n_distinct += __shfl_xor_sync (FULL_MASK, n_distinct, o); n_solid += __shfl_xor_sync (FULL_MASK, n_solid, o);
        n_emitted += __shfl_xor_sync (FULL_MASK, n_emitted, o); n_once += __shfl_xor_sync (FULL_MASK, n_once, o);
    }
    if (lane == 0)
    {
        if (n_once) atomicAdd (&P.histogram[P.histo_max >= 1 ? 1 : P.histo_max], n_once);
        if (n_distinct) atomicAdd (&P.counters[1], n_distinct);
        if (n_solid)    atomicAdd (&P.counters[2], n_solid);
        if (n_emitted)  atomicAdd (&P.counters[6], n_emitted);
    }
}
static size_t k2b_warp_smem_bytes (int table_log2, int nt)
{
    const size_t T = (size_t)1 << table_log2, occ = (T * 3) / 4;
    const size_t per_warp = T * 8 + K2_RETRY_CAP * 8 + T * 4 + ((occ * 2 + 15) & ~(size_t)15) + K2_RETRY_CAP * 2;
    return per_warp * (nt / 32) + K2_HB * 4;
}
template<int NT, bool ORI>
static cudaError_t k2b_warp_launch (const LaunchCtx& L, const K2Params& P)
{
    const size_t smem = k2b_warp_smem_bytes (P.table_log2, NT);
    cudaError_t e = cudaFuncSetAttribute (k2b_warp_bins<NT, ORI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor (&per_sm, k2b_warp_bins<NT, ORI>, NT, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    uint64_t grid = (uint64_t)L.sm_count * per_sm;
    const uint64_t n_work = P.bin_list ? P.n_list : P.nbins;
    const uint64_t need = (n_work + NT / 32 - 1) / (NT / 32);
    if (grid > need) grid = need;
    k2b_warp_bins<NT, ORI><<<(unsigned)grid, NT, smem, L.stream>>> (P);
    (*L.launches)++;
    return cudaGetLastError ();
}

// ------------------------------------------------------------------------------------------------ k2b, 32 <= k <= 63, warp per bin
// The warp-per-bin scheme of k2b_warp_bins for 128-bit k-mers (Kmer<64>, 32-byte records): private table (low and high
// halves of the keys in two arrays), records staged per 32 in the warp's shared memory (a chunk reads the six words it
// needs at its own word offset), chunks of four k-mers decoded by k2_chunk2_* (k2_decode.cuh).
// Claiming a 128-bit key: the high half is taken with a 64-bit CAS from EMPTY to LOCK, the winner publishes the low
// half and then the real high half.  The table is private to the warp and the warp stays converged through a step, so
// after the __syncwarp that follows the publication nobody can still see LOCK: no spinning in the converged path.
// Collisions go to the warp's retry list and are re-inserted 32 at a time with table_insert_w2 (general probing).
#define K2W2_RETRY_CAP 96       // 31 left over + 2 steps x 32 lanes between two drain checks
template<int NT>
__global__ void __launch_bounds__(NT, 512 / NT) k2b_warp_bins_w2 (const K2Params P)
{
    constexpr int NWARP = NT / 32;
    constexpr unsigned WBLOCK = 2048;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int T = 1 << P.table_log2;
    const int hshift = 32 - P.table_log2;
    const uint32_t OCC_W = (uint32_t)(T * 3) / 4;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1;
    const int k = P.k;
    // per warp: staging 32 x 32 B | keys lo T*8 | keys hi T*8 | retry lo/hi CAP*16 | counts T*4 | claimed OCC_W*2 (16-aligned)
    const size_t per_warp = 1024 + (size_t)T * 16 + K2W2_RETRY_CAP * 16 + (size_t)T * 4 + (((size_t)OCC_W * 2 + 15) & ~(size_t)15);
    unsigned char* wbase = smem_raw + per_warp * wid;
    uint4* stage = (uint4*)wbase;
    const uint32_t* stage32 = (const uint32_t*)wbase;
    unsigned long long* s_klo = (unsigned long long*)(wbase + 1024);
    unsigned long long* s_khi = s_klo + T;
    unsigned long long* retry_lo = s_khi + T;
    unsigned long long* retry_hi = retry_lo + K2W2_RETRY_CAP;
    uint32_t* s_cnt = (uint32_t*)(retry_hi + K2W2_RETRY_CAP);
    uint16_t* occ_w = (uint16_t*)(s_cnt + T);
    uint32_t* s_hist = (uint32_t*)(smem_raw + per_warp * NWARP);

    for (int i = lane; i < T; i += 32) { s_khi[i] = EMPTY64; s_klo[i] = 0; s_cnt[i] = 0; }
    for (int i = tid; i < K2_HB; i += NT) s_hist[i] = 0;
    __syncthreads ();

    uint32_t n_distinct = 0, n_solid = 0, n_emitted = 0;           // per lane; flushed as 64-bit sums at the end
    unsigned long long out_pos = 0, out_end = 0;
    const uint32_t G = gridDim.x * NWARP;
    const uint4 zero4 = make_uint4 (0, 0, 0, 0);
    const uint4* recs = (const uint4*)P.recs;

    // software pipeline over this warp's bins: (bin0, d0, base0, ra0, rb0) is current; d1/co1 of the next bin are in flight
    uint32_t bin0 = blockIdx.x * NWARP + wid;
    uint2 d0 = bin0 < P.nbins ? P.bin_desc[bin0] : make_uint2 (0, 0);
    unsigned long long base0 = bin0 < P.nbins ? P.coarse_off[bin0 >> P.fine_bits] + d0.x : 0;
    uint4 ra0 = ((uint32_t)lane < d0.y) ? __ldg (recs + 2 * (base0 + lane)) : zero4;
    uint4 rb0 = ((uint32_t)lane < d0.y) ? __ldg (recs + 2 * (base0 + lane) + 1) : zero4;
    uint32_t bin1 = bin0 + G;
    uint2 d1 = bin1 < P.nbins ? P.bin_desc[bin1] : make_uint2 (0, 0);
    unsigned long long co1 = bin1 < P.nbins ? P.coarse_off[bin1 >> P.fine_bits] : 0;

    while (bin0 < P.nbins)
    {
        const unsigned long long base1 = co1 + d1.x;
        const uint4 ra1 = ((uint32_t)lane < d1.y) ? __ldg (recs + 2 * (base1 + lane)) : zero4;
        const uint4 rb1 = ((uint32_t)lane < d1.y) ? __ldg (recs + 2 * (base1 + lane) + 1) : zero4;
        const uint32_t bin2 = bin1 + G;
        const uint2 d2 = bin2 < P.nbins ? P.bin_desc[bin2] : make_uint2 (0, 0);
        const unsigned long long co2 = bin2 < P.nbins ? P.coarse_off[bin2 >> P.fine_bits] : 0;

        const uint32_t n = d0.y;
        if (n)
        {
            uint32_t wn = 0, rn = 0;                              // warp-uniform: claimed slots, pending retries
            bool w_ovf = false;
            auto append_new = [&] (bool isnew, uint32_t slot)
            {
                const unsigned m = __ballot_sync (FULL_MASK, isnew);
                if (m)
                {
                    const uint32_t idx = wn + __popc (m & lt_mask);
                    if (isnew && idx < OCC_W) occ_w[idx] = (uint16_t)slot;
                    wn += __popc (m);
                }
            };
            auto drain_retries = [&] ()
            {
                __syncwarp ();
                for (uint32_t e0 = 0; e0 < rn; e0 += 32)
                {
                    const uint32_t e = e0 + lane;
                    int res = 0;
                    if (e < rn)
                    {
                        u128 key; key.lo = retry_lo[e]; key.hi = retry_hi[e];
                        const uint32_t slot = ((uint32_t)key.lo * 0x9E3779B1u + (uint32_t)(key.lo >> 32) * 0x85EBCA77u +
                                               (uint32_t)key.hi * 0xC2B2AE3Du + (uint32_t)(key.hi >> 32) * 0x27D4EB2Fu) >> hshift;
                        res = table_insert_w2<false> (s_klo, s_khi, s_cnt, P.table_log2, slot, key, K2_MAXPROBE);
                        if (res < 0) w_ovf = true;
                    }
                    __syncwarp ();
                    append_new (e < rn && res >= 0 && (res & 0x40000000), (uint32_t)res & 0xFFFFu);
                }
                rn = 0;
                __syncwarp ();
            };

            for (uint32_t g0 = 0; g0 < n; g0 += 32)
            {
                uint4 ra = ra0, rb = rb0;
                if (g0)
                {
                    ra = (g0 + lane < n) ? __ldg (recs + 2 * (base0 + g0 + lane)) : zero4;
                    rb = (g0 + lane < n) ? __ldg (recs + 2 * (base0 + g0 + lane) + 1) : zero4;
                }
                __syncwarp ();                                      // the previous group (or bin) is fully consumed
                const uint32_t len_own = (rb.w >> (REC_LEN_SHIFT_W2 - 32)) & 63u;
                rb.w &= (1u << (REC_LEN_SHIFT_W2 - 32)) - 1;        // nucleotides only in the staged copy
                stage[2 * lane] = ra; stage[2 * lane + 1] = rb;
                const uint32_t nch = (len_own + 3u) >> 2;           // 0 for the zero record
                uint32_t incl = nch;
                #pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync (FULL_MASK, incl, o); if (lane >= o) incl += y; }
                const uint32_t excl = incl - nch;
                const uint32_t total = __shfl_sync (FULL_MASK, incl, 31);
                const uint32_t own = (excl << 8) | len_own;          // first chunk id and k-mer count of the lane's record
                __syncwarp ();
                uint32_t r0 = 0;
                for (uint32_t wb = 0; wb < total; wb += 32)
                {
                    const uint32_t h = excl - wb;
                    const uint32_t M = __reduce_or_sync (FULL_MASK, (nch > 0 && h < 32u) ? (1u << h) : 0u);
                    const uint32_t gk = wb + lane;
                    const bool act = gk < total;
                    uint32_t r = r0 + __popc (M & (0xFFFFFFFFu >> (31 - lane))) - 1;
                    r0 += __popc (M);
                    r &= 31u;
                    const uint32_t ow = __shfl_sync (FULL_MASK, own, r);
                    const int c = act ? (int)(gk - (ow >> 8)) : 0;
                    const int nkc = act ? (int)(ow & 0xFFu) - 4 * c : 0;      // k-mers of this chunk
                    K2Chunk2 C;
                    k2_chunk2_begin (C, stage32 + 8 * r, c, k);
                    uint32_t v[4][4], slot[4];
                    k2_chunk2_kmer<0> (C, v[0]); k2_chunk2_kmer<1> (C, v[1]); k2_chunk2_kmer<2> (C, v[2]); k2_chunk2_kmer<3> (C, v[3]);
                    unsigned long long cur[4];
                    #pragma unroll
                    for (int i = 0; i < 4; i++)
                    {
                        slot[i] = (v[i][0] * 0x9E3779B1u + v[i][1] * 0x85EBCA77u + v[i][2] * 0xC2B2AE3Du + v[i][3] * 0x27D4EB2Fu) >> hshift;
                        cur[i] = (i < nkc) ? s_khi[slot[i]] : 0ULL;
                    }
                    #pragma unroll
                    for (int i = 0; i < 4; i++)
                    {
                        const bool valid = i < nkc;
                        const unsigned long long klo = ((unsigned long long)v[i][1] << 32) | v[i][0];
                        const unsigned long long khi = ((unsigned long long)v[i][3] << 32) | v[i][2];
                        unsigned long long ch = cur[i];
                        const bool isE = valid && ch == EMPTY64;
                        if (isE) ch = atomicCAS (&s_khi[slot[i]], EMPTY64, LOCK64);
                        const bool isnew = isE && ch == EMPTY64;
                        if (isnew) { s_klo[slot[i]] = klo; s_khi[slot[i]] = khi; }            // publish: low half, then the real high half
                        __syncwarp ();
                        bool hit = isnew;
                        if (valid && !isnew) hit = (s_khi[slot[i]] == khi) && (s_klo[slot[i]] == klo);
                        if (hit) atomicAdd (&s_cnt[slot[i]], 1u);
                        append_new (isnew, slot[i]);
                        const bool miss = valid && !hit;
                        const unsigned mm = __ballot_sync (FULL_MASK, miss);
                        if (mm)
                        {
                            const uint32_t at = rn + __popc (mm & lt_mask);
                            if (miss) { retry_lo[at] = klo; retry_hi[at] = khi; }
                            rn += __popc (mm);
                        }
                        if ((i & 1) && rn >= 32) drain_retries ();
                    }
                    if (wn > OCC_W) break;                          // the bin cannot fit any more: it goes to the fallback as a whole
                }
                if (wn > OCC_W) break;
            }
            if (rn) drain_retries ();
            __syncwarp ();
            if (__any_sync (FULL_MASK, w_ovf) || wn > OCC_W)
            {   // the bin goes to the global-memory fallback (k2c): wipe the warp's table
                if (lane == 0) { const uint32_t idx = (uint32_t) atomicAdd (&P.counters[4], 1ULL); P.ovf_list[idx] = bin0; }
                for (int i = lane; i < T; i += 32) { s_khi[i] = EMPTY64; s_klo[i] = 0; s_cnt[i] = 0; }
                __syncwarp ();
            }
            else
            {
                for (uint32_t q0 = 0; q0 < wn; q0 += 32)
                {
                    const uint32_t q = q0 + lane;
                    bool emit = false; uint64_t klo = 0, khi = 0; uint32_t c = 0;
                    if (q < wn)
                    {
                        const uint32_t slot = occ_w[q];
                        c = s_cnt[slot]; klo = s_klo[slot]; khi = s_khi[slot];
                        s_khi[slot] = EMPTY64; s_klo[slot] = 0; s_cnt[slot] = 0;
                        n_distinct++;
                        const uint32_t hb = c >= (uint32_t)P.histo_max ? (uint32_t)P.histo_max : c;
                        if (hb < K2_HB) atomicAdd (&s_hist[hb], 1u); else atomicAdd (&P.histogram[hb], 1ULL);
                        if (c >= P.solid_min && c <= P.solid_max) n_solid++;
                        emit = (c >= P.emit_min && c <= P.emit_max);
                    }
                    const unsigned ballot = __ballot_sync (FULL_MASK, emit);
                    if (ballot)
                    {
                        const unsigned ne = __popc (ballot);
                        if (out_pos + ne > out_end)
                        {
                            for (unsigned long long hpos = out_pos + lane; hpos < out_end; hpos += 32)
                                if (hpos < P.out_cap) P.out_hi[hpos] = EMPTY64;
                            unsigned long long b0 = 0;
                            if (lane == 0) b0 = atomicAdd (&P.counters[0], (unsigned long long)WBLOCK);
                            b0 = __shfl_sync (FULL_MASK, b0, 0);
                            out_pos = b0; out_end = b0 + WBLOCK;
                        }
                        if (emit)
                        {
                            const unsigned long long pos = out_pos + __popc (ballot & lt_mask);
                            n_emitted++;
                            if (pos < P.out_cap) { P.out_lo[pos] = klo; P.out_hi[pos] = khi; P.out_cnt[pos] = c; }
                        }
                        out_pos += ne;
                    }
                }
                __syncwarp ();
            }
        }
        bin0 = bin1; d0 = d1; base0 = base1; ra0 = ra1; rb0 = rb1;
        bin1 = bin2; d1 = d2; co1 = co2;
    }
    for (unsigned long long hpos = out_pos + lane; hpos < out_end; hpos += 32)
        if (hpos < P.out_cap) P.out_hi[hpos] = EMPTY64;
    __syncthreads ();
    for (int i = tid; i < K2_HB; i += NT) { uint32_t v = s_hist[i]; if (v) atomicAdd (&P.histogram[i], (unsigned long long)v); }
    unsigned long long t_distinct = n_distinct, t_solid = n_solid, t_emitted = n_emitted;
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        t_distinct += __shfl_xor_sync (FULL_MASK, t_distinct, o); t_solid += __shfl_xor_sync (FULL_MASK, t_solid, o);
        t_emitted += __shfl_xor_sync (FULL_MASK, t_emitted, o);
    }
    if (lane == 0)
    {
        if (t_distinct) atomicAdd (&P.counters[1], t_distinct);
        if (t_solid)    atomicAdd (&P.counters[2], t_solid);
        if (t_emitted)  atomicAdd (&P.counters[6], t_emitted);
    }
}
static size_t k2b_warp_w2_smem_bytes (int table_log2, int nt)
{
    const size_t T = (size_t)1 << table_log2, occ = (T * 3) / 4;
    const size_t per_warp = 1024 + T * 16 + K2W2_RETRY_CAP * 16 + T * 4 + ((occ * 2 + 15) & ~(size_t)15);
    return per_warp * (nt / 32) + K2_HB * 4;
}
template<int NT>
static cudaError_t k2b_warp_w2_launch (const LaunchCtx& L, const K2Params& P)
{
    const size_t smem = k2b_warp_w2_smem_bytes (P.table_log2, NT);
    cudaError_t e = cudaFuncSetAttribute (k2b_warp_bins_w2<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor (&per_sm, k2b_warp_bins_w2<NT>, NT, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    uint64_t grid = (uint64_t)L.sm_count * per_sm;
    const uint64_t need = (P.nbins + NT / 32 - 1) / (NT / 32);
    if (grid > need) grid = need;
    k2b_warp_bins_w2<NT><<<(unsigned)grid, NT, smem, L.stream>>> (P);
    (*L.launches)++;
    return cudaGetLastError ();
}

static size_t k2b_smem_bytes (int W, int table_log2);
// second tier: the bins the warp kernel could not hold, counted by CTAs with the table size of P.table_log2
cudaError_t launch_k2b_count_list (const LaunchCtx& L, const K2Params& P)
{
    if (P.n_list == 0) return cudaSuccess;
    if (P.W == 2)
    {   // 32 <= k <= 63: the CTA-per-bin kernel with a larger table over the list
        const size_t smem2 = k2b_smem_bytes (2, P.table_log2);
        cudaError_t e2 = cudaFuncSetAttribute (k2b_bucket_hash_count<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (e2 != cudaSuccess) return e2;
        int per_sm2 = 0;
        e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor (&per_sm2, k2b_bucket_hash_count<2>, K2_THREADS, smem2);
        if (e2 != cudaSuccess) return e2;
        if (per_sm2 < 1) per_sm2 = 1;
        uint64_t grid2 = (uint64_t)L.sm_count * per_sm2;
        if (grid2 > P.n_list) grid2 = P.n_list;
        k2b_bucket_hash_count<2><<<(unsigned)grid2, K2_THREADS, smem2, L.stream>>> (P);
        (*L.launches)++;
        return cudaGetLastError ();
    }
    // tables of up to 2048 slots still belong to one warp each (k2b_warp_bins over the bin list): five times the rate of the CTA tiers
    if (P.W == 1 && P.table_log2 <= 11 && k2b_variant (P.path_flags) == 1)
        return P.oriented ? k2b_warp_launch<128, true> (L, P) : k2b_warp_launch<128, false> (L, P);
    const size_t smem = k2b_w1_smem_bytes (P.table_log2, 256);
    cudaError_t e = cudaFuncSetAttribute (k2b_count_w1<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor (&per_sm, k2b_count_w1<256>, 256, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    uint64_t grid = (uint64_t)L.sm_count * per_sm;
    if (grid > P.n_list) grid = P.n_list;
    k2b_count_w1<256><<<(unsigned)grid, 256, smem, L.stream>>> (P);
    (*L.launches)++;
    return cudaGetLastError ();
}

// which counting kernel serves k <= 31 (gatb_gpu_params.path_flags & GATB_PATH_K2B_MASK): 1 = warp per bin (default),
// 128 / 256 = CTA per bin with the chunked insert, 0 = CTA per bin, one k-mer per lane.  The table size the planner picks follows.
int k2b_variant (int path_flags)
{
    switch (path_flags & 6) { case 2: return 128; case 4: return 256; case 6: return 0; }
    return 1;
}
// 32 <= k <= 63: the warp-per-bin kernel (k2b_warp_bins_w2) passes the whole GPU parity suite, but its 512-slot tables are too
// small for the long super-k-mers of k = 63 and it has no overflow tiers yet: opt-in (GATB_PATH_K2B_W2_WARP).
static bool k2b_w2_warp (int path_flags) { return (path_flags & 8) != 0; }
int k2b_default_table_log2 (int W, int path_flags) { return (k2b_variant (path_flags) == 1 && (W == 1 || k2b_w2_warp (path_flags))) ? 9 : 11; }

static size_t k2b_smem_bytes (int W, int table_log2)
{
    size_t T = (size_t)1 << table_log2, occ = (T * 3) / 4; occ += occ & 1;
    size_t CHR = (W == 1) ? 256 : 128;
    return 2 * CHR * 16 * W + T * 8 * W + T * 4 + occ * 2 + K2_HB * 4 + 32;
}

cudaError_t launch_k2b_count (const LaunchCtx& L, const K2Params& P)
{
    if (P.nbins == 0) return cudaSuccess;
    size_t smem = k2b_smem_bytes (P.W, P.table_log2);
    const int variant = k2b_variant (P.path_flags);
    if (P.W == 1 && variant == 1) return P.oriented ? k2b_warp_launch<128, true> (L, P) : k2b_warp_launch<128, false> (L, P);
    if (P.W == 2 && variant == 1 && k2b_w2_warp (P.path_flags)) return k2b_warp_w2_launch<128> (L, P);
    if (P.W == 1 && variant == 128) return k2b_w1_launch<128> (L, P);
    if (P.W == 1 && variant == 256) return k2b_w1_launch<256> (L, P);
    const void* fn = (P.W == 1) ? (const void*)k2b_bucket_hash_count<1> : (const void*)k2b_bucket_hash_count<2>;
    cudaError_t e = cudaFuncSetAttribute (fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor (&per_sm, fn, K2_THREADS, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    uint64_t grid = (uint64_t)L.sm_count * per_sm;
    if (grid > P.nbins) grid = P.nbins;
    if (P.W == 1) k2b_bucket_hash_count<1><<<(unsigned)grid, K2_THREADS, smem, L.stream>>> (P);
    else          k2b_bucket_hash_count<2><<<(unsigned)grid, K2_THREADS, smem, L.stream>>> (P);
    (*L.launches)++;
    return cudaGetLastError ();
}
unsigned k2b_max_grid (const LaunchCtx& L) { return (unsigned)L.sm_count * 4; }

// ------------------------------------------------------------------------------------------------ k2c: global fallback
// For the (rare) bins whose distinct k-mers exceed the shared-memory table: all of them share ONE global-memory table
// (different bins hold disjoint k-mers), then the table is scanned like in k2b.
template<int W>
__global__ void __launch_bounds__(256) k2c_measure (const K2Params P, uint32_t n_ovf)
{
    unsigned long long sum = 0;
    for (uint32_t o = blockIdx.x; o < n_ovf; o += gridDim.x)
    {
        const uint32_t bin = P.ovf_list[o];
        uint2 d = P.bin_desc[bin]; d.y &= K2_DESC_COUNT;
        const uint4* base = (const uint4*)P.recs + (P.coarse_off[bin >> P.fine_bits] + d.x) * W;
        for (uint32_t i = threadIdx.x; i < d.y; i += blockDim.x)
        {
            uint4 last = __ldg (&base[(uint64_t)i * W + (W - 1)]);
            uint64_t hi = (uint64_t)last.z | ((uint64_t)last.w << 32);
            sum += (W == 1) ? ((hi >> DEV_LEN_SHIFT_W1) & 31) : ((hi >> REC_LEN_SHIFT_W2) & 63);
        }
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync (FULL_MASK, sum, o);
    if ((threadIdx.x & 31) == 0 && sum) atomicAdd (&P.counters[5], sum);
}

template<int W>
__global__ void __launch_bounds__(256) k2c_insert (const K2Params P, uint32_t n_ovf)
{
    const int k = P.k;
    for (uint32_t o = blockIdx.y; o < n_ovf; o += gridDim.y)
    {
        const uint32_t bin = P.ovf_list[o];
        uint2 d = P.bin_desc[bin]; d.y &= K2_DESC_COUNT;
        const uint4* base = (const uint4*)P.recs + (P.coarse_off[bin >> P.fine_bits] + d.x) * W;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < d.y; i += gridDim.x * blockDim.x)
        {
            if (W == 1)
            {
                uint4 r = __ldg (&base[i]);
                uint64_t lo = (uint64_t)r.x | ((uint64_t)r.y << 32), hi = (uint64_t)r.z | ((uint64_t)r.w << 32);
                const int len = (int)((hi >> DEV_LEN_SHIFT_W1) & 31);
                const uint32_t mult = (uint32_t)(hi >> DEV_FINE_SHIFT_W1);              // multiplicity of the record (k2a)
                hi &= (1ULL << DEV_LEN_SHIFT_W1) - 1;
                for (int j = 0; j < len; j++)
                    table_insert_w1 ((unsigned long long*)P.g_lo, P.g_cnt, P.g_log2, rec_kmer_w1 (lo, hi, j, k), 1 << 30, mult);
            }
            else
            {
                uint4 a = __ldg (&base[2*(uint64_t)i]), b = __ldg (&base[2*(uint64_t)i + 1]);
                uint64_t r[4] = { (uint64_t)a.x | ((uint64_t)a.y << 32), (uint64_t)a.z | ((uint64_t)a.w << 32),
                                  (uint64_t)b.x | ((uint64_t)b.y << 32), (uint64_t)b.z | ((uint64_t)b.w << 32) };
                const int len = (int)((r[3] >> REC_LEN_SHIFT_W2) & 63);
                r[3] &= (1ULL << REC_LEN_SHIFT_W2) - 1;
                for (int j = 0; j < len; j++)
                    { u128 key = rec_kmer_w2 (r, j, k); table_insert_w2<true> ((unsigned long long*)P.g_lo, (unsigned long long*)P.g_hi, P.g_cnt, P.g_log2, slot_hash128 (key, P.g_log2), key, 1 << 30); }
            }
        }
    }
}

template<int W>
__global__ void __launch_bounds__(256) k2c_scan (const K2Params P)
{
    const uint64_t T = 1ULL << P.g_log2;
    __shared__ uint32_t s_hist[K2_HB];                     // abundances below K2_HB: one global atomic per CTA and value at the end
    for (int i = threadIdx.x; i < K2_HB; i += blockDim.x) s_hist[i] = 0;
    __syncthreads ();
    EmitState st; st.distinct = 0; st.solid = 0; st.emitted = 0;
    // every thread of a warp runs the same number of iterations (consume_entry uses warp collectives)
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t iters = (T + stride - 1) / stride;
    for (uint64_t it = 0; it < iters; it++)
    {
        uint64_t i = it * stride + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        bool occ = false; uint64_t klo = 0, khi = 0; uint32_t c = 0;
        if (i < T)
        {
            klo = P.g_lo[i]; khi = (W == 2) ? P.g_hi[i] : 0; c = P.g_cnt[i];
            occ = (W == 1) ? (klo != EMPTY64) : (khi != EMPTY64);
        }
        consume_entry (P, occ, klo, khi, c, st, s_hist);
    }
    __syncthreads ();
    for (int i = threadIdx.x; i < K2_HB; i += blockDim.x) { const uint32_t v = s_hist[i]; if (v) atomicAdd (&P.histogram[i], (unsigned long long)v); }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) { st.distinct += __shfl_xor_sync (FULL_MASK, st.distinct, o); st.solid += __shfl_xor_sync (FULL_MASK, st.solid, o); st.emitted += __shfl_xor_sync (FULL_MASK, st.emitted, o); }
    if ((threadIdx.x & 31) == 0) { if (st.distinct) atomicAdd (&P.counters[1], st.distinct); if (st.solid) atomicAdd (&P.counters[2], st.solid); if (st.emitted) atomicAdd (&P.counters[6], st.emitted); }
}

cudaError_t launch_k2c_measure (const LaunchCtx& L, const K2Params& P, uint32_t n_ovf)
{
    unsigned grid = n_ovf < 1024 ? n_ovf : 1024;
    if (P.W == 1) k2c_measure<1><<<grid, 256, 0, L.stream>>> (P, n_ovf); else k2c_measure<2><<<grid, 256, 0, L.stream>>> (P, n_ovf);
    (*L.launches)++;
    return cudaGetLastError ();
}
cudaError_t launch_k2c_insert (const LaunchCtx& L, const K2Params& P, uint32_t n_ovf)
{
    dim3 grid (n_ovf < 64 ? (unsigned)(L.sm_count * 8 / (n_ovf ? n_ovf : 1) + 1) : 8, n_ovf < 4096 ? n_ovf : 4096);
    if (P.W == 1) k2c_insert<1><<<grid, 256, 0, L.stream>>> (P, n_ovf); else k2c_insert<2><<<grid, 256, 0, L.stream>>> (P, n_ovf);
    (*L.launches)++;
    return cudaGetLastError ();
}
cudaError_t launch_k2c_scan (const LaunchCtx& L, const K2Params& P)
{
    unsigned grid = L.sm_count * 8;
    if (P.W == 1) k2c_scan<1><<<grid, 256, 0, L.stream>>> (P); else k2c_scan<2><<<grid, 256, 0, L.stream>>> (P);
    (*L.launches)++;
    return cudaGetLastError ();
}
