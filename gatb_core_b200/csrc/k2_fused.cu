// k2_fused.cu -- k2f_count_coarse: counting straight out of the coarse bins (k <= 31), no fine split.
//
// Replaces (paths relative to /root/reference/gatb-core/src/gatb/):
//   ReadSuperKCommand::execute / hash-mode decode    kmer/impl/PartitionsCommand.cpp:944-1128, 420-501
//   SortCommand + executeDump (sort, merge, count)   kmer/impl/PartitionsCommand.cpp:1400-1445, 1599-1805
//   CountProcessorHistogram::process, CountProcessorSoliditySum::check   kmer/impl/CountProcessorHistogram.hpp:173, CountProcessorSolidity.hpp:186
//
// One CTA owns a coarse bin from its first record to its last emitted k-mer:
//   * a producer lane streams the bin's records (gathered from up to 16 source pieces, blocks of 64 records at the stride
//     of the round-interleaved layout, kernels.h) into a ring of three 16 KB shared-memory tiles with TMA bulk copies
//     (cp.async.bulk + mbarrier: SASS UBLKCP), two tiles ahead of the consumers;
//   * per tile the consumer warps (1) collapse identical records in a small shared-memory table -- oriented records
//     (k1_scan.cuh) of the reads that cover one locus without an error in the span are bit-identical whatever the strand,
//     about half of all records -- (2) compact the survivors, (3) insert their k-mers into ONE table shared by the CTA
//     (2^table_log2 slots, 8-byte keys claimed by CAS, 32-bit counts), lane <-> chunk of four k-mers exactly like
//     k2b_warp_bins, adding the record's multiplicity; with oriented records the key of a k-mer is a plain slice of the
//     record (no reverse complement, no min per occurrence);
//   * after the last tile of the bin the table is scanned once: histogram, solidity statistics, k-mers in the emission
//     range appended to per-warp output blocks (canonical VALUE rebuilt here for oriented keys), slots cleared.
// A bin whose distinct k-mers outgrow the table is dropped untouched into an overflow list; the host sends those bins
// through the fine-split pipeline of k2_count.cu (k2a_fine_split over a bin list + the tier kernels).
// Block barriers per tile: four (tile hand-over, table cleared, records collapsed, survivors compacted), per bin two more.
#include "common.cuh"
#include "kernels.h"
#include "k2_decode.cuh"
#include "k2_common.cuh"

#define K2F_TILE   1024       // records per tile (16 KB)
#define K2F_NBUF   3
#define K2F_DT     2048       // slots of the per-tile record table
#define K2F_RETRY  160        // per warp: 31 left over + 4 steps x 32 lanes at the very worst
#define K2F_WBLOCK 2048       // output slots a warp reserves at a time
#define K2F_END    2u
#define K2F_LAST   1u
#define EMPTY32    0xFFFFFFFFu

template<int NWARP, bool ORI>
__global__ void __launch_bounds__(NWARP * 32 + 32) k2f_count_coarse (const K2Params P, const K2aSrc S, const uint32_t nb, const uint32_t cap,
                                                                       const uint32_t n_bins, const int dedup)
{
    constexpr int NT = NWARP * 32;
    extern __shared__ __align__(128) unsigned char k2f_smem[];
    const int T = 1 << P.table_log2;
    const uint32_t tmask = (uint32_t)T - 1;
    const int hshift = 32 - P.table_log2;
    uint4* tiles = (uint4*)k2f_smem;                                          // [K2F_NBUF][K2F_TILE]
    unsigned long long* keys = (unsigned long long*)(tiles + K2F_NBUF * K2F_TILE);   // [T]
    unsigned long long* retry = keys + T;                                     // [NWARP][K2F_RETRY]
    uint32_t* cnts  = (uint32_t*)(retry + NWARP * K2F_RETRY);                 // [T]
    uint32_t* dtbl  = cnts + T;                                               // [K2F_DT]
    uint32_t* ulist = dtbl + K2F_DT;                                          // [K2F_TILE]
    uint32_t* hist  = ulist + K2F_TILE;                                       // [K2_HB]
    uint16_t* retry_mult = (uint16_t*)(hist + K2_HB);                         // [NWARP][K2F_RETRY]
    __shared__ __align__(8) uint64_t s_bar[K2F_NBUF];
    __shared__ uint32_t s_tn[K2F_NBUF], s_tflags[K2F_NBUF], s_tbin[K2F_NBUF];
    __shared__ uint32_t s_nu, s_ndist, s_fail;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const bool producer = tid >= NT;
    const uint32_t lt_mask = (1u << lane) - 1;
    const int k = P.k;
    const uint32_t dist_limit = (uint32_t)T - (uint32_t)T / 4;               // the bin is given up beyond 75 % load

    for (int i = tid; i < T; i += NT + 32) { keys[i] = EMPTY64; cnts[i] = 0; }
    for (int i = tid; i < K2_HB; i += NT + 32) hist[i] = 0;
    if (tid == 0) { for (int b = 0; b < K2F_NBUF; b++) mbar_init (&s_bar[b], 1); s_ndist = 0; s_fail = 0; s_nu = 0; }
    asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads ();

    // ---- producer state: the tiles of this CTA's bins, in order ----
    uint32_t p_bin = blockIdx.x, p_src = 0, p_off = 0, p_left = 0; bool p_open = false;
    uint32_t p_cnt[K2A_MAXSRC];
    auto issue_tile = [&] (int buf)
    {   // only lane 0 of the producer warp runs this
        for (;;)
        {
            if (!p_open)
            {
                if (p_bin >= n_bins) { s_tflags[buf] = K2F_END; s_tn[buf] = 0; s_tbin[buf] = 0; return; }
                uint32_t tot = 0;
                #pragma unroll
                for (int s = 0; s < K2A_MAXSRC; s++) { p_cnt[s] = (s < S.n) ? min (__ldg (&S.cursors[s][p_bin]), cap) : 0u; tot += p_cnt[s]; }
                if (tot == 0) { p_bin += gridDim.x; continue; }
                p_left = tot; p_src = 0; p_off = 0; p_open = true;
            }
            const uint32_t n = p_left < K2F_TILE ? p_left : K2F_TILE;
            fence_proxy_async ();
            mbar_expect_tx (&s_bar[buf], n * 16);
            uint4* dst = tiles + buf * K2F_TILE;
            uint32_t filled = 0;
            while (filled < n)
            {
                while (p_off >= p_cnt[p_src]) { p_src++; p_off = 0; }
                uint32_t c = S.off[p_src] ? 1024u : COARSE_BLK - (p_off % COARSE_BLK);      // dense sources are contiguous
                if (c > p_cnt[p_src] - p_off) c = p_cnt[p_src] - p_off;
                if (c > n - filled) c = n - filled;
                tma_bulk_g2s (dst + filled, S.bins[p_src] + k2a_record_index (S, (int)p_src, p_bin, p_off, nb), c * 16, &s_bar[buf]);
                filled += c; p_off += c;
            }
            p_left -= n;
            s_tn[buf] = n; s_tbin[buf] = p_bin; s_tflags[buf] = p_left ? 0u : K2F_LAST;
            if (p_left == 0) { p_open = false; p_bin += gridDim.x; }
            return;
        }
    };
    if (producer && lane == 0) { issue_tile (0); issue_tile (1); }

    unsigned long long n_distinct = 0, n_solid = 0, n_emitted = 0, n_once = 0;
    unsigned long long out_pos = 0, out_end = 0;                             // warp-uniform: this warp's output block
    unsigned long long* retry_w = retry + (producer ? 0 : wid) * K2F_RETRY;
    uint16_t* retry_m = retry_mult + (producer ? 0 : wid) * K2F_RETRY;
    const uint4 zero4 = make_uint4 (0, 0, 0, 0);
    bool bin_ovf = false;                                                    // CTA-uniform: the current bin is given up

    for (uint32_t it = 0; ; it++)
    {
        const int buf = (int)(it % K2F_NBUF);
        __syncthreads ();                                                    // tile it-1 is done with; meta of tile it is visible
        if (producer && lane == 0) issue_tile ((int)((it + 2) % K2F_NBUF));
        const uint32_t flags = s_tflags[buf], tn = s_tn[buf];
        if (flags & K2F_END) break;
        bin_ovf = bin_ovf || s_ndist > dist_limit || s_fail != 0;            // read by everybody before anybody changes them again
        mbar_wait (&s_bar[buf], (it / K2F_NBUF) & 1u);
        const uint4* tile = tiles + buf * K2F_TILE;
        uint32_t nu = tn;
        if (dedup && !bin_ovf)
        {   // ---- (1) collapse identical records of the tile ----
            for (int i = tid; i < K2F_DT; i += NT + 32) dtbl[i] = EMPTY32;
            if (tid == 0) s_nu = 0;
            __syncthreads ();
            if (!producer)
                for (uint32_t g = tid; g < tn; g += NT)
                {
                    const uint4 r = tile[g];
                    uint32_t h = (r.x * 0x9E3779B1u) ^ (r.y * 0x85EBCA77u) ^ (r.z * 0xC2B2AE3Du) ^ (r.w * 0x27D4EB2Fu);
                    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 13;
                    h &= K2F_DT - 1;
                    for (;;)
                    {
                        uint32_t e = *(volatile uint32_t*)&dtbl[h];
                        if (e == EMPTY32) { e = atomicCAS (&dtbl[h], EMPTY32, g | (1u << 16)); if (e == EMPTY32) break; }
                        const uint4 o = tile[e & 0xFFFFu];
                        if (o.x == r.x && o.y == r.y && o.z == r.z && o.w == r.w) { atomicAdd (&dtbl[h], 1u << 16); break; }
                        h = (h + 1) & (K2F_DT - 1);
                    }
                }
            __syncthreads ();
            // ---- (2) survivors {record index : 16, multiplicity : 16}, compacted ----
            if (!producer)
                for (int i0 = wid * 32; i0 < K2F_DT; i0 += NT)
                {
                    const uint32_t e = dtbl[i0 + lane];
                    const unsigned m = __ballot_sync (FULL_MASK, e != EMPTY32);
                    if (m)
                    {
                        uint32_t base = 0;
                        if (lane == 0) base = atomicAdd (&s_nu, (uint32_t)__popc (m));
                        base = __shfl_sync (FULL_MASK, base, 0);
                        if (e != EMPTY32) ulist[base + __popc (m & lt_mask)] = e;
                    }
                }
            __syncthreads ();
            nu = s_nu;
        }
        else __syncthreads ();
        // ---- (3) the k-mers of the survivors go into the CTA's table ----
        if (!producer && !bin_ovf)
        {
            uint32_t rn = 0, wnew = 0;                                      // warp-uniform: pending retries, slots claimed
            bool w_fail = false;
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
                        res = k2_probe_loop (keys, cnts, k2_slot32 ((uint32_t)key, (uint32_t)(key >> 32), hshift), key, tmask, (uint32_t)retry_m[e]);
                        if (res == -1) w_fail = true;
                    }
                    wnew += __popc (__ballot_sync (FULL_MASK, e < rn && res != -1 && (res & 0x80000000)));
                }
                rn = 0;
                __syncwarp ();
            };
            for (uint32_t b0 = (uint32_t)wid * 32; b0 < nu; b0 += NT)
            {
                uint4 rec = zero4;
                if (b0 + lane < nu)
                {
                    const uint32_t e = dedup ? ulist[b0 + lane] : ((b0 + lane) | (1u << 16));
                    rec = tile[e & 0xFFFFu];
                    rec.w = (rec.w & ((1u << (DEV_FINE_SHIFT_W1 - 32)) - 1)) | ((e >> 16) << (DEV_FINE_SHIFT_W1 - 32));   // multiplicity in place of the fine-bin id
                }
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
                    const uint32_t mult = q.w >> (DEV_FINE_SHIFT_W1 - 32);
                    K2Chunk C;
                    uint32_t lo[4], hi[4], slot[4];
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
                    for (int i = 0; i < 4; i++) { slot[i] = k2_slot32 (lo[i], hi[i], hshift); cur[i] = (i < nkc) ? keys[slot[i]] : 0ULL; }
                    #pragma unroll
                    for (int i = 0; i < 4; i++)
                    {
                        const bool valid = i < nkc;
                        const unsigned long long key = ((unsigned long long)hi[i] << 32) | lo[i];
                        unsigned long long cv = cur[i];
                        const bool isE = valid && cv == EMPTY64;
                        if (isE) cv = atomicCAS (&keys[slot[i]], EMPTY64, key);
                        const bool isnew = isE && cv == EMPTY64;
                        const bool hit = valid && (isnew || cv == key);
                        if (hit && mult > (isnew ? 1u : 0u)) atomicAdd (&cnts[slot[i]], isnew ? mult - 1u : mult);     // a claim counts one by itself
                        wnew += __popc (__ballot_sync (FULL_MASK, isnew));
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
            if (rn) drain_retries ();
            if (lane == 0 && wnew) atomicAdd (&s_ndist, wnew);
            if (__any_sync (FULL_MASK, w_fail) && lane == 0) s_fail = 1;
        }
        if (flags & K2F_LAST)
        {   // ---- the bin is complete: scan (or wipe) the table ----
            __syncthreads ();
            const bool give_up = bin_ovf || s_ndist > dist_limit || s_fail != 0;
            __syncthreads ();
            if (tid == 0)
            {
                if (give_up) { const uint32_t idx = (uint32_t) atomicAdd (&P.counters[4], 1ULL); P.ovf_list[idx] = s_tbin[buf]; }
                s_ndist = 0; s_fail = 0;
            }
            if (!producer)
                for (int s0 = wid * 32; s0 < T; s0 += NT)
                {
                    const int slot = s0 + lane;
                    const unsigned long long key = keys[slot];
                    bool emit = false; uint32_t c = 0;
                    if (key != EMPTY64)
                    {
                        c = cnts[slot] + 1u;                                 // the claim itself counts one (k2_common.cuh)
                        keys[slot] = EMPTY64; cnts[slot] = 0;
                        if (!give_up)
                        {
                            n_distinct++;
                            const uint32_t hb = c >= (uint32_t)P.histo_max ? (uint32_t)P.histo_max : c;
                            // k-mers seen once are most of the distinct k-mers of a sequencing run (one per error and position):
                            // they are counted in a register, not by 32 lanes hammering one shared-memory word
                            if (hb == 1) n_once++; else if (hb < K2_HB) atomicAdd (&hist[hb], 1u); else atomicAdd (&P.histogram[hb], 1ULL);
                            if (c >= P.solid_min && c <= P.solid_max) n_solid++;
                            emit = (c >= P.emit_min && c <= P.emit_max);
                        }
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
                            if (lane == 0) b0 = atomicAdd (&P.counters[0], (unsigned long long)K2F_WBLOCK);
                            b0 = __shfl_sync (FULL_MASK, b0, 0);
                            out_pos = b0; out_end = b0 + K2F_WBLOCK;
                        }
                        if (emit)
                        {
                            const unsigned long long pos = out_pos + __popc (ballot & lt_mask);
                            n_emitted++;
                            if (pos < P.out_cap) { P.out_lo[pos] = ORI ? k2_raw_to_canonical (key, k) : key; P.out_cnt[pos] = c; }
                        }
                        out_pos += ne;
                    }
                }
            bin_ovf = false;
        }
    }
    if (!producer)
        for (unsigned long long hpos = out_pos + lane; hpos < out_end; hpos += 32)
            if (hpos < P.out_cap) P.out_lo[hpos] = EMPTY64;
    __syncthreads ();
    for (int i = tid; i < K2_HB; i += NT + 32) { const uint32_t v = hist[i]; if (v) atomicAdd (&P.histogram[i], (unsigned long long)v); }
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

size_t k2f_smem_bytes (int table_log2, int nwarp)
{
    const size_t T = (size_t)1 << table_log2;
    return (size_t)K2F_NBUF * K2F_TILE * 16 + T * 8 + (size_t)nwarp * K2F_RETRY * 8 + T * 4 + K2F_DT * 4 + K2F_TILE * 4 + K2_HB * 4
           + (size_t)nwarp * K2F_RETRY * 2 + 64;
}
int k2f_max_table_log2 () { int t = 13; while (k2f_smem_bytes (t + 1, 16) <= 227 * 1024) t++; return t; }

template<int NWARP, bool ORI>
static cudaError_t k2f_launch (const LaunchCtx& L, const K2Params& P, const K2aSrc& S, uint32_t nb, uint32_t cap, uint32_t n_bins, int dedup)
{
    const size_t smem = k2f_smem_bytes (P.table_log2, NWARP);
    cudaError_t e = cudaFuncSetAttribute (k2f_count_coarse<NWARP, ORI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor (&per_sm, k2f_count_coarse<NWARP, ORI>, NWARP * 32 + 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    uint64_t grid = (uint64_t)L.sm_count * per_sm;
    if (grid > n_bins) grid = n_bins;
    k2f_count_coarse<NWARP, ORI><<<(unsigned)grid, NWARP * 32 + 32, smem, L.stream>>> (P, S, nb, cap, n_bins, dedup);
    (*L.launches)++;
    return cudaGetLastError ();
}

// counts the n_bins coarse bins of the gathered sources; P.counters[4] / P.ovf_list receive the bins that did not fit
cudaError_t launch_k2f_count (const LaunchCtx& L, const K2Params& P, const K2aSrc& S, uint32_t nb, uint32_t cap, uint32_t n_bins, int dedup)
{
    if (n_bins == 0) return cudaSuccess;
    return P.oriented ? k2f_launch<16, true> (L, P, S, nb, cap, n_bins, dedup) : k2f_launch<16, false> (L, P, S, nb, cap, n_bins, dedup);
}
