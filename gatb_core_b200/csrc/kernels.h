// kernels.h -- parameter blocks and host-side launchers of the sm_100a kernels (definitions in k*.cu).
#pragma once
#ifndef KERNELS_H_NO_CUDA_RUNTIME      // tests/cpp/test_layout.cpp compiles the layout helpers with plain g++
#include <cuda_runtime.h>
#endif
#include <stdint.h>

// ----------------------------------------------------------------------------------------------------------------
// Record format written by k1 and consumed by k2a/k2b (replaces the reference's SuperKmerBinFiles byte stream,
// tools/storage/impl/Storage.cpp:310-589, by fixed-size HBM records so one aligned vector store/load moves one):
//   W=1 (k<=31, Kmer<32>): 16 B.  bits [0,116)  = the k+nbK-1 nucleotides of the super-k-mer, stream order, 2 bits each
//                                 bits [116,121) = nbK (1..28)           bits [121,128) = fine-bin id (7 bits)
//   W=2 (k<=63, Kmer<64>): 32 B.  bits [0,244)  = nucleotides            bits [244,250) = nbK (1..60)
//                                 bits [250,256) = fine-bin id (6 bits)
// nbK <= maxs = min((8*sizeof(Type)-8)/2, 255) = 28 / 60 exactly as Sequence2SuperKmer.hpp:147.
// ----------------------------------------------------------------------------------------------------------------
enum { REC_LEN_SHIFT_W1 = 52, REC_FINE_SHIFT_W1 = 57, FINE_BITS_W1 = 7,
       REC_LEN_SHIFT_W2 = 52, REC_FINE_SHIFT_W2 = 58, FINE_BITS_W2 = 6 };
// Device-mode records of k <= 31 trade nucleotides for fine-bin bits: at most DEV_MAXLEN_W1 = 24 k-mers (54 nucleotides,
// bits [0,108)), nbK in bits [108,113), fine id in bits [113,128) (up to 15 bits; the planner uses 9 on one GPU -- eight ids per planned counting
// bin, merged into bins of even load by k2a_dedup_split -- and one more per doubling of the ranks, so that the number of COARSE
// bins a partition kernel scatters into stays put).
// The shifts below are relative to the upper 64-bit half, like the ones above.
enum { DEV_LEN_SHIFT_W1 = 44, DEV_FINE_SHIFT_W1 = 49, DEV_MAXLEN_W1 = 24, DEV_FINE_BITS_MAX_W1 = 12 };

// ----------------------------------------------------------------------------------------------------------------
// Layout of a region of coarse bins (device mode): ROUNDS of COARSE_BLK (= 64) records.  Record 'slot' of bin 'b' of a region
// of 'nb' bins lives at ((slot / BLK) * nb + b) * BLK + slot % BLK: block r of every bin sits in round r.  All bins
// fill at about the same rate (hashed minimizers), so at any time the partition kernel writes into a few neighbouring
// rounds -- a window of tens of MB instead of one open line in every 100 KB of a 25 GB buffer.  Measured on B200: the
// scattered record stores cost 4.0 ms per 3*10^8 records in the bin-major layout and 0.9 ms inside a small window
// (address translation, not bandwidth).  Readers take a bin's blocks at stride nb*BLK records; CTAs working on
// neighbouring bins share the pages.  cap is a multiple of COARSE_BLK.
// ----------------------------------------------------------------------------------------------------------------
enum { COARSE_BLK = 64 };          // 1 KB blocks (k <= 31): the readers' address translations amortise over 64 records
__host__ __device__ __forceinline__ uint64_t coarse_index (uint32_t b, uint32_t slot, uint32_t nb)
{ return ((uint64_t)(slot / COARSE_BLK) * nb + b) * COARSE_BLK + (slot % COARSE_BLK); }

enum { K1_MODE_DEVICE = 0,   // hashed-order signature, device bins (the counting path)
       K1_MODE_GATB   = 1 }; // GATB minimizer (lexicographic + AA rule), bins = pass*nb_partitions + repart[minimizer]

struct K1Params
{
    const uint64_t* words;          // packed reads, 64-bit words
    const uint64_t* offsets;        // [n_reads+1] nucleotide offsets, or NULL with read_len
    const uint32_t* nmask;          // invalid-nucleotide bitmask or NULL
    uint64_t n_reads;               // reads handled by this launch: [first_read, first_read + n_reads)
    uint64_t first_read;
    int      read_len;
    int      k, m, w, maxlen;       // m = size of the m-mers ranked; w = k-m+1; maxlen = max k-mers per record
    uint32_t mmask, mask_ma1;
    int      mode;
    const uint16_t* repart;         // GATB mode
    int      nb_partitions, nb_passes;
    uint32_t nb1;                   // number of (coarse) bins
    uint32_t n_regions, bins_per_region;   // device mode: nb1 = n_regions * bins_per_region; region r (one per owner rank) is a
                                    // contiguous block of bins_per_region*cap records laid out by coarse_index()
    const uint64_t* bin_off;        // NULL: bins of fixed capacity 'cap' laid out by coarse_index; else DENSE layout: bin b owns the
                                    // records [bin_off[b], bin_off[b+1]) (exact sizes from a counting run: skewed inputs)
    uint32_t cap;                   // records per bin
    int      fine_bits;
    void*    bins;                  // nb1*cap records
    uint32_t* cursors;              // [nb1] demand per bin (may exceed cap)
    unsigned long long* stats;      // [0] valid k-mers [1] invalid k-mers [2] records stored [3] records dropped (overflow)
    int      count_only;            // 1: only count demand (cursors), store nothing
    int      force_general;         // 1: never take the register-scanner kernel (GATB_PATH_K1_GENERAL, reads of 2^20 nucleotides and more)
    int      oriented;              // 1: records oriented by the strand of their minimizer (k1_scan.cuh); see k1_oriented()
    int      max_len;               // longest read of the batch when offsets are given (0 = unknown: no shared-memory staging)
    int      no_staging;            // 1 (default): the register scanner reads the stream from global memory; 0: TMA-staged tiles (GATB_PATH_K1_STAGING)
};

struct K2Params
{
    int      k, W;
    const void* recs;               // fine-split records
    const uint2* bin_desc;          // [nbins] {offset inside coarse region, record count}
    uint32_t nbins;                 // nb1 << fine_bits
    const uint64_t* coarse_off;     // [nb1+1] first record of each coarse bin in the fine-split copy
    int fine_bits;
    int      table_log2;            // log2 slots of the shared-memory table
    int      path_flags;            // gatb_gpu_params.path_flags
    int      oriented;              // the records are oriented: keys are plain slices of the records (k2b_warp_bins<NT,true>)
    uint32_t emit_min, emit_max;    // emit k-mers with emit_min <= count <= emit_max
    uint32_t solid_min, solid_max;  // solidity range (stats only)
    int      histo_max;
    unsigned long long* histogram;  // [histo_max+1]
    uint64_t* out_lo; uint64_t* out_hi; uint32_t* out_cnt; unsigned long long out_cap;
    unsigned long long* counters;   // [0] output cursor (incl. holes) [1] distinct [2] solid [3] work counter [4] overflow bins [5] k-mers in overflow bins [6] emitted [7] overflow bins of the second tier
    uint32_t* ovf_list;             // [nbins] ids of bins whose table overflowed
    const uint32_t* bin_list; uint32_t n_list;   // second-tier run (k2b_count_w1 only): count these bins instead of [0, nbins)
    int      ovf_counter;           // index into counters[] of the overflow count this launch appends to (4, or 7 for the second tier)
    // global-memory fallback table
    uint64_t* g_lo; uint64_t* g_hi; uint32_t* g_cnt; int g_log2;
};

struct K3Params
{
    int k, m, W;
    uint32_t mmask, mask_ma1;
    const uint16_t* repart; int nb_partitions, nb_passes; uint32_t n_keys;
    int t_bits;                     // value-range bits per key: bucket = key << t_bits | top t_bits of the k-mer
    uint64_t n;
    const uint64_t* in_lo; const uint64_t* in_hi; const uint32_t* in_cnt;
    const uint16_t* in_key;         // NULL: the partition key is computed from the k-mer; else it came with the item (routed items, k3r_*)
    uint32_t* bucket_of;            // [n]
    uint32_t* bucket_count;         // [n_buckets]   -> cursors during scatter
    const uint64_t* bucket_off;     // [n_buckets+1]
    uint64_t* tmp_lo; uint64_t* tmp_hi; uint32_t* tmp_cnt;     // scattered
    uint64_t* out_lo; uint64_t* out_hi; int32_t* out_cnt;      // sorted
    uint32_t n_buckets;
    uint32_t bucket_begin, bucket_end;   // k3c sorts buckets [bucket_begin, bucket_end) (chunked so that copies overlap)
    unsigned long long* big_list;   // buckets too large for shared memory
    unsigned long long* counters;   // [0] number of big buckets
    // pooled scatter (k3s): bucket b keeps its items in blocks of K3_BLK items taken from one bump allocator; block j of
    // bucket b is pool block dir[j * n_buckets + b].  NULL pool = the exact two-pass layout (tmp_*) is used instead.
    uint4*    pool; uint32_t pool_blocks; uint32_t* pool_ptr; uint32_t* dir; uint32_t dir_rounds; uint32_t* ovf_flag;
};
enum { K3_BLK = 32 };

struct LaunchCtx { cudaStream_t stream; int sm_count; uint64_t* launches; };

// k1_partition.cu
cudaError_t launch_k1 (const LaunchCtx&, const K1Params&);
int         k1_fast_window (int k);      // window (k-m+1) the register-scanner kernel is compiled for, 0 = none
bool        k1_oriented (int k, int m, int w, int path_flags);   // does launch_k1 write oriented records for these parameters?
// k2_count.cu
// the same coarse bins gathered from n sources.  off[s] == NULL: source s is laid out by coarse_index (fixed capacity per bin);
// off[s] != NULL: DENSE layout, bin b of source s holds its records at [off[s][b], off[s][b+1]) (exact sizes: skewed inputs)
#define K2A_MAXSRC 32           // = GATB_GPU_MAX_SOURCES (ranks x pieces per rank)
struct K2aSrc { const uint4* bins[K2A_MAXSRC]; const uint32_t* cursors[K2A_MAXSRC]; const uint64_t* off[K2A_MAXSRC]; int n; };
__host__ __device__ __forceinline__ uint64_t k2a_record_index (const K2aSrc& S, int s, uint32_t b, uint32_t slot, uint32_t nb)
{ return S.off[s] ? S.off[s][b] + slot : coarse_index (b, slot, nb); }
// pre-split of gathered bins by the leading bits of the fine id (k2a_fine_split): first record and record count of every sub-bin
struct K2aPresplit { uint64_t* sub_off; uint32_t* sub_cnt; int shift; };
cudaError_t launch_k2a_split (const LaunchCtx&, int W, const K2aSrc& src, void* dst, const uint64_t* coarse_off,
                              uint32_t nb1, uint32_t cap, int fine_bits, uint2* bin_desc, const uint32_t* bin_list = 0, uint32_t n_list = 0,
                              int desc_abs = 0, uint64_t desc_base = 0, const K2aPresplit* presplit = 0);
uint32_t    k2a_dedup_rmax (uint32_t max_bin_records, uint32_t mean_bin_records, int fine_bits);
uint32_t    k2a_two_cta_capacity (int fine_bits);
cudaError_t launch_k2a_dedup_split (const LaunchCtx&, const K2aSrc& src, void* dst, const uint64_t* coarse_off, uint32_t nb1, uint32_t cap,
                                    int fine_bits, uint2* bin_desc, uint32_t rmax, uint32_t* big_list, unsigned long long* counters,
                                    uint32_t target /* k-mers of surviving records per counting bin; 0: one bin per fine id */,
                                    uint32_t big_load /* bins with more k-mers are flagged K2_DESC_BIG; 0: none */);
enum : uint32_t { K2_DESC_BIG = 0x80000000u, K2_DESC_COUNT = 0x7FFFFFFFu };      // bin descriptor .y = record count | flag
cudaError_t launch_k2b_count (const LaunchCtx&, const K2Params&);
cudaError_t launch_k2b_count_list (const LaunchCtx&, const K2Params&);     // CTA-per-bin kernel over P.bin_list (k <= 31)
int         k2b_variant (int path_flags);            // 1 warp per bin, 128 / 256 CTA per bin (chunked insert), 0 one k-mer per lane
int         k2b_default_table_log2 (int W, int path_flags);
cudaError_t launch_k2c_measure (const LaunchCtx&, const K2Params&, uint32_t n_ovf);
cudaError_t launch_k2c_insert (const LaunchCtx&, const K2Params&, uint32_t n_ovf);
cudaError_t launch_k2c_scan (const LaunchCtx&, const K2Params&);
// k3_sort.cu
cudaError_t launch_k3a_classify (const LaunchCtx&, const K3Params&);
cudaError_t launch_k3b_scatter (const LaunchCtx&, const K3Params&);
cudaError_t launch_k3r_route (const LaunchCtx&, const K3Params&, uint32_t n_ranks, uint64_t dest_cap, unsigned long long* dest_cursor,
                              uint64_t* o_lo, uint64_t* o_hi, uint32_t* o_cnt, uint16_t* o_key, uint32_t* ovf_flag);
cudaError_t launch_k3s_pool_scatter (const LaunchCtx&, const K3Params&);
uint32_t    k3_sort_cap ();
cudaError_t launch_k3s_pool_scatter (const LaunchCtx&, const K3Params&);
uint32_t    k3_sort_cap ();
cudaError_t launch_k3b_scatter_coarse (const LaunchCtx&, const K3Params&, int shift, uint32_t* group_cursor,
                                       uint64_t* g_lo, uint64_t* g_hi, uint32_t* g_cnt, uint32_t* g_bucket);
cudaError_t launch_k3c_sort (const LaunchCtx&, const K3Params&);
cudaError_t launch_k3d_sort_big (const LaunchCtx&, const K3Params&, uint32_t n_big);
cudaError_t launch_scan_u32_to_u64 (const LaunchCtx&, const uint32_t* in, uint64_t* out, uint64_t n, uint64_t* scratch);
uint64_t    scan_scratch_elems (uint64_t n);
// gatb_serialize (k1 GATB mode -> reference byte streams)
cudaError_t launch_serialize_sizes (const LaunchCtx&, int W, int k, const void* bins, const uint32_t* cursors, uint32_t nb1,
                                    uint32_t cap, unsigned long long* key_bytes);
cudaError_t launch_serialize_write (const LaunchCtx&, int W, int k, const void* bins, const uint32_t* cursors, uint32_t nb1,
                                    uint32_t cap, const uint64_t* key_off, unsigned long long* key_cur, uint8_t* out);
// bloom.cu
cudaError_t launch_bloom_insert (const LaunchCtx&, int kind, int W, int k, int nb_hash, uint64_t tai, int pow2, uint64_t reduced,
                                 const uint64_t* lo, const uint64_t* hi, uint64_t n, uint32_t* words);
// synth.cu
cudaError_t launch_synth_reads (const LaunchCtx&, uint64_t seed, uint64_t genome_len, uint64_t first_read, uint64_t n_reads,
                                int L, uint8_t* packed);
cudaError_t launch_pack_ascii (const LaunchCtx&, const char* ascii, uint64_t n, uint32_t* packed_words, uint32_t* nmask,
                               unsigned long long* n_invalid);
cudaError_t launch_pack_ascii_at (const LaunchCtx&, const char* ascii, uint64_t n, uint64_t base, uint32_t* packed_words, uint32_t* nmask,
                                  unsigned long long* n_invalid);
cudaError_t launch_rebase_offsets (const LaunchCtx&, const uint64_t* in, uint64_t n, uint64_t base, uint64_t* out);
// k2_fused.cu
size_t      k2f_smem_bytes (int table_log2, int nwarp);
cudaError_t launch_k2f_count (const LaunchCtx&, const K2Params&, const K2aSrc&, uint32_t nb, uint32_t cap, uint32_t n_bins, int dedup);
// k_parse.cu (FASTA / FASTQ text on the device)
uint64_t    text_blocks (uint64_t n);
cudaError_t launch_text_count_newlines (const LaunchCtx&, const char* text, uint64_t n, uint32_t* block_counts);
cudaError_t launch_text_line_starts (const LaunchCtx&, const char* text, uint64_t n, const uint64_t* block_off, uint64_t* line_start);
cudaError_t launch_text_lines (const LaunchCtx&, const char* text, uint64_t n, const uint64_t* line_start, uint64_t n_lines, int format,
                               uint32_t* seq_len, uint32_t* is_header);
cudaError_t launch_text_offsets (const LaunchCtx&, const uint32_t* is_header, const uint64_t* line_dst, const uint64_t* line_rec, uint64_t n_lines,
                                 uint64_t base, uint64_t* out, uint64_t n_recs, uint64_t total_nt);
cudaError_t launch_text_pack (const LaunchCtx&, const char* text, const uint64_t* line_start, const uint64_t* line_dst, uint64_t n_lines,
                              uint64_t total_nt, uint64_t base, uint32_t* words, uint32_t* nmask, unsigned long long* n_invalid);
cudaError_t launch_text_stats (const LaunchCtx&, const uint64_t* offsets, uint64_t n, unsigned long long* out);
cudaError_t launch_synth_reads_zipf (const LaunchCtx&, uint64_t seed, uint64_t n_species, const uint64_t* d_cdf, const uint64_t* d_goff,
                                     uint64_t first_read, uint64_t n_reads, int len, uint8_t* packed);
// k_repart.cu
cudaError_t launch_repart_sample (const LaunchCtx&, const uint64_t* words, const uint64_t* offsets, int read_len, const uint32_t* nmask,
                                  uint64_t n_reads, int k, int m, uint32_t* sk_count, unsigned long long* kx_table);
