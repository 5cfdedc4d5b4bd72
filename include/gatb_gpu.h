/* gatb_gpu.h -- C ABI of the B200-native k-mer counting path (libgatb_b200.so).
 *
 * This is the drop-in boundary for ONE hot path of GATB-core: DSK / SortingCountAlgorithm (+ Bloom insertion of the
 * solid k-mers).  Plain pointers and sizes only; no C++ or torch types.  The C++ host shim that keeps GATB's own
 * API (gatb_core_b200/host/: Kmer<span>::Model*, SortingCountAlgorithm<span>, ICountProcessor<span>) and the Python
 * harness (gatb_core_b200/__init__.py, ctypes) are both thin callers of these entry points.
 *
 * Reference interfaces replaced (paths relative to /root/reference/gatb-core/src/gatb/):
 *   gatb_gpu_count / _dev     SortingCountAlgorithm<span>::execute            kmer/impl/SortingCountAlgorithm.cpp:636-781
 *                             = fillPartitions (:1211-1344) + fillSolidKmers (:1384-1602) + the default
 *                             ICountProcessor chain histogram -> solidity -> dump (kmer/api/ICountProcessor.hpp:91-183,
 *                             kmer/impl/CountProcessorHistogram.hpp:173, CountProcessorSolidity.hpp:186, CountProcessorDump.hpp:148)
 *   gatb_gpu_superkmers       Sequence2SuperKmer<span>::operator()            kmer/impl/Sequence2SuperKmer.hpp:138-159
 *                             + FillPartitions<span,true>::processSuperkmer   kmer/impl/SortingCountAlgorithm.cpp:1081-1151
 *                             + Kmer<span>::SuperKmer::save                   kmer/impl/Model.hpp:1386-1471
 *   gatb_gpu_bloom_params     BloomAlgorithm<span>::execute sizing            kmer/impl/BloomAlgorithm.cpp:158-166
 *   gatb_gpu_bloom / _dev     BloomBuilder<span>::build -> Bloom*::insert     kmer/impl/BloomBuilder.hpp:102-131,
 *                                                                             tools/collections/impl/Bloom.hpp:394-412,445-459,555-588
 *   gatb_gpu_histogram_cutoff Histogram::compute_threshold                    tools/misc/impl/Histogram.cpp:61-190
 *
 * Conventions
 *   - nucleotides: A=0 C=1 T=2 G=3 (tools/misc/api/Data.hpp:185), a k-mer value holds its FIRST nucleotide in the most
 *     significant position (kmer/impl/Model.hpp:636-657); canonical = min(forward, reverse complement).
 *   - packed reads: 2 bits per nucleotide, nucleotide i of the stream in bits [2(i%4), 2(i%4)+2) of byte i/4; read r
 *     occupies stream positions [read_offsets_nt[r], read_offsets_nt[r+1]).  Buffers must be 16-byte aligned and
 *     readable for 32 bytes past the last nucleotide (gatb_gpu_count copies host input into such a buffer itself).
 *   - n_mask (optional, may be NULL): 1 bit per stream position, bit (i%32) of 32-bit word i/32, set = the
 *     nucleotide is not A/C/G/T (k-mers overlapping it are invalid and dropped, Sequence2SuperKmer.hpp:95-108).
 *   - k-mers are returned as (lo, hi) 64-bit halves; hi arrays are NULL when kmer_size < 32 (Kmer<32>).
 *   - every function returns 0 on success, non-zero on error; gatb_gpu_last_error() describes the last failure.
 *     There is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef GATB_GPU_H
#define GATB_GPU_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gatb_gpu_ctx gatb_gpu_ctx;

/* Mirrors the fields of kmer/impl/Configuration.hpp:56-100 that decide the OUTPUT. */
typedef struct gatb_gpu_params
{
    int32_t  kmer_size;          /* _kmerSize        1 <= m < k <= 63                                          */
    int32_t  minimizer_size;     /* _minim_size      m <= 12 (size of the Repartitor table is 4^m)             */
    int32_t  nb_partitions;      /* _nb_partitions                                                              */
    int32_t  nb_passes;          /* _nb_passes       pass = minimizer % nb_passes (SortingCountAlgorithm.cpp:1083)*/
    int32_t  abundance_min;      /* _abundance[0].getBegin()  (solid <=> min <= count <= max)                   */
    int32_t  abundance_max;      /* _abundance[0].getEnd()    (default 2^31-1)                                  */
    int32_t  histo_max;          /* _abundanceUserNb / -histo-max (default 10000)                               */
    int32_t  minimizer_type;     /* _minimizerType   0 = lexicographic (supported); 1 = frequency (not yet)     */
    int32_t  emit_all;           /* 1: return EVERY distinct k-mer (for custom ICountProcessor chains); 0: solid only */
    int32_t  read_len;           /* >0: all reads have this length and read_offsets_nt may be NULL               */
    int32_t  table_log2;         /* 0 = default (9 for kmer_size < 32: one warp per fine bin; 11 otherwise; 13 with GATB_PATH_FUSED);
                                    log2 slots of the first-tier shared-memory table, 5..13                        */
    int32_t  path_flags;         /* 0 = the product path.  Selectors of the alternate code paths (GATB_PATH_*), so that one
                                    process can run the parity suite through every kernel variant                    */
    int32_t  k3_dir_rounds;      /* 0 = default; >0: rounds of the block directory of the bucket scatter (a tiny one forces the
                                    exact two-pass fallback)                                                         */
    int32_t  bin_load_pct;       /* 0 = default; planned k-mer occurrences per fine bin in % of the first-tier table slots */
    int32_t  fine_bits;          /* 0 = default (9 on one GPU); k <= 31: log2 of the fine ids per coarse bin (experiments)    */
    int32_t  bin_target_pct;     /* 0 = default (45); k <= 31: k-mers of distinct records per counting bin in % of the table slots:
                                    the fine split merges consecutive fine ids into counting bins of that load                   */
} gatb_gpu_params;
/* gatb_gpu_params.path_flags */
enum {
    GATB_PATH_K1_GENERAL   = 1,      /* general partition kernel instead of the register scanner (records not oriented)   */
    GATB_PATH_K2B_MASK     = 6,      /* k <= 31 counting kernel: 0 warp per bin (default), 2 CTA per bin 128 threads,      */
    GATB_PATH_K2B_CTA128   = 2,      /*   4 CTA per bin 256 threads, 6 CTA per bin one k-mer per lane                      */
    GATB_PATH_K2B_CTA256   = 4,
    GATB_PATH_K2B_LANE     = 6,
    GATB_PATH_K2B_W2_WARP  = 8,      /* 32 <= k <= 63: warp-per-bin kernel instead of CTA per bin                          */
    GATB_PATH_NO_TIER2     = 16,     /* overflowing bins go straight to the global-memory table                            */
    GATB_PATH_K3_NO_POOL   = 32,     /* exact two-pass bucket scatter instead of the pooled single pass                    */
    GATB_PATH_CANONICAL    = 64,     /* register scanner without orientation: k2b rebuilds min(forward, revcomp) per k-mer  */
    GATB_PATH_NO_DEDUP     = 128,    /* identical records are not collapsed (every multiplicity is 1)                       */
    GATB_PATH_K1_STAGING   = 512,    /* the register scanner reads TMA-staged tiles (one bulk copy per 32 reads) instead of the global
                                        stream: measured 5 ms slower and the same DRAM traffic, kept as a tested variant */
    GATB_PATH_K2A_SMALL_STAGE = 1024,/* test selector: the dedup split stages at most 256 records, so that ordinary bins take its
                                        several-passes path and the largest ones its two-pass fallback                                  */
    GATB_PATH_K2A_PRESPLIT = 2048,   /* test selector: the pre-split of gathered bins (several ranks) on a single source too            */
    GATB_PATH_FUSED        = 256     /* k <= 31: one CTA counts a whole coarse bin straight out of the partition buffers (k2_fused.cu:
                                        TMA-streamed tiles, CTA-wide table) instead of fine split + warp-per-fine-bin counting  */
};

enum { GATB_GPU_NSTATS = 16, GATB_GPU_MAX_RANKS = 8, GATB_GPU_MAX_SOURCES = 32 };   /* sources = ranks x pieces per rank */
/* indices into gatb_gpu_result.stats */
enum {
    GATB_STAT_KMERS_VALID = 0,   /* kmers_nb_valid   (SortingCountAlgorithm.cpp:737)   */
    GATB_STAT_KMERS_INVALID = 1, /* kmers_nb_invalid                                   */
    GATB_STAT_DISTINCT = 2,      /* kmers_nb_distinct (CountProcessorSolidity.hpp:163) */
    GATB_STAT_SOLID = 3,         /* kmers_nb_solid                                     */
    GATB_STAT_RECORDS = 4,       /* super-k-mer records written by the partition kernel */
    GATB_STAT_SEQUENCES = 5,
    GATB_STAT_NUCLEOTIDES = 6,
    GATB_STAT_BINS = 7,          /* device bins used                                    */
    GATB_STAT_OVERFLOW_BINS = 8, /* bins the first-tier (per-warp) table could not hold; stats[12] = those that also
                                    overflowed the second-tier (per-CTA, 8192 slots) table and went to the global table,
                                    stats[11] = k-mer occurrences in the latter */
    GATB_STAT_RETRIES = 9,       /* partition-kernel re-runs after a bucket overflow    */
    GATB_STAT_RECORD_BYTES = 10, /* bytes of super-k-mer records (S of SURVEY.md 8d)    */
    GATB_STAT_UNIQUE_RECORDS = 13 /* records left after identical ones were collapsed (k <= 31) */
};

typedef struct gatb_gpu_result
{
    uint64_t  n_keys;            /* nb_passes * nb_partitions; key = pass*nb_partitions + partition              */
    uint64_t  n_items;           /* total k-mers returned                                                         */
    uint64_t* part_offsets;      /* [n_keys+1] offsets into the arrays below; ascending k-mer order inside a key  */
    uint64_t* kmers_lo;          /* [n_items]                                                                     */
    uint64_t* kmers_hi;          /* [n_items] or NULL                                                             */
    int32_t*  counts;            /* [n_items] CountNumber (system/api/types.hpp:49)                               */
    uint64_t* histogram;         /* [histo_max+1], index clamped like Histogram::inc (tools/misc/impl/Histogram.hpp:92) */
    uint64_t  stats[GATB_GPU_NSTATS];
    double    seconds[8];        /* stream time per stage (CUDA events): 0 h2d, 1 partition, 2 split, 3 count, 4 sort, 5 d2h,
                                    6 device total, 7 end to end */
    double    kernel_seconds[8]; /* kernel-only durations: 0 partition (k1), 1 fine split (k2a), 2 first-tier count (k2b),
                                    3 k3 (classify + scatter + scan + sort), 4 overflow tiers + global fallback;
                                    diagnostics of the sort stage, not times: 5 buckets sorted in global memory, 6 = 1 when the exact two-pass
                                    scatter replaced the pooled one, 7 value-range bits per key */
    int32_t   on_device;         /* 1: the arrays above are DEVICE pointers (gatb_gpu_count_dev), 0: host            */
    int32_t   pad;
    void*     owner;             /* internal */
} gatb_gpu_result;

/* ---- context ------------------------------------------------------------------------------------------------ */
gatb_gpu_ctx* gatb_gpu_create  (int device);         /* NULL when the device cannot be opened                     */
void          gatb_gpu_destroy (gatb_gpu_ctx*);
const char*   gatb_gpu_last_error (gatb_gpu_ctx*);   /* ctx may be NULL: error of gatb_gpu_create                 */
void*         gatb_gpu_stream (gatb_gpu_ctx*);       /* the cudaStream_t every kernel of this ctx is launched on  */
uint64_t      gatb_gpu_kernel_launches (gatb_gpu_ctx*); /* kernels launched by this ctx so far                    */
int           gatb_gpu_sm_count (gatb_gpu_ctx*);

/* ---- DSK: reads -> sorted (k-mer, count) per partition + histogram ------------------------------------------- */
/* HOST buffers in, HOST arrays out (copies are part of the call).  The returned host arrays live in a pinned staging
 * buffer owned by the context: they stay valid until the next gatb_gpu_count on this context or gatb_gpu_result_free. */
int gatb_gpu_count (gatb_gpu_ctx*, const gatb_gpu_params*, const uint16_t* repart_table /* [4^m] or NULL when n_keys==1 */,
                    const uint32_t* freq_order /* NULL */, const uint8_t* packed_reads, const uint64_t* read_offsets_nt,
                    uint64_t n_reads, const uint32_t* n_mask, gatb_gpu_result* out);
/* DEVICE buffers in, DEVICE arrays out (owned by the context, valid until its next count call). repart_table is a HOST pointer. */
int gatb_gpu_count_dev (gatb_gpu_ctx*, const gatb_gpu_params*, const uint16_t* repart_table, const uint32_t* freq_order,
                        const uint8_t* d_packed_reads, const uint64_t* d_read_offsets_nt, uint64_t n_reads,
                        const uint32_t* d_n_mask, gatb_gpu_result* out);
void gatb_gpu_result_free (gatb_gpu_ctx*, gatb_gpu_result*);

/* ---- streaming input (SURVEY.md 8b "push_reads / finish"): the bank is pushed batch by batch as ASCII, packed to 2 bits on
 * the device as it arrives (row A1; replaces the per-sequence Data::ConvertASCII of bank::Sequence, tools/misc/api/Data.hpp:185,
 * inside the IteratorCommand loop tools/designpattern/impl/ICommand.hpp:304-330), and counted once at the end.  The host never
 * holds more than one batch.
 *   gatb_gpu_reads_begin        forgets the reads pushed so far (expected_nt: capacity hint, may be 0)
 *   gatb_gpu_reads_push_ascii   n_seqs sequences concatenated without separators in ascii[seq_offsets[0] .. seq_offsets[n_seqs]);
 *                               characters other than ACGTacgt are invalid nucleotides (their k-mers are dropped)
 *   gatb_gpu_reads_count        gatb_gpu_count over everything pushed: host arrays out, same ownership rules
 *   gatb_gpu_reads_push_text    a batch of FASTA (multi-line records) or FASTQ (four-line records) TEXT, cut by the caller at record
 *                               boundaries; lines, records, offsets and the packing are found on the device (k_parse.cu) -- the
 *                               device-side replacement of the reference's line parser bank/impl/BankFasta.cpp:391-620
 *   gatb_gpu_reads_info         what was pushed so far: [0] sequences [1] nucleotides [2] shortest [3] longest record
 *                               [4] sum of squared lengths (bits of a double) [5] invalid nucleotides (BankStats, kmer/impl/BankKmers.hpp:164-215) */
enum { GATB_TEXT_FASTA = 0, GATB_TEXT_FASTQ = 1 };
int gatb_gpu_reads_begin (gatb_gpu_ctx*, uint64_t expected_nt);
int gatb_gpu_reads_push_text (gatb_gpu_ctx*, const char* text, uint64_t n_bytes, int format);
int gatb_gpu_reads_info (gatb_gpu_ctx*, uint64_t* info6);
int gatb_gpu_reads_push_ascii (gatb_gpu_ctx*, const char* ascii, const uint64_t* seq_offsets, uint64_t n_seqs);
int gatb_gpu_reads_count (gatb_gpu_ctx*, const gatb_gpu_params*, const uint16_t* repart_table, const uint32_t* freq_order,
                          gatb_gpu_result* out);

/* ---- the same path in stages, for multi-GPU runs (one process per GPU) -------------------------------------------
 * The device binning: every rank partitions ITS reads into nb1 coarse bins with the SAME geometry;
 * coarse bin b is owned by rank b / bins_per_rank; the caller moves each bin region to its owner (one all-to-all of
 * [bins_per_rank][cap] records + cursors, or only the used rounds of it) and the owner counts the bins it gathered
 * from all sources.
 * Replaces the reference's only exchange medium, the SuperKmerBinFiles temp files (tools/storage/impl/Storage.cpp:310-347). */
typedef struct gatb_gpu_geometry
{
    uint64_t total_kmers;       /* k-mer positions of the WHOLE job (all ranks)                                          */
    uint32_t nb1;               /* coarse bins (multiple of n_ranks)                                                     */
    uint32_t cap;               /* record slots per coarse bin in a rank's partition buffer                              */
    int32_t  fine_bits;         /* fine bins per coarse bin = 1 << fine_bits                                             */
    int32_t  table_log2;
    int32_t  m_device, w, maxlen, words;
    uint32_t n_ranks, bins_per_rank, record_bytes;
    uint32_t coarse_blk;        /* records per block of the round-interleaved region layout; cap is a multiple of it, and a region
                                   that holds at most c records per bin only uses its first ceil(c/coarse_blk) rounds:
                                   ceil(c/coarse_blk) * bins_per_rank * coarse_blk * record_bytes bytes                          */
} gatb_gpu_geometry;
int gatb_gpu_plan (gatb_gpu_ctx*, const gatb_gpu_params*, uint64_t total_kmers, uint64_t n_reads, int n_ranks, gatb_gpu_geometry* out);
/* k1 into caller buffers: d_bins [nb1*cap records], d_cursors [nb1] (demand; > cap means overflow);
 * stats4 (host): valid k-mers, invalid k-mers, records stored, records dropped. */
int gatb_gpu_partition_into (gatb_gpu_ctx*, const gatb_gpu_params*, const gatb_gpu_geometry*,
                             const uint8_t* d_packed_reads, const uint64_t* d_read_offsets_nt, uint64_t n_reads, const uint32_t* d_n_mask,
                             void* d_bins, uint32_t* d_cursors, uint64_t* stats4);
/* the same for the reads [first_read, first_read + n_reads) of the batch only: a rank that partitions its reads in
 * several pieces (each into its own buffers) can send piece i while piece i+1 is being partitioned; the owner hands all
 * pieces of all ranks to gatb_gpu_count_bins as separate sources (n_src <= GATB_GPU_MAX_SOURCES). */
int gatb_gpu_partition_range_into (gatb_gpu_ctx*, const gatb_gpu_params*, const gatb_gpu_geometry*,
                             const uint8_t* d_packed_reads, const uint64_t* d_read_offsets_nt, uint64_t first_read, uint64_t n_reads, const uint32_t* d_n_mask,
                             void* d_bins, uint32_t* d_cursors, uint64_t* stats4);
/* counts nb1_local coarse bins gathered from n_src sources: d_src_bins[s] = [nb1_local*cap records], d_src_cursors[s] =
 * [nb1_local]; kmers_bound >= k-mers in these bins (the fine split counts its own bins, nothing else is exchanged). */
int gatb_gpu_count_bins (gatb_gpu_ctx*, const gatb_gpu_params*, const gatb_gpu_geometry*, int n_src,
                         const void* const* d_src_bins, const uint32_t* const* d_src_cursors,
                         uint32_t nb1_local, const uint16_t* repart_table, uint64_t kmers_bound, gatb_gpu_result* out);

/* Several devices driven by ONE process (no collective library: one host thread per device, peer copies of the bin regions): the whole
 * staged sequence above for the reads of one batch, split evenly over the contexts (one per device, from gatb_gpu_create), and the
 * per-device ascending runs of every partition merged on the host -- the result is what gatb_gpu_count returns (HOST arrays, owned by
 * ctxs[0], valid until its next call).  For C / C++ callers such as SortingCountAlgorithm<span>::execute (); the multi-process path
 * (one process per GPU, NCCL) is gatb_core_b200/multigpu.py. */
int gatb_gpu_count_multi (gatb_gpu_ctx* const* ctxs, int n_dev, const gatb_gpu_params*, const uint16_t* repart_table,
                          const uint8_t* packed_reads, const uint64_t* read_offsets_nt, uint64_t n_reads, const uint32_t* n_mask,
                          gatb_gpu_result* out);

/* Second exchange of a multi-GPU run: the result of a partition must be ONE ascending sequence (what ICountProcessor::process sees in
 * the reference, kmer/impl/PartitionsCommand.cpp:1599-1805), but a k-mer's device bin -- hence the rank that counted it -- is unrelated
 * to its GATB partition.  gatb_gpu_count_bins_routed counts like gatb_gpu_count_bins and, instead of sorting, groups the emitted
 * k-mers by the rank that owns their partition key (key % n_ranks): out->kmers_lo / kmers_hi / counts and *d_keys (16-bit keys) are
 * DEVICE arrays of n_ranks regions of send_counts[n_ranks] items each, the first send_counts[r] items of region r going to rank r
 * (send_counts: host [n_ranks + 1]; a key travels as key / n_ranks, its index among the keys of its owner); out->n_items is their total, out->part_offsets is NULL, out->histogram the device histogram of
 * this rank's bins.  The caller exchanges the groups (all-to-all) and hands what it received to gatb_gpu_sort_routed: ascending order
 * of the partitions this rank owns, result as gatb_gpu_count_bins (every key not owned is empty; the keys travel with the items and are
 * not computed twice). */
int gatb_gpu_count_bins_routed (gatb_gpu_ctx*, const gatb_gpu_params*, const gatb_gpu_geometry*, int n_src,
                                const void* const* d_src_bins, const uint32_t* const* d_src_cursors,
                                uint32_t nb1_local, const uint16_t* repart_table, uint64_t kmers_bound, int n_ranks,
                                uint64_t* send_counts, uint16_t** d_keys, gatb_gpu_result* out);
int gatb_gpu_sort_routed (gatb_gpu_ctx*, const gatb_gpu_params*, const uint64_t* d_kmers_lo, const uint64_t* d_kmers_hi,
                          const uint32_t* d_counts, const uint16_t* d_keys, uint64_t n_items, int n_ranks, int rank, gatb_gpu_result* out);

/* ---- GATB-exact super-k-mer partitioning (rows A3-A6): per key, the record stream [u8 nbK][packed bytes]... that
 *      the reference writes to its SuperKmerBinFiles (order of records inside a key is unspecified).
 *      streams[key] is malloc'ed host memory (gatb_gpu_free_host); stats_out: [0] nb super-k-mers [1] nb k-mers
 *      [2] valid k-mers [3] invalid k-mers. ---- */
int gatb_gpu_superkmers (gatb_gpu_ctx*, const gatb_gpu_params*, const uint16_t* repart_table,
                         const uint8_t* packed_reads, const uint64_t* read_offsets_nt, uint64_t n_reads,
                         const uint32_t* n_mask, uint8_t** streams, uint64_t* stream_sizes, uint64_t* stats_out);
void gatb_gpu_free_host (void*);

/* ---- Repartitor table: replaces RepartitorAlgorithm<span>::computeRepartition (kmer/impl/RepartitionAlgorithm.cpp:394-492;
 *      minimizer_type 0, one bank, k <= 63).  The serial sampling pass over the first reads of the bank (SampleRepart :157-243: super-k-mers
 *      in GATB's minimizer order, kx-mers charged to their minimizer, cancelled after nb_seqs_to_see super-k-mers -- the reference uses
 *      max (5 % of the estimated number of reads, 10^6), :451) runs on the device, thread <-> read; the distribution (largest minimizer bin
 *      into the emptiest partition, Repartitor::computeDistrib kmer/impl/PartiInfo.cpp:48-106) is host arithmetic on the 4^m counters.
 *      HOST buffers in; table_out: u16[4^m], the table gatb_gpu_count takes; info3 (may be NULL): reads sampled, super-k-mers seen,
 *      kx-mers charged. ---- */
int gatb_gpu_repartition (gatb_gpu_ctx*, const gatb_gpu_params*, const uint8_t* packed_reads, const uint64_t* read_offsets_nt,
                          uint64_t n_reads, const uint32_t* n_mask, uint64_t nb_seqs_to_see, uint16_t* table_out, uint64_t* info3);

/* ---- Bloom filter of solid k-mers ---------------------------------------------------------------------------- */
enum { GATB_BLOOM_BASIC = 0, GATB_BLOOM_CACHE = 1, GATB_BLOOM_NEIGHBOR = 2 };
/* bloom_size = (u64)((float)nb_solid * (float)rvalues[k][1]); nb_hash = floorf(0.7 * bits) */
int gatb_gpu_bloom_params (int kmer_size, uint64_t nb_solid, uint64_t* bloom_size, int32_t* nb_hash);
/* byte size of the array (1 + tai/8) and the value the reference reports as getBitSize() */
int gatb_gpu_bloom_layout (int kind, uint64_t bloom_size, uint64_t* nbytes, uint64_t* bit_size);
/* host k-mers in, host bytes out (out_bytes has nbytes from gatb_gpu_bloom_layout) */
int gatb_gpu_bloom (gatb_gpu_ctx*, int kind, uint64_t bloom_size, int nb_hash, int kmer_size,
                    const uint64_t* kmers_lo, const uint64_t* kmers_hi, uint64_t n, uint8_t* out_bytes);
/* device k-mers in, device bytes out (d_out_bytes zeroed by the call; must be padded to a multiple of 4 bytes) */
int gatb_gpu_bloom_dev (gatb_gpu_ctx*, int kind, uint64_t bloom_size, int nb_hash, int kmer_size,
                        const uint64_t* d_kmers_lo, const uint64_t* d_kmers_hi, uint64_t n, uint8_t* d_out_bytes);

/* ---- Histogram cutoff (host arithmetic on the small histogram; doubles like the reference) -------------------- */
int gatb_gpu_histogram_cutoff (const uint64_t* histogram, int histo_max, int min_auto_threshold,
                               uint32_t* cutoff, uint64_t* nb_solids, uint32_t* first_peak);

/* ---- device utilities used by bench.py and the tests ---------------------------------------------------------- */
void* gatb_gpu_malloc (gatb_gpu_ctx*, uint64_t bytes);       /* cudaMalloc on the ctx device                      */
void  gatb_gpu_free   (gatb_gpu_ctx*, void*);
int   gatb_gpu_memcpy_h2d (gatb_gpu_ctx*, void* dst, const void* src, uint64_t bytes);
int   gatb_gpu_memcpy_d2h (gatb_gpu_ctx*, void* dst, const void* src, uint64_t bytes);
int   gatb_gpu_synchronize (gatb_gpu_ctx*);
/* Synthetic reads (DESIGN.md "Synthetic workload"; bit-identical to oracle/kmer_oracle.c orc_synth_reads + orc_pack_2bit):
 * writes reads [first_read, first_read+n_reads) of length L, packed back to back, into d_packed. */
int   gatb_gpu_synth_reads_dev (gatb_gpu_ctx*, uint64_t seed, uint64_t genome_len, uint64_t first_read,
                                uint64_t n_reads, int L, uint8_t* d_packed);
/* Metagenome-like reads (BASELINE.json config 5; bit-identical to oracle/kmer_oracle.c orc_synth_reads_zipf): n_species genomes laid
 * end to end (genome_off[n_species+1], host), the species of a read drawn through the threshold table cdf[n_species] (host,
 * built by the caller, e.g. Zipf with exponent 1.1), then start / strand / 1 % substitutions as above. */
int   gatb_gpu_synth_zipf_dev (gatb_gpu_ctx*, uint64_t seed, uint64_t n_species, const uint64_t* cdf, const uint64_t* genome_off,
                               uint64_t first_read, uint64_t n_reads, int L, uint8_t* d_packed);
/* 2-bit packer for ASCII reads concatenated without separators (host in, host out; n_mask_out may be NULL):
 * the device-side analogue of bank::Sequence -> Data::ConvertASCII (tools/misc/api/Data.hpp:185). */
int   gatb_gpu_pack_ascii (gatb_gpu_ctx*, const char* ascii, uint64_t n, uint8_t* packed_out, uint32_t* n_mask_out,
                           uint64_t* n_invalid);

#ifdef __cplusplus
}
#endif
#endif /* GATB_GPU_H */
