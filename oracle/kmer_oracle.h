/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of GATB-core's k-mer counting path (DSK / SortingCountAlgorithm
 * + Bloom insertion).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this; the product (gatb_core_b200/) never does.
 *
 * PARITY IS PINNED: every function here is checked in tests/ against (1) the golden vectors of the reference's own
 * unit tests (test/unit/src/kmer/TestKmer.cpp, TestDSK.cpp -- transcribed in tests/golden/), and (2) the unmodified
 * reference itself compiled from /root/reference into oracle/_ref/libgatbref.so (ref_harness.cpp), plus committed
 * fixtures generated from it (tests/golden/make_golden.py).
 *
 * Paths below are relative to /root/reference/gatb-core/src/gatb/.
 */
#ifndef KMER_ORACLE_H
#define KMER_ORACLE_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- integer helpers (tools/math/LargeInt1.pri:137-211, LargeInt2.pri:168-251, NativeInt64.hpp:211-221) ---- */
void     orc_revcomp      (uint64_t lo, uint64_t hi, int k, int words, uint64_t* rlo, uint64_t* rhi);
uint64_t orc_hash1        (uint64_t lo, uint64_t hi, int words, uint64_t seed);
uint64_t orc_simplehash16 (uint64_t lo, uint64_t hi, int words, int shift);

/* ---- A1: nucleotide encoding (tools/misc/api/Data.hpp:185-189) ---- */
void orc_encode_ascii (const char* seq, uint64_t n, uint8_t* code, uint8_t* invalid);

/* ---- A3: minimizer LUT (kmer/impl/Model.hpp:1040-1064, is_allowed :1220-1251).  out[4^m] ---- */
void orc_mmer_lut (int m, uint32_t* out);

/* ---- A2+A3: every k-mer of one ASCII sequence: canonical value, minimizer, valid flag, strand (1 = forward) ---- */
int orc_kmers (const char* seq, uint64_t len, int k, int m,
               uint64_t* lo, uint64_t* hi, uint32_t* minim, uint8_t* valid, uint8_t* strand);

/* ---- A4+A5+A6: super-k-mer record streams per partition, byte-identical to what the reference writes into its
 *      SuperKmerBinFiles (block headers removed).  bytes_out[p] is malloc'ed (free with orc_free).
 *      stats_out[0] = nb super-k-mers, [1] = nb k-mers written, [2] = nb valid k-mers, [3] = nb invalid k-mers ---- */
int orc_superkmers (const char* seqs, const uint64_t* offsets, uint64_t nseq, int k, int m, int nb_passes, int pass,
                    const uint16_t* repart, int nb_partitions,
                    uint8_t** bytes_out, uint64_t* sizes_out, uint64_t* stats_out);
void orc_free (void* p);

/* ---- B1: decode one record stream into canonical k-mers (kmer/impl/PartitionsCommand.cpp:420-501).
 *      Returns the number of k-mers; lo/hi may be NULL to count only. ---- */
uint64_t orc_decode_superkmers (const uint8_t* bytes, uint64_t nbytes, int k, uint64_t* lo, uint64_t* hi);

/* ---- full DSK (A1..E): counts every canonical k-mer of the ASCII reads.
 *      Output per (pass,partition) key = pass*nb_partitions+part, ascending k-mer order inside a partition, like
 *      ICountProcessor::process receives them.  Results are owned by the returned handle. ---- */
typedef struct orc_dsk orc_dsk;
orc_dsk* orc_dsk_run (const char* seqs, const uint64_t* offsets, uint64_t nseq, int k, int m, int nb_passes,
                      const uint16_t* repart, int nb_partitions, int abundance_min, int64_t abundance_max,
                      int histo_max, int nthreads);
uint64_t        orc_dsk_part_size   (const orc_dsk*, uint32_t key);      /* distinct k-mers in the partition        */
void            orc_dsk_get_part    (const orc_dsk*, uint32_t key, uint64_t* lo, uint64_t* hi, int32_t* counts);
uint64_t        orc_dsk_solid_size  (const orc_dsk*, uint32_t key);      /* k-mers passing the solidity range       */
void            orc_dsk_get_solid   (const orc_dsk*, uint32_t key, uint64_t* lo, uint64_t* hi, int32_t* counts);
const uint64_t* orc_dsk_histogram   (const orc_dsk*);                    /* u64[histo_max+1], index clamped         */
/* stats: [0] kmers_nb_valid [1] kmers_nb_invalid [2] kmers_nb_distinct [3] kmers_nb_solid [4] nb_superkmers
 *        [5] nb sequences [6] total nucleotides */
const uint64_t* orc_dsk_stats       (const orc_dsk*);
void            orc_dsk_free        (orc_dsk*);

/* ---- C: Histogram::compute_threshold (tools/misc/impl/Histogram.cpp:61-190) ---- */
void orc_histogram_cutoff (const uint64_t* table, int histo_max, int min_auto_threshold,
                           uint32_t* cutoff, uint64_t* nbsolids, uint32_t* first_peak);

/* ---- F: Bloom sizing (kmer/impl/BloomAlgorithm.cpp:155-166, DebloomAlgorithm.cpp:628-650) ---- */
float orc_nbits_per_kmer (int k);
void  orc_bloom_params   (int k, uint64_t nb_solid, uint64_t* bloom_size, int* nb_hash);

/* ---- G: Bloom insertion, kinds "basic" | "cache" | "neighbor" (tools/collections/impl/Bloom.hpp).
 *      bytes_out may be NULL to query nbytes (= 1+tai/8) and bitsize (= getBitSize()). ---- */
int orc_bloom (const char* kind, uint64_t bit_size, int nb_hash, int k, int words,
               const uint64_t* lo, const uint64_t* hi, uint64_t n,
               uint8_t* bytes_out, uint64_t* nbytes_out, uint64_t* bitsize_out);

/* ---- synthetic reads shared with the CUDA generator (definition in DESIGN.md "Synthetic workload") ---- */
uint64_t orc_splitmix64 (uint64_t x);
/* codes A=0 C=1 T=2 G=3, one byte per nucleotide, n_reads*L bytes */
void orc_synth_reads (uint64_t seed, uint64_t genome_len, uint64_t first_read, uint64_t n_reads, int L, uint8_t* codes);
/* packs codes (one per byte) to the 2-bit little-endian stream used by the C-ABI: nt i -> bits [2(i%4), 2(i%4)+2) of byte i/4 */
void orc_zipf_tables (uint64_t seed, uint64_t n_species, double exponent, uint64_t* cdf, uint64_t* genome_off);
void orc_synth_reads_zipf (uint64_t seed, uint64_t n_species, const uint64_t* cdf, const uint64_t* genome_off, uint64_t first_read,
                           uint64_t n_reads, int L, uint8_t* codes);
void orc_pack_2bit (const uint8_t* codes, uint64_t n, uint8_t* packed);
void orc_codes_to_ascii (const uint8_t* codes, uint64_t n, char* ascii);

#ifdef __cplusplus
}
#endif
#endif
