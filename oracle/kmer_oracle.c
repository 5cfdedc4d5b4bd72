/* TEST INFRASTRUCTURE ONLY -- see kmer_oracle.h.  Plain-C restatement of GATB-core's DSK k-mer counting path.
 * Each function cites the reference lines it follows (paths relative to /root/reference/gatb-core/src/gatb/).
 * K-mers are held in unsigned __int128 (covers Kmer<32> = LargeInt<1> and Kmer<64> = LargeInt<2>, k <= 63).
 */
#include "kmer_oracle.h"
#include "gatb_tables.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;

static const uint64_t random_values[256] = GATB_RANDOM_VALUES_INIT;
static const double   rvalues_col1[129]  = GATB_RVALUES_COL1_INIT;
static const uint8_t  comp_NT[4] = {2, 3, 0, 1};                     /* kmer/impl/ModelData.cpp:41 */

static inline u128 mk128 (uint64_t lo, uint64_t hi) { return ((u128)hi << 64) | lo; }
static inline u128 kmask (int k) { return k >= 64 ? ~(u128)0 : (((u128)1 << (2*k)) - 1); }

/* ------------------------------------------------------------------------------------------------
 * tools/math/LargeInt1.pri:137-154 (revcomp64) -- the 2-bit-group reversal by successive swaps, complement by
 * XOR 0xAA.., then right-aligned to sizeKmer nucleotides.
 * ---------------------------------------------------------------------------------------------- */
static uint64_t revcomp64 (uint64_t x, int sizeKmer)
{
    uint64_t res = x;
    res = ((res >>  2) & 0x3333333333333333ULL) | ((res & 0x3333333333333333ULL) <<  2);
    res = ((res >>  4) & 0x0F0F0F0F0F0F0F0FULL) | ((res & 0x0F0F0F0F0F0F0F0FULL) <<  4);
    res = ((res >>  8) & 0x00FF00FF00FF00FFULL) | ((res & 0x00FF00FF00FF00FFULL) <<  8);
    res = ((res >> 16) & 0x0000FFFF0000FFFFULL) | ((res & 0x0000FFFF0000FFFFULL) << 16);
    res = ((res >> 32) & 0x00000000FFFFFFFFULL) | ((res & 0x00000000FFFFFFFFULL) << 32);
    res ^= 0xAAAAAAAAAAAAAAAAULL;
    if (sizeKmer <= 0) return 0;                 /* (res >> 64) is undefined in C; the reference zeroes that case, LargeInt2.pri:185 */
    return res >> (2*(32 - sizeKmer));
}

/* words==1: LargeInt1.pri:214-219 ; words==2: LargeInt2.pri:168-197 (high/low halves reversed separately, then glued) */
static u128 revcomp_k (u128 x, int k, int words)
{
    if (words == 1)  return revcomp64 ((uint64_t)x, k);
    uint64_t high = (uint64_t)(x >> 64), low = (uint64_t)x;
    int nb_high = k > 32 ? k - 32 : 0;
    int nb_low  = k > 32 ? 32 : k;
    uint64_t rh = (k <= 32) ? 0 : revcomp64 (high, nb_high);
    uint64_t rl = revcomp64 (low, nb_low);
    u128 res = rl;
    res <<= 2*nb_high;
    res += rh;
    return res;
}

/* tools/math/LargeInt1.pri:157-170 (hash64) */
static uint64_t hash64 (uint64_t key, uint64_t seed)
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
/* LargeInt1.pri:251 / LargeInt2.pri:200-206 (xor of the two halves' hashes) */
static uint64_t hash1_k (u128 x, int words, uint64_t seed)
{
    if (words == 1)  return hash64 ((uint64_t)x, seed);
    return hash64 ((uint64_t)(x >> 64), seed) ^ hash64 ((uint64_t)x, seed);
}
/* LargeInt1.pri:190-211 (adds random_values[key&255]) vs LargeInt2.pri:248-251 -> NativeInt64.hpp:211-221 (does not) */
static uint64_t simplehash16_k (u128 x, int words, int shift)
{
    uint64_t key = (uint64_t)x;
    uint64_t input = key >> shift;
    uint64_t res = random_values[input & 255];
    input >>= 8;
    res ^= random_values[input & 255];
    if (words == 1)  res ^= random_values[key & 255];
    return res;
}

void orc_revcomp (uint64_t lo, uint64_t hi, int k, int words, uint64_t* rlo, uint64_t* rhi)
{ u128 r = revcomp_k (mk128 (lo, hi), k, words); *rlo = (uint64_t)r; *rhi = (uint64_t)(r >> 64); }
uint64_t orc_hash1 (uint64_t lo, uint64_t hi, int words, uint64_t seed) { return hash1_k (mk128 (lo, hi), words, seed); }
uint64_t orc_simplehash16 (uint64_t lo, uint64_t hi, int words, int shift) { return simplehash16_k (mk128 (lo, hi), words, shift); }

/* ------------------------------------------------------------------------------------------------
 * A1. tools/misc/api/Data.hpp:185 ConvertASCII: value (c>>1)&3 (A=0 C=1 T=2 G=3), invalid unless c in ACGTacgt
 * (validNucleotide[], tools/misc/api/Data.cpp:3).  'N' therefore encodes as 3 (=G) with the invalid flag set.
 * ---------------------------------------------------------------------------------------------- */
static inline int valid_nt (unsigned char c)
{ return c=='A'||c=='C'||c=='G'||c=='T'||c=='a'||c=='c'||c=='g'||c=='t'; }
void orc_encode_ascii (const char* seq, uint64_t n, uint8_t* code, uint8_t* invalid)
{
    for (uint64_t i=0; i<n; i++)  { code[i] = (seq[i] >> 1) & 3; invalid[i] = !valid_nt ((unsigned char)seq[i]); }
}

/* ------------------------------------------------------------------------------------------------
 * A3. kmer/impl/Model.hpp:1040-1064: lut[x] = min(x, revcomp_m(x)), replaced by 4^m-1 when not allowed;
 * is_allowed :1220-1251 bans an "AA" (two consecutive zero nucleotides) anywhere but at the prefix.
 * ---------------------------------------------------------------------------------------------- */
static int is_allowed (uint32_t mmer, uint32_t len)
{
    uint64_t mmask_m1  = ((uint64_t)1 << ((len-2)*2)) - 1;
    uint64_t mask_0101 = 0x5555555555555555ULL;
    uint64_t mask_ma1  = mask_0101 & mmask_m1;
    uint64_t a1 = mmer;
    a1 = ~(a1 | (a1 >> 2));
    a1 = ((a1 >> 1) & a1) & mask_ma1;
    return a1 == 0;
}
void orc_mmer_lut (int m, uint32_t* out)
{
    uint64_t nb = (uint64_t)1 << (2*m);
    uint32_t mask = (uint32_t)(nb - 1);
    for (uint64_t ii=0; ii<nb; ii++)
    {
        uint32_t mmer = (uint32_t)ii;
        uint32_t rev  = (uint32_t) revcomp64 (ii, m);
        if (rev < mmer)  mmer = rev;
        if (!is_allowed (mmer, m))  mmer = mask;
        out[ii] = mmer;
    }
}

/* ------------------------------------------------------------------------------------------------
 * A2 + A3 state machine: ModelCanonical::first/next (Model.hpp:857-884), ModelAbstract::iterate (:725-765),
 * ModelMinimizer::first/next (:1082-1139), computeNewMinimizerOriginal (:1254-1287); comparator = integer '<'
 * (ComparatorMinimizerFrequencyOrLex without frequencies, :953-971).
 * ---------------------------------------------------------------------------------------------- */
typedef struct
{
    int k, m, words, nbMinimizers;
    u128 kmerMask; u128 revcompTable[4];
    uint32_t mmask; const uint32_t* lut; uint32_t minimizerDefault;
    /* current k-mer */
    u128 fwd, rev; int isValid; int choice;
    uint32_t minimizer; int position; int changed;
} kmodel;

static void kmodel_init (kmodel* M, int k, int m, const uint32_t* lut)
{
    M->k = k; M->m = m; M->words = (k < 32) ? 1 : 2;
    M->kmerMask = kmask (k);
    for (int i=0; i<4; i++)  M->revcompTable[i] = (u128)comp_NT[i] << (2*(k-1));         /* Model.hpp:417 */
    M->nbMinimizers = k - m + 1;
    M->mmask = (uint32_t)(((uint64_t)1 << (2*m)) - 1);
    M->lut = lut; M->minimizerDefault = M->mmask;                                         /* :1032-1034 */
}
static void kmodel_new_minimizer (kmodel* M)                                              /* :1254-1287 */
{
    M->minimizer = M->minimizerDefault; M->position = -1; M->changed = 1;
    u128 val = M->fwd;
    uint32_t best = M->minimizer;
    for (int idx = M->nbMinimizers-1; idx >= 0; idx--)
    {
        uint32_t cand = M->lut[(uint32_t)val & M->mmask];
        if (cand < best)  { M->minimizer = cand; M->position = idx; best = cand; }
        val >>= 2;
    }
}
/* returns index of the last bad character of the first k-mer, or -1 (polynom, Model.hpp:636-657) */
static int kmodel_first (kmodel* M, const char* seq)
{
    int bad = -1; u128 v = 0;
    for (int i=0; i<M->k; i++)
    {
        unsigned char c = (unsigned char)seq[i];
        v = (v << 2) + ((c >> 1) & 3);
        if (!valid_nt (c))  bad = i;
    }
    M->fwd = v; M->isValid = bad < 0;
    M->rev = revcomp_k (v, M->k, M->words);
    M->choice = (M->fwd < M->rev) ? 0 : 1;                                                /* updateChoice :294 */
    kmodel_new_minimizer (M);
    return bad;
}
static void kmodel_next (kmodel* M, int c, int isValid)
{
    M->fwd = ((M->fwd << 2) + (u128)c) & M->kmerMask;                                      /* :878 */
    M->rev = ((M->rev >> 2) + M->revcompTable[c]) & M->kmerMask;                           /* :879 */
    M->isValid = isValid;
    M->choice = (M->fwd < M->rev) ? 0 : 1;
    uint32_t mmer = M->lut[(uint32_t)M->fwd & M->mmask];                                   /* extract :304-315 */
    M->position--; M->changed = 0;
    if (mmer < M->minimizer)  { M->minimizer = mmer; M->position = M->nbMinimizers - 1; M->changed = 1; }
    else if (M->position < 0) { kmodel_new_minimizer (M); }
}
static inline u128 kmodel_value (const kmodel* M) { return M->choice == 0 ? M->fwd : M->rev; }

int orc_kmers (const char* seq, uint64_t len, int k, int m, uint64_t* lo, uint64_t* hi, uint32_t* minim, uint8_t* valid, uint8_t* strand)
{
    int64_t nbKmers = (int64_t)len - k + 1;
    if (nbKmers <= 0) return 0;
    uint32_t* lut = (uint32_t*) malloc (sizeof(uint32_t) << (2*m));
    orc_mmer_lut (m, lut);
    kmodel M; kmodel_init (&M, k, m, lut);
    int indexBadChar = kmodel_first (&M, seq);
    size_t out = 0;
    for (uint64_t idx = k; ; idx++)
    {
        u128 v = kmodel_value (&M);
        lo[out] = (uint64_t)v; if (hi) hi[out] = (uint64_t)(v >> 64);
        minim[out] = M.minimizer; valid[out] = (uint8_t)M.isValid; strand[out] = (M.choice == 0);
        out++;
        if (idx >= len) break;
        unsigned char c = (unsigned char)seq[idx];
        if (!valid_nt (c))  indexBadChar = k - 1;  else indexBadChar--;                   /* :753-754 */
        kmodel_next (&M, (c >> 1) & 3, indexBadChar < 0);
    }
    free (lut);
    return (int)out;
}

/* ------------------------------------------------------------------------------------------------
 * A4..A6: Sequence2SuperKmer (kmer/impl/Sequence2SuperKmer.hpp:81-159), FillPartitions::processSuperkmer
 * (kmer/impl/SortingCountAlgorithm.cpp:1081-1091), SuperKmer::save (Model.hpp:1386-1471),
 * CacheSuperKmerBinFiles::insertSuperkmer (tools/storage/impl/Storage.cpp:567-580).
 * ---------------------------------------------------------------------------------------------- */
typedef struct { uint8_t* p; uint64_t n, cap; } bytebuf;
static void bb_push (bytebuf* b, const uint8_t* src, uint64_t n)
{
    if (b->n + n > b->cap)  { b->cap = (b->cap ? b->cap*2 : 4096); while (b->cap < b->n+n) b->cap *= 2; b->p = (uint8_t*) realloc (b->p, b->cap); }
    memcpy (b->p + b->n, src, n); b->n += n;
}

#define DEFAULT_MINIMIZER 1000000000ULL

typedef struct
{
    uint64_t minimizer;        /* DEFAULT_MINIMIZER = not valid */
    int size;                  /* number of k-mers */
    u128 fwd[256];             /* forward value of each k-mer, like superKmer[i].forward() */
} superkmer;

/* SuperKmer::save, Model.hpp:1386-1471: k/4 full bytes of the first k-mer's forward value (low byte first), then
 * the k%4 remaining nucleotides and the last nucleotide of every following k-mer, packed low-to-high 2 bits. */
static int superkmer_serialize (const superkmer* sk, int k, uint8_t* buf)
{
    int idx = 0;
    u128 base = sk->fwd[0];
    int rem = k;
    while (rem >= 4)  { buf[idx++] = (uint8_t)(base & 255); rem -= 4; base >>= 8; }
    uint8_t newbyte = (uint8_t)(base & 255);
    int uid = rem; int skid = 1;
    for (;;)
    {
        while (uid < 4 && skid < sk->size)  { uint8_t nt = (uint8_t)(sk->fwd[skid] & 3); newbyte |= nt << (uid*2); uid++; skid++; }
        if (uid > 0)  buf[idx++] = newbyte;
        if (skid >= sk->size)  break;
        newbyte = 0; uid = 0;
    }
    return idx;
}

typedef struct { bytebuf* parts; int nb_partitions, nb_passes, pass, k; const uint16_t* repart;
                 uint64_t nbSuperKmers, nbKmersWritten, nbValid, nbInvalid; } sk_sink;

static void process_superkmer (sk_sink* S, const superkmer* sk)                            /* SortingCountAlgorithm.cpp:1081-1091 */
{
    if (sk->minimizer != DEFAULT_MINIMIZER && (sk->minimizer % S->nb_passes) == (uint64_t)S->pass)
    {
        int p = S->repart[sk->minimizer];
        uint8_t rec[1 + 80];
        rec[0] = (uint8_t) sk->size;
        int n = superkmer_serialize (sk, S->k, rec + 1);
        bb_push (&S->parts[p], rec, n + 1);
        S->nbSuperKmers++; S->nbKmersWritten += sk->size;
    }
}

static void sequence_to_superkmers (sk_sink* S, kmodel* M, const char* seq, uint64_t len, int maxs)
{
    int k = M->k;
    int64_t nbKmers = (int64_t)len - k + 1;
    if (nbKmers <= 0) return;                                                              /* Sequence2SuperKmer.hpp:144-145 */
    superkmer sk; sk.minimizer = DEFAULT_MINIMIZER; sk.size = 0;
    int indexBadChar = kmodel_first (M, seq);
    for (uint64_t idx = k; ; idx++)
    {
        /* KmerFunctor::operator(), Sequence2SuperKmer.hpp:90-133 */
        if (!M->isValid)
        {
            process_superkmer (S, &sk);
            sk.size = 0; sk.minimizer = DEFAULT_MINIMIZER;
            S->nbInvalid++;
        }
        else
        {
            S->nbValid++;
            uint64_t h = M->minimizer;
            if (sk.minimizer == DEFAULT_MINIMIZER)  sk.minimizer = h;
            if (h != sk.minimizer || sk.size >= maxs)  { process_superkmer (S, &sk); sk.size = 0; }
            sk.minimizer = h;
            sk.fwd[sk.size++] = M->fwd;
        }
        if (idx >= len) break;
        unsigned char c = (unsigned char)seq[idx];
        if (!valid_nt (c))  indexBadChar = k - 1;  else indexBadChar--;
        kmodel_next (M, (c >> 1) & 3, indexBadChar < 0);
    }
    process_superkmer (S, &sk);                                                            /* :155 */
}

/* maxs = min((Type::getSize()-8)/2, 255), Sequence2SuperKmer.hpp:147: 28 for Kmer<32>, 60 for Kmer<64> */
static int maxs_for (int k) { int bits = (k < 32) ? 64 : 128; int v = (bits - 8)/2; return v < 255 ? v : 255; }

int orc_superkmers (const char* seqs, const uint64_t* offsets, uint64_t nseq, int k, int m, int nb_passes, int pass,
                    const uint16_t* repart, int nb_partitions, uint8_t** bytes_out, uint64_t* sizes_out, uint64_t* stats_out)
{
    uint32_t* lut = (uint32_t*) malloc (sizeof(uint32_t) << (2*m));
    orc_mmer_lut (m, lut);
    kmodel M; kmodel_init (&M, k, m, lut);
    sk_sink S; memset (&S, 0, sizeof(S));
    S.parts = (bytebuf*) calloc (nb_partitions, sizeof(bytebuf));
    S.nb_partitions = nb_partitions; S.nb_passes = nb_passes; S.pass = pass; S.k = k; S.repart = repart;
    int maxs = maxs_for (k);
    for (uint64_t i=0; i<nseq; i++)  sequence_to_superkmers (&S, &M, seqs + offsets[i], offsets[i+1]-offsets[i], maxs);
    for (int p=0; p<nb_partitions; p++)  { bytes_out[p] = S.parts[p].p ? S.parts[p].p : (uint8_t*) malloc (1); sizes_out[p] = S.parts[p].n; }
    stats_out[0] = S.nbSuperKmers; stats_out[1] = S.nbKmersWritten; stats_out[2] = S.nbValid; stats_out[3] = S.nbInvalid;
    free (S.parts); free (lut);
    return 0;
}
void orc_free (void* p) { free (p); }

/* ------------------------------------------------------------------------------------------------
 * B1: kmer/impl/PartitionsCommand.cpp:420-501 -- rebuild the seed k-mer from k/4 (+1) bytes, then roll one
 * nucleotide per following k-mer; canonical = min(forward, revcomp).
 * ---------------------------------------------------------------------------------------------- */
uint64_t orc_decode_superkmers (const uint8_t* bytes, uint64_t nbytes, int k, uint64_t* lo, uint64_t* hi)
{
    const uint8_t* ptr = bytes; const uint8_t* end = bytes + nbytes;
    int words = (k < 32) ? 1 : 2;
    u128 kmerMask = kmask (k); int shift = 2*(k-1);
    uint64_t out = 0;
    while (ptr < end)
    {
        uint8_t nbK = *ptr++;
        int rem_size = k; int nbr = 0; u128 seedk = 0; uint8_t newbyte = 0;
        while (rem_size >= 4)  { newbyte = *ptr++; seedk |= (u128)newbyte << (8*nbr); rem_size -= 4; nbr++; }
        int uid = 4;
        if (rem_size > 0)  { newbyte = *ptr++; seedk |= (u128)newbyte << (8*nbr); uid = rem_size; }
        seedk &= kmerMask;
        uint8_t rem = nbK;
        u128 temp = seedk, rev_temp = revcomp_k (temp, k, words);
        for (int ii=0; ii<nbK; ii++, rem--)
        {
            u128 mink = rev_temp < temp ? rev_temp : temp;
            if (lo)  { lo[out] = (uint64_t)mink; if (hi) hi[out] = (uint64_t)(mink >> 64); }
            out++;
            if (rem < 2) break;
            if (uid >= 4)  { newbyte = *ptr++; uid = 0; }
            u128 newnt = (newbyte >> (2*uid)) & 3; uid++;
            temp = ((temp << 2) | newnt) & kmerMask;
            rev_temp = ((rev_temp >> 2) | ((u128)comp_NT[(int)newnt] << shift)) & kmerMask;
        }
    }
    return out;
}

/* ------------------------------------------------------------------------------------------------
 * Full DSK: SortingCountAlgorithm::execute (kmer/impl/SortingCountAlgorithm.cpp:636-781): per pass, stage 1
 * fillPartitions (super-k-mers into partition p = repart[minimizer]), stage 2 fillSolidKmers (decode, sort,
 * count, emit ascending: PartitionsCommand.cpp:1205-1239, executeDump :1599-1805), then the default processor
 * chain: histogram of every distinct k-mer (CountProcessorHistogram.hpp:173, Histogram.hpp:92), solidity
 * abundance_min <= sum <= abundance_max (CountProcessorSolidity.hpp:186), dump (CountProcessorDump.hpp:148).
 * ---------------------------------------------------------------------------------------------- */
typedef struct { u128* kmers; int32_t* counts; uint64_t n; uint64_t nsolid; } dsk_part;
struct orc_dsk
{
    int nb_partitions, nb_passes, histo_max, abundance_min; int64_t abundance_max;
    dsk_part* parts;           /* nb_passes * nb_partitions */
    uint64_t* histogram;       /* histo_max + 1 */
    uint64_t stats[8];
};

static int cmp_u128 (const void* a, const void* b)
{ u128 x = *(const u128*)a, y = *(const u128*)b; return x < y ? -1 : (x > y ? 1 : 0); }

orc_dsk* orc_dsk_run (const char* seqs, const uint64_t* offsets, uint64_t nseq, int k, int m, int nb_passes,
                      const uint16_t* repart, int nb_partitions, int abundance_min, int64_t abundance_max,
                      int histo_max, int nthreads)
{
    orc_dsk* D = (orc_dsk*) calloc (1, sizeof(orc_dsk));
    D->nb_partitions = nb_partitions; D->nb_passes = nb_passes; D->histo_max = histo_max;
    D->abundance_min = abundance_min; D->abundance_max = abundance_max;
    D->parts = (dsk_part*) calloc ((size_t)nb_passes * nb_partitions, sizeof(dsk_part));
    D->histogram = (uint64_t*) calloc (histo_max + 1, sizeof(uint64_t));
    uint32_t* lut = (uint32_t*) malloc (sizeof(uint32_t) << (2*m));
    orc_mmer_lut (m, lut);
    int maxs = maxs_for (k);
#ifdef _OPENMP
    if (nthreads <= 0)  nthreads = omp_get_max_threads ();
#else
    nthreads = 1;
#endif
    for (uint64_t i=0; i<nseq; i++)  D->stats[6] += offsets[i+1] - offsets[i];
    D->stats[5] = nseq;

    for (int pass=0; pass<nb_passes; pass++)
    {
        /* ---- stage 1: each thread fills private per-partition record streams (like the reference's per-thread caches) */
        sk_sink* sinks = (sk_sink*) calloc (nthreads, sizeof(sk_sink));
        #pragma omp parallel num_threads(nthreads)
        {
#ifdef _OPENMP
            int t = omp_get_thread_num ();
#else
            int t = 0;
#endif
            sk_sink* S = &sinks[t];
            S->parts = (bytebuf*) calloc (nb_partitions, sizeof(bytebuf));
            S->nb_partitions = nb_partitions; S->nb_passes = nb_passes; S->pass = pass; S->k = k; S->repart = repart;
            kmodel M; kmodel_init (&M, k, m, lut);
            #pragma omp for schedule(dynamic, 1000)
            for (uint64_t i=0; i<nseq; i++)  sequence_to_superkmers (S, &M, seqs + offsets[i], offsets[i+1]-offsets[i], maxs);
        }
        for (int t=0; t<nthreads; t++)
        {
            if (pass == 0)  { D->stats[0] += sinks[t].nbValid; D->stats[1] += sinks[t].nbInvalid; }
            D->stats[4] += sinks[t].nbSuperKmers;
        }
        /* ---- stage 2: per partition decode + sort + run-length count, ascending emission */
        #pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
        for (int p=0; p<nb_partitions; p++)
        {
            uint64_t nk = 0;
            for (int t=0; t<nthreads; t++)  nk += orc_decode_superkmers (sinks[t].parts[p].p, sinks[t].parts[p].n, k, 0, 0);
            u128* all = (u128*) malloc ((nk ? nk : 1) * sizeof(u128));
            uint64_t* tlo = (uint64_t*) malloc ((nk ? nk : 1) * sizeof(uint64_t));
            uint64_t* thi = (uint64_t*) malloc ((nk ? nk : 1) * sizeof(uint64_t));
            uint64_t pos = 0;
            for (int t=0; t<nthreads; t++)  pos += orc_decode_superkmers (sinks[t].parts[p].p, sinks[t].parts[p].n, k, tlo + pos, thi + pos);
            for (uint64_t i=0; i<nk; i++)  all[i] = mk128 (tlo[i], thi[i]);
            free (tlo); free (thi);
            qsort (all, nk, sizeof(u128), cmp_u128);
            uint64_t nd = 0;
            for (uint64_t i=0; i<nk; i++)  if (i == 0 || all[i] != all[i-1])  nd++;
            dsk_part* P = &D->parts[(size_t)pass * nb_partitions + p];
            P->kmers = (u128*) malloc ((nd ? nd : 1) * sizeof(u128));
            P->counts = (int32_t*) malloc ((nd ? nd : 1) * sizeof(int32_t));
            P->n = nd; P->nsolid = 0;
            uint64_t j = 0;
            for (uint64_t i=0; i<nk; )
            {
                uint64_t e = i + 1; while (e < nk && all[e] == all[i]) e++;
                P->kmers[j] = all[i]; P->counts[j] = (int32_t)(e - i); j++;
                i = e;
            }
            free (all);
        }
        for (int t=0; t<nthreads; t++)
        {
            for (int p=0; p<nb_partitions; p++)  free (sinks[t].parts[p].p);
            free (sinks[t].parts);
        }
        free (sinks);
    }
    /* processor chain */
    for (size_t q=0; q<(size_t)nb_passes*nb_partitions; q++)
    {
        dsk_part* P = &D->parts[q];
        for (uint64_t i=0; i<P->n; i++)
        {
            int64_t c = P->counts[i];
            D->histogram[c >= histo_max ? histo_max : c]++;                                /* Histogram.hpp:92 */
            if (c >= abundance_min && c <= abundance_max)  P->nsolid++;                    /* Range.hpp:75 */
        }
        D->stats[2] += P->n; D->stats[3] += P->nsolid;
    }
    free (lut);
    return D;
}
uint64_t orc_dsk_part_size (const orc_dsk* D, uint32_t key) { return D->parts[key].n; }
void orc_dsk_get_part (const orc_dsk* D, uint32_t key, uint64_t* lo, uint64_t* hi, int32_t* counts)
{
    const dsk_part* P = &D->parts[key];
    for (uint64_t i=0; i<P->n; i++)  { lo[i] = (uint64_t)P->kmers[i]; if (hi) hi[i] = (uint64_t)(P->kmers[i] >> 64); counts[i] = P->counts[i]; }
}
uint64_t orc_dsk_solid_size (const orc_dsk* D, uint32_t key) { return D->parts[key].nsolid; }
void orc_dsk_get_solid (const orc_dsk* D, uint32_t key, uint64_t* lo, uint64_t* hi, int32_t* counts)
{
    const dsk_part* P = &D->parts[key]; uint64_t j = 0;
    for (uint64_t i=0; i<P->n; i++)
    {
        int64_t c = P->counts[i];
        if (c >= D->abundance_min && c <= D->abundance_max)
        { lo[j] = (uint64_t)P->kmers[i]; if (hi) hi[j] = (uint64_t)(P->kmers[i] >> 64); counts[j] = P->counts[i]; j++; }
    }
}
const uint64_t* orc_dsk_histogram (const orc_dsk* D) { return D->histogram; }
const uint64_t* orc_dsk_stats (const orc_dsk* D) { return D->stats; }
void orc_dsk_free (orc_dsk* D)
{
    if (!D) return;
    for (size_t q=0; q<(size_t)D->nb_passes*D->nb_partitions; q++)  { free (D->parts[q].kmers); free (D->parts[q].counts); }
    free (D->parts); free (D->histogram); free (D);
}

/* ------------------------------------------------------------------------------------------------
 * C: tools/misc/impl/Histogram.cpp:61-190 (compute_threshold).  table[0..histo_max]; _length = histo_max.
 * ---------------------------------------------------------------------------------------------- */
void orc_histogram_cutoff (const uint64_t* h, int histo_max, int min_auto_threshold, uint32_t* cutoff_out, uint64_t* nbsolids_out, uint32_t* first_peak_out)
{
    size_t length = histo_max;
    uint64_t* sm = (uint64_t*) calloc (length + 2, sizeof(uint64_t));
    uint64_t sum_allk = 0;
    uint32_t cutoff = 0; uint32_t firstPeak = 0; uint64_t nbsolids = 0;
    if (length >= 2)  { sm[1] = (uint64_t)(0.6 * (double)h[1] + 0.4 * (double)h[2]); sum_allk += h[1]; }
    int index_first_increase = -1, index_maxval_after_first_increase = -1;
    uint64_t max_val = 0;
    for (size_t i=2; i<length; i++)
    {
        sum_allk += h[i] * i;
        sm[i] = (uint64_t)(0.2 * (double)h[i-1] + 0.6 * (double)h[i] + 0.2 * (double)h[i+1]);
        if (index_first_increase == -1 && sm[i-1] < sm[i])  index_first_increase = (int)i - 1;
        if (index_first_increase > 0 && sm[i] > max_val)  { max_val = sm[i]; index_maxval_after_first_increase = (int)i; }
    }
    sum_allk += h[length] * length;
    if (index_first_increase == -1)
    {
        *cutoff_out = (uint32_t)min_auto_threshold; *nbsolids_out = 0; *first_peak_out = 0;
        free (sm); return;
    }
    firstPeak = index_maxval_after_first_increase;
    uint64_t min_val = 10000000000ULL; int index_minval = -1;
    for (int i=index_first_increase; i<=index_maxval_after_first_increase; i++)
        if (sm[i] < min_val)  { min_val = sm[i]; index_minval = i; }
    if (index_minval != -1)  cutoff = index_minval;
    uint64_t sum_elim = 0; size_t max_cutoff = 0;
    for (size_t i=0; i<length+1; i++)
    {
        sum_elim += h[i] * i;
        double ratio = (double)sum_elim / sum_allk;
        if (ratio >= 0.25)  { max_cutoff = i + 1; break; }
    }
    if (cutoff > max_cutoff)  cutoff = (uint32_t)max_cutoff;
    if (cutoff < (size_t)min_auto_threshold)  cutoff = (uint32_t)min_auto_threshold;
    for (size_t i=cutoff; i<length+1; i++)  nbsolids += h[i];
    *cutoff_out = cutoff; *nbsolids_out = nbsolids; *first_peak_out = firstPeak;
    free (sm);
}

/* ------------------------------------------------------------------------------------------------
 * F: kmer/impl/BloomAlgorithm.cpp:158-166: float32 product, nbHash = floorf(0.7*bits); bits = rvalues[k][1]
 * (kmer/impl/DebloomAlgorithm.cpp:638, cascading debloom = default), 1 if 0 (:648).
 * ---------------------------------------------------------------------------------------------- */
float orc_nbits_per_kmer (int k)
{
    float v = (float) rvalues_col1[k];
    if (v == 0) v = 1;
    return v;
}
void orc_bloom_params (int k, uint64_t nb_solid, uint64_t* bloom_size, int* nb_hash)
{
    float NBITS = orc_nbits_per_kmer (k);
    uint64_t est = (uint64_t)(nb_solid * NBITS);             /* u64 * float -> float32 product, as in the reference */
    *nb_hash = (int) floorf (0.7 * NBITS);
    if (est == 0)  est = 1000;
    *bloom_size = est;
}

/* ------------------------------------------------------------------------------------------------
 * G: tools/collections/impl/Bloom.hpp: HashFunctors :59-98, BloomContainer ctor :184-199, BloomSynchronized::insert
 * :394-412, BloomCacheCoherent ctor/insert :437-459, BloomNeighborCoherent ctor/insert :523-588.
 * ---------------------------------------------------------------------------------------------- */
static void bloom_seeds (uint64_t seed_tab[10])
{
    static const uint64_t rbase[10] = {
        0xAAAAAAAA55555555ULL, 0x33333333CCCCCCCCULL, 0x6666666699999999ULL, 0xB5B5B5B54B4B4B4BULL,
        0xAA55AA5555335533ULL, 0x33CC33CCCC66CC66ULL, 0x6699669999B599B5ULL, 0xB54BB54B4BAA4BAAULL,
        0xAA33AA3355CC55CCULL, 0x33663366CC99CC99ULL };
    for (int i=0; i<10; i++)  seed_tab[i] = rbase[i];
    for (int i=0; i<10; i++)  seed_tab[i] = seed_tab[i] * seed_tab[(i+3) % 10] + 0;      /* in-place, sequential: :92-93 */
}
static inline void setbit (uint8_t* a, uint64_t h) { a[h >> 3] |= (uint8_t)(1u << (h & 7)); }

int orc_bloom (const char* kind, uint64_t bit_size, int nb_hash, int k, int words,
               const uint64_t* lo, const uint64_t* hi, uint64_t n, uint8_t* bytes_out, uint64_t* nbytes_out, uint64_t* bitsize_out)
{
    int is_basic = !strcmp (kind, "basic"), is_cache = !strcmp (kind, "cache"), is_neigh = !strcmp (kind, "neighbor");
    if (!is_basic && !is_cache && !is_neigh) return -1;
    uint64_t tai = is_basic ? bit_size : bit_size + 2*4096;                                /* :438 block_nbits = 12 */
    uint64_t nchar = 1 + tai/8;                                                            /* :187 */
    int pow2 = (tai && !(tai & (tai - 1)));
    if (pow2) tai--;                                                                       /* :193-198 */
    uint64_t reduced = tai - 2*4096;                                                       /* :441 (cache, neighbor) */
    *nbytes_out = nchar; *bitsize_out = is_basic ? tai : reduced;
    if (!bytes_out) return 0;
    memset (bytes_out, 0, nchar);
    uint64_t seeds[10]; bloom_seeds (seeds);
    uint64_t mask_block = 4095;
    static const uint8_t cano2[16] = {0,1,2,3,4,5,3,7,8,9,0,4,9,13,1,5};                   /* :526-541 */
    for (uint64_t i=0; i<n; i++)
    {
        u128 item = mk128 (lo[i], hi ? hi[i] : 0);
        if (is_basic)
        {
            for (int f=0; f<nb_hash; f++)
            {
                uint64_t h = hash1_k (item, words, seeds[f]);
                h = pow2 ? (h & tai) : (h % tai);
                setbit (bytes_out, h);
            }
        }
        else if (is_cache)
        {
            uint64_t h0 = hash1_k (item, words, seeds[0]) % reduced;
            setbit (bytes_out, h0);
            for (int f=1; f<nb_hash; f++)  setbit (bytes_out, h0 + (simplehash16_k (item, words, f) & mask_block));
        }
        else
        {
            u128 maskkm2  = kmask (k-2);
            u128 prefmask = (u128)3 << ((k-1)*2);
            u128 suffix = item & 3;
            u128 prefix = (item & prefmask) >> ((k-2)*2);
            prefix += suffix; prefix &= 15;
            uint64_t pref_val = cano2[(int)prefix];
            u128 hashpart = (item >> 2) & maskkm2;
            u128 rev = revcomp_k (hashpart, k-2, words);
            if (rev < hashpart) hashpart = rev;
            uint64_t racine = hash1_k (hashpart, words, seeds[0]) % reduced;
            uint64_t h0 = racine + pref_val;
            setbit (bytes_out, h0);
            for (int f=1; f<nb_hash; f++)  setbit (bytes_out, h0 + (simplehash16_k (hashpart, words, f) & mask_block));
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Synthetic workload (ours, not the reference's; mirrored bit-for-bit by the CUDA generator in
 * gatb_core_b200/csrc/synth.cu).  Counter-based so any slice can be produced independently.
 * ---------------------------------------------------------------------------------------------- */
uint64_t orc_splitmix64 (uint64_t x)
{
    uint64_t z = x + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline uint8_t genome_base (uint64_t seed, uint64_t p) { return (uint8_t)(orc_splitmix64 (seed * 0x100000001B3ULL + p) >> 61) & 3; }
void orc_synth_reads (uint64_t seed, uint64_t genome_len, uint64_t first_read, uint64_t n_reads, int L, uint8_t* codes)
{
    uint64_t sR = orc_splitmix64 (seed ^ 0x5EEDC0DE00000001ULL), sE = orc_splitmix64 (seed ^ 0x5EEDC0DE00000002ULL);
    #pragma omp parallel for schedule(static)
    for (uint64_t i=0; i<n_reads; i++)
    {
        uint64_t r = first_read + i;
        uint64_t h0 = orc_splitmix64 (sR + 2*r), h1 = orc_splitmix64 (sR + 2*r + 1);
        uint64_t start = h0 % (genome_len - L + 1);
        int flip = (int)(h1 >> 63);
        uint8_t* out = codes + i * (uint64_t)L;
        for (int j=0; j<L; j++)
        {
            uint8_t b = genome_base (seed, start + j);
            uint64_t e = orc_splitmix64 (sE + r * (uint64_t)L + j);
            if ((e % 100) == 0)  b = (uint8_t)((b + 1 + ((e >> 32) % 3)) & 3);            /* 1 % substitutions */
            if (flip)  out[L-1-j] = b ^ 2;  else out[j] = b;                                /* complement: A<->T, C<->G is xor 2 */
        }
    }
}
/* ---- metagenome-like workload (BASELINE.json config 5): n_species genomes of 10^5..10^6 nt laid end to end in one
 * coordinate space (genome_off[s] .. genome_off[s+1]), the species of a read drawn from Zipf(exponent) through a table of
 * 64-bit thresholds (species = first s with h < cdf[s]); otherwise like orc_synth_reads (start, strand flip, 1 % substitutions).
 * The tables are computed ONCE on the host (doubles, pow) and handed to the device generator as they are, so both sides use the
 * same bits. ---- */
void orc_zipf_tables (uint64_t seed, uint64_t n_species, double exponent, uint64_t* cdf, uint64_t* genome_off)
{
    double total = 0, run = 0;
    for (uint64_t s=0; s<n_species; s++)  total += 1.0 / pow ((double)(s+1), exponent);
    genome_off[0] = 0;
    for (uint64_t s=0; s<n_species; s++)
    {
        run += 1.0 / pow ((double)(s+1), exponent);
        double f = run / total;
        cdf[s] = (s+1 == n_species || f >= 1.0) ? ~0ULL : (uint64_t)(f * 18446744073709551615.0);
        genome_off[s+1] = genome_off[s] + 100000 + orc_splitmix64 ((seed ^ 0x5EEDC0DE00000003ULL) + s) % 900001;
    }
}
void orc_synth_reads_zipf (uint64_t seed, uint64_t n_species, const uint64_t* cdf, const uint64_t* genome_off, uint64_t first_read,
                           uint64_t n_reads, int L, uint8_t* codes)
{
    uint64_t sR = orc_splitmix64 (seed ^ 0x5EEDC0DE00000001ULL), sE = orc_splitmix64 (seed ^ 0x5EEDC0DE00000002ULL);
    #pragma omp parallel for schedule(static)
    for (uint64_t i=0; i<n_reads; i++)
    {
        uint64_t r = first_read + i;
        uint64_t h0 = orc_splitmix64 (sR + 3*r), h1 = orc_splitmix64 (sR + 3*r + 1), h2 = orc_splitmix64 (sR + 3*r + 2);
        uint64_t lo = 0, hi = n_species - 1;                       /* first s with h2 <= cdf[s] */
        while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (h2 <= cdf[mid]) hi = mid; else lo = mid + 1; }
        uint64_t glen = genome_off[lo+1] - genome_off[lo];
        uint64_t start = genome_off[lo] + h0 % (glen - L + 1);
        int flip = (int)(h1 >> 63);
        uint8_t* out = codes + i * (uint64_t)L;
        for (int j=0; j<L; j++)
        {
            uint8_t b = genome_base (seed, start + j);
            uint64_t e = orc_splitmix64 (sE + r * (uint64_t)L + j);
            if ((e % 100) == 0)  b = (uint8_t)((b + 1 + ((e >> 32) % 3)) & 3);
            if (flip)  out[L-1-j] = b ^ 2;  else out[j] = b;
        }
    }
}
void orc_pack_2bit (const uint8_t* codes, uint64_t n, uint8_t* packed)
{
    memset (packed, 0, (n + 3)/4);
    for (uint64_t i=0; i<n; i++)  packed[i >> 2] |= (uint8_t)((codes[i] & 3) << (2*(i & 3)));
}
void orc_codes_to_ascii (const uint8_t* codes, uint64_t n, char* ascii)
{
    static const char bin2NT[4] = {'A','C','T','G'};
    for (uint64_t i=0; i<n; i++)  ascii[i] = bin2NT[codes[i] & 3];
}
