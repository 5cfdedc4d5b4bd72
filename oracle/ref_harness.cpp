/* TEST INFRASTRUCTURE ONLY -- never linked or loaded by the product path.
 *
 * oracle/_ref/libgatbref.so: the UNMODIFIED reference (GATB-core) k-mer counting path, compiled by oracle/Makefile
 * with plain g++ from the sources where they lie under /root/reference (no cmake, no HDF5: see ref_shim/), behind a
 * small extern "C" surface so that tests/ and bench.py's cpu_baseline / --impl reference legs can call it via ctypes.
 *
 * What is the reference's and what is ours: every algorithmic step below is executed by reference classes
 * (Kmer<span>::ModelMinimizer, Sequence2SuperKmer, SuperKmer::save, SuperKmerBinFiles, SortingCountAlgorithm,
 * Repartitor, Bloom*, hash1/simplehash16/revcomp).  This file only instantiates them and copies results out.
 *
 * Reference entry points used:
 *   SortingCountAlgorithm<span>(IProperties*)            gatb/kmer/impl/SortingCountAlgorithm.hpp:85, execute :156
 *   ICountProcessor<span>::process                        gatb/kmer/api/ICountProcessor.hpp:91-183
 *   Sequence2SuperKmer<span>::operator()                  gatb/kmer/impl/Sequence2SuperKmer.hpp:138-159
 *   Kmer<span>::SuperKmer::save                           gatb/kmer/impl/Model.hpp:1386-1471
 *   BloomFactory::createBloom                             gatb/tools/collections/impl/Bloom.hpp:1254-1265
 *   rvalues                                               gatb/kmer/impl/DebloomAlgorithm.pri:27-157
 */
#include <gatb/kmer/impl/Model.hpp>
#include <gatb/kmer/impl/Sequence2SuperKmer.hpp>
#include <gatb/kmer/impl/SortingCountAlgorithm.hpp>
#include <gatb/kmer/impl/CountProcessorAbstract.hpp>
#include <gatb/kmer/impl/PartiInfo.hpp>
#include <gatb/kmer/impl/BankKmers.hpp>
#include <gatb/kmer/impl/DebloomAlgorithm.pri>
#include <gatb/tools/collections/impl/Bloom.hpp>
#include <gatb/tools/misc/impl/Property.hpp>
#include <gatb/tools/misc/impl/Histogram.hpp>
#include <gatb/tools/misc/api/StringsRepository.hpp>
#include <gatb/tools/storage/impl/Storage.hpp>
#include <gatb/bank/impl/Bank.hpp>
#include <gatb/bank/impl/BankStrings.hpp>
#include <gatb/system/impl/System.hpp>

#include <map>
#include <vector>
#include <string>
#include <sstream>
#include <cstring>
#include <cstdio>
#include <pthread.h>

using namespace gatb::core;
using namespace gatb::core::kmer;
using namespace gatb::core::kmer::impl;
using namespace gatb::core::tools::misc;
using namespace gatb::core::tools::misc::impl;
using namespace gatb::core::tools::collections::impl;
using namespace gatb::core::tools::storage::impl;
using namespace gatb::core::bank;
using namespace gatb::core::bank::impl;
using namespace gatb::core::system;
using namespace gatb::core::system::impl;

static std::string g_error;

/* ---- 128-bit split helpers (LargeInt<2> wraps __uint128_t, LargeInt<1> a u64) ---- */
template<typename T> static inline void split(const T& v, uint64_t& lo, uint64_t& hi);
template<> inline void split(const tools::math::LargeInt<1>& v, uint64_t& lo, uint64_t& hi) { lo = v.getVal(); hi = 0; }
template<> inline void split(const tools::math::LargeInt<2>& v, uint64_t& lo, uint64_t& hi)
{ tools::math::LargeInt<2> t = v; lo = t.getVal(); t = t >> 64; hi = t.getVal(); }
template<typename T> static inline T join(uint64_t lo, uint64_t hi);
template<> inline tools::math::LargeInt<1> join(uint64_t lo, uint64_t) { tools::math::LargeInt<1> r; r.setVal(lo); return r; }
template<> inline tools::math::LargeInt<2> join(uint64_t lo, uint64_t hi)
{ tools::math::LargeInt<2> r; r.setVal(hi); r = r << 64; tools::math::LargeInt<2> l; l.setVal(lo); r = r + l; return r; }

/* =====================================================================================================
 *  1. Full DSK run through SortingCountAlgorithm with a capturing count processor
 * ===================================================================================================== */
struct RefRec { uint64_t lo, hi; int32_t count; };

struct RefDsk
{
    int span, nb_partitions, nb_passes, kmer_size, minim_size;
    std::map<uint32_t, std::vector<RefRec> > parts;    /* key = passId*nb_partitions + partId */
    std::vector<uint16_t> repart;
    double seconds, fill_partitions_s, fill_solid_s;
    uint64_t kmers_nb_valid, kmers_nb_invalid, nb_seqs;
    std::vector<uint64_t> histogram;                   /* filled from the reference's own Histogram object */
    std::string info_xml;
};

template<size_t span>
class CaptureProcessor : public CountProcessorAbstract<span>
{
public:
    typedef typename Kmer<span>::Type Type;
    CaptureProcessor (RefDsk* sink, pthread_mutex_t* mtx) : _sink(sink), _mtx(mtx), _key(0), _nbParts(1) {}
    CountProcessorAbstract<span>* clone ()  { CaptureProcessor* c = new CaptureProcessor (_sink, _mtx); c->_nbParts = _nbParts; return c; }
    void begin (const Configuration& config) { _nbParts = config._nb_partitions; }
    void beginPart (size_t passId, size_t partId, size_t cacheSize, const char* name) { _key = passId * _nbParts + partId; _local.clear(); }
    void endPart (size_t passId, size_t partId)
    {
        pthread_mutex_lock (_mtx);
        std::vector<RefRec>& dst = _sink->parts[_key];
        dst.insert (dst.end(), _local.begin(), _local.end());
        pthread_mutex_unlock (_mtx);
        _local.clear();
    }
    bool process (size_t partId, const Type& kmer, const CountVector& count, CountNumber sum)
    {
        if (sum == 0)  { for (size_t i=0; i<count.size(); i++) sum += count[i]; }   /* like CountProcessorChain::process :130 */
        RefRec r; split (kmer, r.lo, r.hi); r.count = sum; _local.push_back (r);
        return true;
    }
private:
    RefDsk* _sink; pthread_mutex_t* _mtx; uint32_t _key; size_t _nbParts; std::vector<RefRec> _local;
};

template<size_t span>
static RefDsk* run_dsk (IProperties* props, int k, int m)
{
    RefDsk* res = new RefDsk();
    res->span = span; res->kmer_size = k; res->minim_size = m;
    pthread_mutex_t mtx = PTHREAD_MUTEX_INITIALIZER;

    SortingCountAlgorithm<span> algo (props);
    CaptureProcessor<span>* proc = new CaptureProcessor<span> (res, &mtx);
    algo.addProcessor (proc);
    algo.execute ();

    const Configuration& cfg = algo.getConfig();
    res->nb_partitions = cfg._nb_partitions;
    res->nb_passes     = cfg._nb_passes;
    Repartitor* rep = algo.getRepartitor();
    uint64_t nbm = (uint64_t)1 << (2*m);
    res->repart.resize (nbm);
    for (uint64_t i=0; i<nbm; i++)  res->repart[i] = (*rep)(i);

    IProperties* info = algo.getInfo();
    res->info_xml = info->getXML();
    res->seconds = 0; res->fill_partitions_s = 0; res->fill_solid_s = 0;
    res->kmers_nb_valid = res->kmers_nb_invalid = res->nb_seqs = 0;
    if (info->get("time"))               res->seconds           = info->getDouble("time");
    if (info->get("fill_partitions"))    res->fill_partitions_s = info->getDouble("fill_partitions");
    if (info->get("fill_solid_kmers"))   res->fill_solid_s      = info->getDouble("fill_solid_kmers");
    if (info->get("kmers_nb_valid"))     res->kmers_nb_valid    = info->getInt("kmers_nb_valid");
    if (info->get("kmers_nb_invalid"))   res->kmers_nb_invalid  = info->getInt("kmers_nb_invalid");
    if (info->get("sequences_number"))   res->nb_seqs           = info->getInt("sequences_number");
    return res;
}

#define API extern "C"

API const char* ref_last_error () { return g_error.c_str(); }

/* Runs the reference DSK.  'input' is a FASTA/FASTQ path (Bank::open).  Returns NULL on exception. */
API RefDsk* ref_dsk_run (const char* input, int k, int m, int abundance_min, int nb_cores, int max_memory_mb,
                     const char* out_prefix, int minimizer_type, int repartition_type)
{
    try
    {
        std::stringstream ss;
        IOptionsParser* parser = SortingCountAlgorithm<>::getOptionsParser (true);
        LOCAL (parser);
        std::vector<std::string> a;
        a.push_back ("ref"); a.push_back (STR_URI_INPUT); a.push_back (input);
        #define PUSHI(opt,val) { std::stringstream s; s << (val); a.push_back (opt); a.push_back (s.str()); }
        PUSHI (STR_KMER_SIZE, k); PUSHI (STR_MINIMIZER_SIZE, m); PUSHI (STR_KMER_ABUNDANCE_MIN, abundance_min);
        PUSHI (STR_MAX_MEMORY, max_memory_mb);
        PUSHI (STR_MINIMIZER_TYPE, minimizer_type); PUSHI (STR_REPARTITION_TYPE, repartition_type);
        a.push_back (STR_STORAGE_TYPE); a.push_back ("file");
        a.push_back (STR_URI_OUTPUT);   a.push_back (out_prefix);
        std::vector<char*> argv; for (size_t i=0; i<a.size(); i++) argv.push_back ((char*)a[i].c_str());
        IProperties* props = parser->parse ((int)argv.size(), argv.data());
        /* -nb-cores / -verbose belong to the Algorithm base option set; set them directly as properties. */
        { std::stringstream s; s << nb_cores; props->add (0, STR_NB_CORES, s.str().c_str()); }
        props->add (0, STR_VERBOSE, "0");
        LOCAL (props);
        if (k < 32)  return run_dsk<32> (props, k, m);
        if (k < 64)  return run_dsk<64> (props, k, m);
        g_error = "ref_dsk_run: k too large for the spans compiled into oracle/_ref (32, 64)";
        return 0;
    }
    catch (Exception& e)        { g_error = e.getMessage(); return 0; }
    catch (std::exception& e)   { g_error = e.what();       return 0; }
    catch (...)                 { g_error = "unknown exception"; return 0; }
}

/* The reference's minimizer -> partition table (Repartitor) for a GIVEN number of partitions, computed by the reference's own
 * RepartitorAlgorithm (kmer/impl/RepartitionAlgorithm.cpp:286-492: samples the bank, balances the minimizer bins) on 'input'.
 * bench.py uses it with the partition count ConfigurationAlgorithm's arithmetic gives for the full-size workload. */
#include <gatb/kmer/impl/RepartitionAlgorithm.hpp>
template<size_t span>
static int run_repartition (const char* input, int k, int m, int nb_partitions, int nb_passes, int nb_cores, const char* tmp_prefix, uint16_t* table)
{
    IBank* bank = Bank::open (input);  LOCAL (bank);
    Configuration config;
    config._kmerSize = k; config._minim_size = m; config._nb_partitions = nb_partitions; config._nb_passes = nb_passes;
    config._repartitionType = 0; config._minimizerType = 0; config._nbCores = nb_cores; config._nb_banks = 1;
    u_int64_t nbSeq = 0, totalSize = 0, maxSize = 0;
    bank->estimate (nbSeq, totalSize, maxSize);
    config._estimateSeqNb = nbSeq; config._estimateSeqTotalSize = totalSize; config._estimateSeqMaxSize = maxSize;
    config._isComputed = true;
    Storage* storage = StorageFactory (STORAGE_FILE).create (tmp_prefix, true, true);  LOCAL (storage);
    RepartitorAlgorithm<span> repart (bank, storage->getGroup ("minimizers"), config, nb_cores);
    repart.execute ();
    Repartitor rep (storage->getGroup ("minimizers"));
    const uint64_t nbm = (uint64_t)1 << (2*m);
    for (uint64_t i=0; i<nbm; i++)  table[i] = rep (i);
    return 0;
}
API int ref_repartition (const char* input, int k, int m, int nb_partitions, int nb_passes, int nb_cores, const char* tmp_prefix, uint16_t* table)
{
    try { return k < 32 ? run_repartition<32> (input, k, m, nb_partitions, nb_passes, nb_cores, tmp_prefix, table)
                        : run_repartition<64> (input, k, m, nb_partitions, nb_passes, nb_cores, tmp_prefix, table); }
    catch (Exception& e)        { g_error = e.getMessage(); return 1; }
    catch (std::exception& e)   { g_error = e.what();       return 1; }
    catch (...)                 { g_error = "unknown exception"; return 1; }
}

API int      ref_dsk_nb_partitions (RefDsk* r) { return r->nb_partitions; }
API int      ref_dsk_nb_passes     (RefDsk* r) { return r->nb_passes; }
API double   ref_dsk_seconds       (RefDsk* r) { return r->seconds; }
API double   ref_dsk_fill_partitions_seconds (RefDsk* r) { return r->fill_partitions_s; }
API double   ref_dsk_fill_solid_seconds      (RefDsk* r) { return r->fill_solid_s; }
API uint64_t ref_dsk_kmers_nb_valid   (RefDsk* r) { return r->kmers_nb_valid; }
API uint64_t ref_dsk_kmers_nb_invalid (RefDsk* r) { return r->kmers_nb_invalid; }
API const char* ref_dsk_info_xml (RefDsk* r) { return r->info_xml.c_str(); }
API uint64_t ref_dsk_nb_distinct (RefDsk* r)
{ uint64_t n=0; for (std::map<uint32_t,std::vector<RefRec> >::iterator it=r->parts.begin(); it!=r->parts.end(); ++it) n += it->second.size(); return n; }
API uint64_t ref_dsk_part_size (RefDsk* r, uint32_t key) { return r->parts.count(key) ? r->parts[key].size() : 0; }
/* Copies partition 'key' (= pass*nb_partitions+part) in the order the reference emitted it. */
API void ref_dsk_get_part (RefDsk* r, uint32_t key, uint64_t* lo, uint64_t* hi, int32_t* counts)
{
    if (!r->parts.count(key)) return;
    std::vector<RefRec>& v = r->parts[key];
    for (size_t i=0; i<v.size(); i++)  { lo[i]=v[i].lo; if (hi) hi[i]=v[i].hi; counts[i]=v[i].count; }
}
API void ref_dsk_get_repart (RefDsk* r, uint16_t* table) { memcpy (table, r->repart.data(), r->repart.size()*sizeof(uint16_t)); }
API void ref_dsk_free (RefDsk* r) { delete r; }

/* =====================================================================================================
 *  2. Histogram semantics (gatb/tools/misc/impl/Histogram.hpp:92, Histogram.cpp:61-190)
 * ===================================================================================================== */
/* Feeds 'n' abundances through the reference Histogram; returns clamped table [0..histo_max] + auto cutoff. */
API void ref_histogram (const int32_t* abundances, uint64_t n, int histo_max, int min_auto_threshold,
                    uint64_t* table_out, uint32_t* cutoff_out, uint64_t* nbsolids_out, uint32_t* first_peak_out)
{
    Histogram h (histo_max);
    for (uint64_t i=0; i<n; i++)  h.inc (abundances[i]);
    h.compute_threshold (min_auto_threshold);
    for (int i=0; i<=histo_max; i++)  table_out[i] = h.get(i);
    *cutoff_out = h.get_solid_cutoff(); *nbsolids_out = h.get_nbsolids_auto(); *first_peak_out = h.get_first_peak();
}

/* =====================================================================================================
 *  3. Per-k-mer canonical value + minimizer, super-k-mer split, super-k-mer serialisation
 * ===================================================================================================== */
template<size_t span>
struct KmerDump
{
    typedef typename Kmer<span>::ModelCanonical MC;
    typedef typename Kmer<span>::template ModelMinimizer<MC> MM;
    uint64_t* lo; uint64_t* hi; uint32_t* minim; uint8_t* valid; uint8_t* strand;
    void operator() (const typename MM::Kmer& kmer, size_t idx)
    {
        split (kmer.value(), lo[idx], hi[idx]);
        minim[idx]  = (uint32_t) kmer.minimizer().value().getVal();
        valid[idx]  = kmer.isValid() ? 1 : 0;
        strand[idx] = kmer.which() ? 1 : 0;
    }
};

template<size_t span>
static int kmers_of (const char* seq, size_t len, int k, int m, uint64_t* lo, uint64_t* hi, uint32_t* minim, uint8_t* valid, uint8_t* strand)
{
    typedef typename Kmer<span>::ModelCanonical MC;
    typedef typename Kmer<span>::template ModelMinimizer<MC> MM;
    MM model (k, m);
    Data data ((char*)seq);
    data.setRef ((char*)seq, len);
    KmerDump<span> f; f.lo=lo; f.hi=hi; f.minim=minim; f.valid=valid; f.strand=strand;
    model.iterate (data, f);
    return (int)len - k + 1;
}

/* canonical value, minimizer value, validity, strand (1 = forward is canonical) of every k-mer of one ASCII sequence */
API int ref_kmers (const char* seq, uint64_t len, int k, int m, uint64_t* lo, uint64_t* hi, uint32_t* minim, uint8_t* valid, uint8_t* strand)
{
    try {
        if (k < 32) return kmers_of<32> (seq, len, k, m, lo, hi, minim, valid, strand);
        if (k < 64) return kmers_of<64> (seq, len, k, m, lo, hi, minim, valid, strand);
        g_error = "k too large"; return -1;
    } catch (Exception& e) { g_error = e.getMessage(); return -1; }
}

/* Sequence2SuperKmer subclass that saves through the reference's own SuperKmer::save into SuperKmerBinFiles */
template<size_t span>
class SuperKmerSaver : public Sequence2SuperKmer<span>
{
public:
    typedef typename Sequence2SuperKmer<span>::Model Model;
    typedef typename Kmer<span>::SuperKmer SuperKmer;
    SuperKmerSaver (Model& model, size_t nbPasses, size_t pass, size_t nbPartitions, BankStats& stats,
                    const uint16_t* repart, SuperKmerBinFiles* files)
        : Sequence2SuperKmer<span> (model, nbPasses, pass, nbPartitions, 0, stats), _repart(repart), _cache (files, 1<<16), nbSuperKmers(0), nbKmers(0) {}
    void processSuperkmer (SuperKmer& superKmer)
    {
        /* same guard as FillPartitions<span,true>::processSuperkmer, gatb/kmer/impl/SortingCountAlgorithm.cpp:1083 */
        if ((superKmer.minimizer % this->_nbPass) == this->_pass && superKmer.isValid())
        {
            size_t p = _repart[superKmer.minimizer];
            superKmer.save (_cache, p);
            nbSuperKmers++; nbKmers += superKmer.size();
        }
    }
    void flush () { _cache.flushAll(); }
    const uint16_t* _repart; CacheSuperKmerBinFiles _cache; uint64_t nbSuperKmers, nbKmers;
};

template<size_t span>
static int superkmers_of (const char* seqs, const uint64_t* offsets, uint64_t nseq, int k, int m, int nb_passes, int pass,
                          const uint16_t* repart, int nb_partitions, const char* tmpdir,
                          uint8_t** bytes_out, uint64_t* sizes_out, uint64_t* stats_out)
{
    typedef typename Kmer<span>::ModelCanonical MC;
    typedef typename Kmer<span>::template ModelMinimizer<MC> MM;
    MM model (k, m);
    BankStats stats;
    SuperKmerBinFiles* files = new SuperKmerBinFiles (tmpdir, "refsk", nb_partitions);
    {
        SuperKmerSaver<span> saver (model, nb_passes, pass, nb_partitions, stats, repart, files);
        for (uint64_t i=0; i<nseq; i++)
        {
            Sequence s ((char*)(seqs + offsets[i]));
            s.getData().setRef ((char*)(seqs + offsets[i]), offsets[i+1]-offsets[i]);
            saver (s);
        }
        saver.flush ();
        stats_out[0] = saver.nbSuperKmers; stats_out[1] = saver.nbKmers;
    }
    files->flushFiles (); files->closeFiles ();
    files->openFiles ("rb");
    for (int p=0; p<nb_partitions; p++)
    {
        std::vector<uint8_t> acc;
        unsigned char* block = 0; unsigned int cap = 0, nb = 0;
        while (files->readBlock (&block, &cap, &nb, p))  acc.insert (acc.end(), block, block+nb);
        if (block) free (block);
        bytes_out[p] = (uint8_t*) malloc (acc.size() ? acc.size() : 1);
        memcpy (bytes_out[p], acc.data(), acc.size());
        sizes_out[p] = acc.size();
    }
    files->closeFiles (); files->eraseFiles ();
    delete files;
    return 0;
}

/* Concatenated ASCII sequences (offsets[nseq+1]); output: per partition, the record stream the reference wrote
 * (block payloads, i.e. the [u32 size] headers of SuperKmerBinFiles::writeBlock removed).  Caller frees with ref_free. */
API int ref_superkmers (const char* seqs, const uint64_t* offsets, uint64_t nseq, int k, int m, int nb_passes, int pass,
                    const uint16_t* repart, int nb_partitions, const char* tmpdir,
                    uint8_t** bytes_out, uint64_t* sizes_out, uint64_t* stats_out)
{
    try {
        if (k < 32) return superkmers_of<32> (seqs, offsets, nseq, k, m, nb_passes, pass, repart, nb_partitions, tmpdir, bytes_out, sizes_out, stats_out);
        if (k < 64) return superkmers_of<64> (seqs, offsets, nseq, k, m, nb_passes, pass, repart, nb_partitions, tmpdir, bytes_out, sizes_out, stats_out);
        g_error = "k too large"; return -1;
    } catch (Exception& e) { g_error = e.getMessage(); return -1; }
}
API void ref_free (void* p) { free (p); }

/* =====================================================================================================
 *  4. Integer helpers and Bloom filters
 * ===================================================================================================== */
API void ref_revcomp (uint64_t lo, uint64_t hi, int k, int words, uint64_t* rlo, uint64_t* rhi)
{
    if (words == 1) { tools::math::LargeInt<1> r = revcomp (join<tools::math::LargeInt<1> >(lo,hi), k); split (r, *rlo, *rhi); }
    else            { tools::math::LargeInt<2> r = revcomp (join<tools::math::LargeInt<2> >(lo,hi), k); split (r, *rlo, *rhi); }
}
API uint64_t ref_hash1 (uint64_t lo, uint64_t hi, int words, uint64_t seed)
{
    return words == 1 ? hash1 (join<tools::math::LargeInt<1> >(lo,hi), seed) : hash1 (join<tools::math::LargeInt<2> >(lo,hi), seed);
}
API uint64_t ref_simplehash16 (uint64_t lo, uint64_t hi, int words, int shift)
{
    return words == 1 ? simplehash16 (join<tools::math::LargeInt<1> >(lo,hi), shift) : simplehash16 (join<tools::math::LargeInt<2> >(lo,hi), shift);
}
API float ref_nbits_per_kmer (int k) { return (float) rvalues[k][1]; }   /* DebloomAlgorithm.cpp:638 (cascading, the default) */

template<typename T>
static int bloom_build (const char* kind, uint64_t bit_size, int nb_hash, int k, const uint64_t* lo, const uint64_t* hi, uint64_t n,
                        uint8_t* bytes_out, uint64_t* nbytes_out, uint64_t* bitsize_out)
{
    BloomKind bk; parse (kind, bk);
    IBloom<T>* bloom = BloomFactory::singleton().createBloom<T> (bk, bit_size, nb_hash, k);
    LOCAL (bloom);
    for (uint64_t i=0; i<n; i++)  bloom->insert (join<T> (lo[i], hi ? hi[i] : 0));
    *nbytes_out  = bloom->getSize();
    *bitsize_out = bloom->getBitSize();
    if (bytes_out)  memcpy (bytes_out, bloom->getArray(), bloom->getSize());
    return 0;
}
/* kind in {"basic","cache","neighbor"}; words = 1 (Kmer<32>) or 2 (Kmer<64>).  bytes_out may be NULL to query sizes. */
API int ref_bloom (const char* kind, uint64_t bit_size, int nb_hash, int k, int words, const uint64_t* lo, const uint64_t* hi, uint64_t n,
               uint8_t* bytes_out, uint64_t* nbytes_out, uint64_t* bitsize_out)
{
    try {
        if (words == 1) return bloom_build<tools::math::LargeInt<1> > (kind, bit_size, nb_hash, k, lo, hi, n, bytes_out, nbytes_out, bitsize_out);
        else            return bloom_build<tools::math::LargeInt<2> > (kind, bit_size, nb_hash, k, lo, hi, n, bytes_out, nbytes_out, bitsize_out);
    } catch (Exception& e) { g_error = e.getMessage(); return -1; }
}

