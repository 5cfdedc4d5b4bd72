/* integration/SortingCountAlgorithmGPU.cpp -- GATB-core's own SortingCountAlgorithm<span>, with its two hot stages served by
 * libgatb_b200.so through the C ABI of include/gatb_gpu.h.
 *
 * This translation unit REPLACES the reference's explicit instantiation of SortingCountAlgorithm<span> (the cmake-generated
 * gatb/template/TemplateSpecialization*.cpp) when a GATB tool is linked against the GPU path.  Nothing of the reference is
 * copied or patched: its template definitions are included from where they lie (kmer/impl/SortingCountAlgorithm.cpp) and
 *   SortingCountAlgorithm<span>::fillPartitions        kmer/impl/SortingCountAlgorithm.cpp:1211-1344   (stage 1: bank -> super-k-mer partitions)
 *   SortingCountAlgorithm<span>::fillSolidKmers_aux    kmer/impl/SortingCountAlgorithm.cpp:1409-1602   (stage 2: partitions -> ICountProcessor)
 * are given explicit specialisations for span 32 and 64 BEFORE the class is explicitly instantiated, so every other member
 * (constructors from IProperties*, configure(), execute(), getSolidCounts(), getStorage(), the option parser, the default
 * processor chain, the info/statistics tree) is the reference's code, unmodified, calling into these two.
 * Downstream code written against GATB -- Graph::create (debruijn/impl/Graph.cpp:399-407), examples/kmer/kmer12.cpp style
 * custom ICountProcessor<span> chains, dbgh5 -- compiles and links unchanged.
 *
 * Contract kept (kmer/api/ICountProcessor.hpp:91-183): per pass, for every partition a clone of the processor receives
 * beginPart -> process(partId, kmer, CountVector{count}, sum) for every distinct k-mer in ascending order -> endPart, from
 * dispatcher threads (one clone per partition), then finishClones on the prototype.  There is no CPU fallback: multi-bank
 * solidity kinds and frequency-ordered minimizers, which the device path does not serve, raise system::Exception.
 */
#include <gatb/kmer/impl/SortingCountAlgorithm.cpp>
#include <gatb/kmer/impl/PartitionsCommand.cpp>

#include "../include/gatb_gpu.h"

#include <map>
#include <string>
#include <vector>
#include <cstdlib>
#include <cstring>
#include <pthread.h>

namespace gatb { namespace core { namespace kmer { namespace impl {

namespace gpu_dsk {

/* One run per algorithm object: the context and the result of the count (all passes at once), alive from pass 0's
 * fillPartitions to the last processor of the last pass. */
struct Run
{
    gatb_gpu_ctx*   ctx;
    gatb_gpu_result res;
    bool            counted;
    size_t          replays_left;
    Run () : ctx(0), counted(false), replays_left(0) { memset (&res, 0, sizeof(res)); }
    void close ()
    {
        if (ctx) { if (counted) gatb_gpu_result_free (ctx, &res); gatb_gpu_destroy (ctx); ctx = 0; }
        counted = false;
    }
};
static std::map<const void*, Run> g_runs;
static pthread_mutex_t g_mtx = PTHREAD_MUTEX_INITIALIZER;

static Run& run_of (const void* algo)
{
    pthread_mutex_lock (&g_mtx);
    Run& r = g_runs[algo];
    pthread_mutex_unlock (&g_mtx);
    return r;
}
static int device_id () { const char* e = getenv ("GATB_GPU_DEVICE"); return e ? atoi (e) : 0; }

static void check (gatb_gpu_ctx* ctx, int rc, const char* what)
{ if (rc != 0) throw system::Exception ("GPU k-mer counting failed in %s: %s", what, gatb_gpu_last_error (ctx)); }

template<typename Type> static inline Type make_kmer (uint64_t lo, uint64_t hi);
template<> inline tools::math::LargeInt<1> make_kmer (uint64_t lo, uint64_t)    { tools::math::LargeInt<1> r; r.setVal (lo); return r; }
template<> inline tools::math::LargeInt<2> make_kmer (uint64_t lo, uint64_t hi)
{ tools::math::LargeInt<2> r; r.setVal (hi); r = r << 64; tools::math::LargeInt<2> l; l.setVal (lo); return r + l; }

/* ---- stage 1: bank -> device.  Sequences are pushed as ASCII in batches and packed on the device. ---- */
static void push_bank (gatb_gpu_ctx* ctx, tools::dp::Iterator<bank::Sequence>* itSeq, BankStats& stats, tools::dp::IteratorListener* progress)
{
    const size_t BATCH = 64u << 20;
    std::string blob; blob.reserve (BATCH + (1u << 20));
    std::vector<uint64_t> offs; offs.push_back (0);
    check (ctx, gatb_gpu_reads_begin (ctx, 0), "gatb_gpu_reads_begin");
    for (itSeq->first (); !itSeq->isDone (); itSeq->next ())
    {
        bank::Sequence& seq = itSeq->item ();
        stats.update (seq);
        const size_t n = seq.getDataSize ();
        const char* d = seq.getDataBuffer ();
        if (seq.getDataEncoding () == tools::misc::Data::ASCII) blob.append (d, n);
        else if (seq.getDataEncoding () == tools::misc::Data::INTEGER) { for (size_t i = 0; i < n; i++) blob.push_back ("ACTG"[d[i] & 3]); }
        else { for (size_t i = 0; i < n; i++) blob.push_back ("ACTG"[(d[i / 4] >> (2 * (3 - (i & 3)))) & 3]); }       /* BINARY: first nt in the top bits */
        offs.push_back (blob.size ());
        if (blob.size () >= BATCH)
        {
            check (ctx, gatb_gpu_reads_push_ascii (ctx, blob.data (), offs.data (), offs.size () - 1), "gatb_gpu_reads_push_ascii");
            if (progress) progress->inc (blob.size ());
            blob.clear (); offs.clear (); offs.push_back (0);
        }
    }
    if (offs.size () > 1) check (ctx, gatb_gpu_reads_push_ascii (ctx, blob.data (), offs.data (), offs.size () - 1), "gatb_gpu_reads_push_ascii");
    itSeq->finalize ();
}

/* ---- stage 1, fast path: the bank is ONE plain FASTA / FASTQ file.  The file is read in batches cut at record boundaries and
 * parsed on the device (gatb_gpu_reads_push_text): the serial line parser of the reference (BankFasta.cpp:391-620, under the
 * iterator lock of ICommand.hpp:313-319) is out of the way.  Returns false when the bank is something else (gzip, album,
 * several files, in-memory bank): the caller then iterates it through the reference's own iterator. ---- */
static bool push_plain_file (gatb_gpu_ctx* ctx, bank::IBank* bank, BankStats& stats, tools::dp::IteratorListener* progress)
{
    if (getenv ("GATB_GPU_NO_TEXT_PARSER")) return false;
    const std::string path = bank->getId ();
    if (path.empty () || path.find (',') != std::string::npos) return false;
    if (bank->getCompositionNb () > 1) return false;
    FILE* f = fopen (path.c_str (), "rb");
    if (!f) return false;
    unsigned char magic[2] = { 0, 0 };
    const size_t got = fread (magic, 1, 2, f);
    int format = -1;
    if (got == 2 && magic[0] == '>') format = GATB_TEXT_FASTA; else if (got == 2 && magic[0] == '@') format = GATB_TEXT_FASTQ;
    if (format < 0) { fclose (f); return false; }                          /* gzip (1f 8b), empty, binary bank, album ... */
    fseek (f, 0, SEEK_SET);
    const size_t BATCH = 256u << 20;
    std::vector<char> buf (BATCH + (8u << 20));
    size_t have = 0; bool eof = false;
    check (ctx, gatb_gpu_reads_begin (ctx, 0), "gatb_gpu_reads_begin");
    while (!eof || have)
    {
        if (!eof) { const size_t r = fread (buf.data () + have, 1, BATCH - have, f); have += r; if (r == 0 || have < BATCH) eof = true; }
        if (have == 0) break;
        size_t cut = have;
        if (!eof)
        {   /* last record boundary inside the batch: "\n>" (FASTA) / a line starting with '@' whose line after next starts with '+' (FASTQ) */
            cut = 0;
            for (size_t i = have - 1; i > 0; i--)
            {
                if (buf[i - 1] != '\n') continue;
                if (format == GATB_TEXT_FASTA) { if (buf[i] == '>') { cut = i; break; } continue; }
                if (buf[i] != '@') continue;
                const char* l1 = (const char*) memchr (buf.data () + i, '\n', have - i);
                if (!l1) continue;
                const char* l2 = (const char*) memchr (l1 + 1, '\n', buf.data () + have - (l1 + 1));
                if (!l2 || l2 + 1 >= buf.data () + have) continue;
                if (l2[1] == '+') { cut = i; break; }
            }
            if (cut == 0) { fclose (f); throw system::Exception ("GPU k-mer counting: no record boundary inside a %zu-byte batch of %s", have, path.c_str ()); }
        }
        check (ctx, gatb_gpu_reads_push_text (ctx, buf.data (), cut, format), "gatb_gpu_reads_push_text");
        if (progress) progress->inc (cut);
        memmove (buf.data (), buf.data () + cut, have - cut);
        have -= cut;
    }
    fclose (f);
    uint64_t info[6];
    check (ctx, gatb_gpu_reads_info (ctx, info), "gatb_gpu_reads_info");
    stats.sequencesNb = info[0]; stats.sequencesTotalLength = info[1];
    stats.sequencesMinLength = info[0] ? info[2] : ~0; stats.sequencesMaxLength = info[3];
    double sq; memcpy (&sq, &info[4], 8); stats.sequencesTotalLengthSquare = (u_int64_t) sq;
    return true;
}

/* ---- stage 2: one partition replayed through a processor clone (one command per partition, like PartitionsCommand) ---- */
template<size_t span>
class ReplayCommand : public tools::dp::ICommand, public system::SmartPointer
{
public:
    typedef typename Kmer<span>::Type Type;
    ReplayCommand (ICountProcessor<span>* proc, const gatb_gpu_result* res, size_t pass, size_t part, size_t nbParts, size_t cacheSize)
        : _proc(proc), _res(res), _pass(pass), _part(part), _nbParts(nbParts), _cacheSize(cacheSize) {}
    void execute ()
    {
        _proc->beginPart (_pass, _part, _cacheSize, "gpu");
        const uint64_t key = (uint64_t)_pass * _nbParts + _part;
        const uint64_t a = _res->part_offsets[key], b = _res->part_offsets[key + 1];
        CountVector cv (1);
        for (uint64_t i = a; i < b; i++)
        {
            const CountNumber c = (CountNumber)_res->counts[i];
            cv[0] = c;
            _proc->process (_part, make_kmer<Type> (_res->kmers_lo[i], _res->kmers_hi ? _res->kmers_hi[i] : 0), cv, c);
        }
        _proc->endPart (_pass, _part);
    }
private:
    ICountProcessor<span>* _proc; const gatb_gpu_result* _res; size_t _pass, _part, _nbParts, _cacheSize;
};

} /* namespace gpu_dsk */

#define GATB_GPU_SPECIALIZE(SPAN)                                                                                                   \
template<> void SortingCountAlgorithm<SPAN>::fillPartitions (size_t pass, Iterator<Sequence>* itSeq, PartiInfo<5>& pInfo)            \
{                                                                                                                                    \
    TIME_INFO (getTimeInfo(), "fill_partitions");                                                                                    \
    if (_config._solidityKind != KMER_SOLIDITY_SUM)                                                                                  \
        throw Exception ("GPU k-mer counting serves single-bank (sum) solidity only; solidity kind %d asked", (int)_config._solidityKind); \
    if (_config._minimizerType != 0)  throw Exception ("GPU k-mer counting: -minimizer-type 1 (frequency order) is not supported");  \
    /* execute() asks the super-k-mer storage for its file statistics and deletes it (SortingCountAlgorithm.cpp:711-721) */          \
    _tmpStorageName_superK = getInput()->getStr(STR_URI_OUTPUT_TMP) + "/" + System::file().getTemporaryFilename("superK_partitions"); \
    if (_superKstorage != 0)  { delete _superKstorage;  _superKstorage = 0; }                                                        \
    _superKstorage = new SuperKmerBinFiles (_tmpStorageName_superK, "superKparts", _config._nb_partitions);                          \
    _superKstorage->flushFiles();  _superKstorage->closeFiles();                                                                     \
    _progress->setMessage (Stringify::format (progressFormat1, pass+1, _config._nb_passes));                                         \
    _progress->init ();                                                                                                              \
    gpu_dsk::Run& run = gpu_dsk::run_of (this);                                                                                      \
    if (pass == 0)                                                                                                                   \
    {                                                                                                                                \
        run.close ();                                                                                                                \
        run.ctx = gatb_gpu_create (gpu_dsk::device_id ());                                                                           \
        if (run.ctx == 0)  throw Exception ("GPU k-mer counting: %s", gatb_gpu_last_error (0));                                      \
        if (!gpu_dsk::push_plain_file (run.ctx, _bank, _bankStats, _progress))                                                       \
            gpu_dsk::push_bank (run.ctx, itSeq, _bankStats, _progress);                                                              \
        else  itSeq->finalize ();                                                                                                    \
        gatb_gpu_params p;  memset (&p, 0, sizeof(p));                                                                               \
        p.kmer_size = _config._kmerSize;  p.minimizer_size = _config._minim_size;                                                    \
        p.nb_partitions = _config._nb_partitions;  p.nb_passes = _config._nb_passes;                                                 \
        p.abundance_min = _config._abundance.empty() ? 1 : _config._abundance[0].getBegin();                                         \
        p.abundance_max = _config._abundance.empty() ? 0x7fffffff : _config._abundance[0].getEnd();                                  \
        p.histo_max = _config._abundanceUserNb > 0 ? 10000 : 10000;   /* device-side histogram is unused here: processors build theirs */ \
        p.emit_all = 1;           /* every distinct k-mer goes through ICountProcessor::process, like PartitionsCommand::insert */   \
        const uint64_t nbm = (uint64_t)1 << (2 * _config._minim_size);                                                               \
        std::vector<uint16_t> table (nbm);                                                                                           \
        for (uint64_t i = 0; i < nbm; i++)  table[i] = (uint16_t)(*_repartitor)(i);                                                  \
        gpu_dsk::check (run.ctx, gatb_gpu_reads_count (run.ctx, &p, table.data(), 0, &run.res), "gatb_gpu_reads_count");             \
        run.counted = true;                                                                                                          \
        run.replays_left = _config._nb_passes * _processors.size();                                                                  \
        _bankStats.kmersNbValid   = run.res.stats[GATB_STAT_KMERS_VALID];                                                            \
        _bankStats.kmersNbInvalid = run.res.stats[GATB_STAT_KMERS_INVALID];                                                          \
    }                                                                                                                                \
    else  { itSeq->finalize (); }                                                                                                    \
    if (!run.counted)  throw Exception ("GPU k-mer counting: pass %d without a counted run", (int)pass);                             \
    /* partition sizes of this pass (PartiInfo is what getNbCoresList and the statistics read) */                                    \
    for (size_t part = 0; part < _config._nb_partitions; part++)                                                                     \
    {                                                                                                                                \
        const uint64_t key = (uint64_t)pass * _config._nb_partitions + part;                                                         \
        const uint64_t nb = run.res.part_offsets[key+1] - run.res.part_offsets[key];                                                 \
        pInfo.incKmer (part, nb);  pInfo.incKxmer (part, nb);                                                                        \
    }                                                                                                                                \
    if (pass == 0)  pInfo.incSuperKmer_per_minimBin (0, 1, 0);                                                                       \
}                                                                                                                                    \
template<> void SortingCountAlgorithm<SPAN>::fillSolidKmers_aux (ICountProcessor<SPAN>* processor, size_t pass, PartiInfo<5>& pInfo) \
{                                                                                                                                    \
    _progress->setMessage (Stringify::format (progressFormat2, pass+1, _config._nb_passes));                                         \
    gpu_dsk::Run& run = gpu_dsk::run_of (this);                                                                                      \
    if (!run.counted)  throw Exception ("GPU k-mer counting: fillSolidKmers without a counted run");                                 \
    const size_t cacheSize = 200*1000;                                                                                               \
    size_t p = 0;                                                                                                                    \
    while (p < _config._nb_partitions)                                                                                               \
    {                                                                                                                                \
        /* groups of partitions replayed in parallel, one clone each, exactly like the reference's dispatchCommands loop */          \
        const size_t group = std::max ((size_t)1, std::min ((size_t)_config._nb_partitions_in_parallel, (size_t)(_config._nb_partitions - p))); \
        std::vector<ICommand*> cmds;  std::vector<CountProcessor*> clones;                                                           \
        for (size_t j = 0; j < group; j++, p++)                                                                                      \
        {                                                                                                                            \
            CountProcessor* clone = processor->clone ();  clone->use ();  clones.push_back (clone);                                  \
            cmds.push_back (new gpu_dsk::ReplayCommand<SPAN> (clone, &run.res, pass, p, _config._nb_partitions, cacheSize));         \
        }                                                                                                                            \
        getDispatcher()->dispatchCommands (cmds, 0);                                                                                 \
        processor->finishClones (clones);                                                                                            \
        for (size_t i = 0; i < clones.size(); i++)  { clones[i]->forget (); }                                                        \
    }                                                                                                                                \
    if (run.replays_left > 0 && --run.replays_left == 0)  run.close ();     /* last processor of the last pass: release the device */ \
    if (_superKstorage != 0)  _superKstorage->closeFiles ();                                                                         \
}

GATB_GPU_SPECIALIZE (32)
GATB_GPU_SPECIALIZE (64)

/* the explicit instantiations the cmake-generated TemplateSpecialization files would have made (for the two spans we serve) */
template class SortingCountAlgorithm<32>;
template class SortingCountAlgorithm<64>;
template class PartitionsCommand<32>;
template class PartitionsCommand<64>;
template class PartitionsByHashCommand<32>;
template class PartitionsByHashCommand<64>;
template class PartitionsByVectorCommand<32>;
template class PartitionsByVectorCommand<64>;

}}}}
