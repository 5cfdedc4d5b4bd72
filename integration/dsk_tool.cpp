/* integration/dsk_tool.cpp -- a small DSK written ONLY against GATB-core's public API (gatb/gatb_core.hpp), in the style of
 * the reference's examples/kmer/kmer12.cpp.  It is compiled once and linked twice (integration/Makefile):
 *     dsk_tool_cpu   with the reference's own instantiation of SortingCountAlgorithm<span>
 *     dsk_tool_gpu   with integration/SortingCountAlgorithmGPU.cpp + libgatb_b200.so
 * Same source, same options, byte-identical dumps expected: tests/test_integration.py compares them.
 *
 *   dsk_tool -in reads.fa -kmer-size 31 -abundance-min 2 -out /tmp/x -storage-type file [any option of the DSK parser]
 *
 * Run 1 (default chain: histogram -> solidity -> dump, SortingCountAlgorithm(IProperties*) + execute()):
 *     <out>.solid.txt   partition by partition, "p kmer abundance" of Partition<Count>* getSolidCounts()
 *     <out>.histo.txt   the histogram collection the default chain saved into the storage
 *     <out>.info.txt    kmers_nb_valid / kmers_nb_invalid / kmers_nb_distinct / kmers_nb_solid / nb_partitions / nb_passes
 * With -storage-type hdf5 (the default) the <out>.h5 file is then re-opened from disk and dumped:
 *     <out>.h5dump.txt  dsk/solid/<p> datasets, dsk attributes, histogram/histogram, a checksum of minimizers/minimRepart
 * Run 2 (a custom ICountProcessor added with addProcessor(), the kmer12 pattern): every distinct k-mer with its count
 *     <out>.all.txt     "key kmer count", key = pass * nb_partitions + partition, in the order process() was called
 */
#include <gatb/gatb_core.hpp>
#include <fstream>
#include <map>
#include <pthread.h>
using namespace std;

struct Sink { pthread_mutex_t mtx; map<size_t, string> text; Sink () { pthread_mutex_init (&mtx, 0); } };

template<size_t span>
class DumpAllProcessor : public CountProcessorAbstract<span>
{
public:
    typedef typename Kmer<span>::Type Type;
    DumpAllProcessor (Sink* sink, size_t kmerSize, size_t nbParts = 1) : _sink(sink), _kmerSize(kmerSize), _nbParts(nbParts), _key(0), _total(0) {}
    CountProcessorAbstract<span>* clone ()  { return new DumpAllProcessor (_sink, _kmerSize, _nbParts); }
    void begin (const Configuration& config)  { _nbParts = config._nb_partitions; }
    void beginPart (size_t passId, size_t partId, size_t cacheSize, const char* name)  { _key = passId * _nbParts + partId; _local.clear(); }
    void endPart (size_t passId, size_t partId)
    {
        pthread_mutex_lock (&_sink->mtx);  _sink->text[_key] += _local;  pthread_mutex_unlock (&_sink->mtx);
    }
    bool process (size_t partId, const Type& kmer, const CountVector& count, CountNumber sum)
    {
        char buf[64];  snprintf (buf, sizeof(buf), " %d\n", (int)count[0]);
        _local += to_string (_key) + " " + kmer.toString (_kmerSize) + buf;
        _total++;
        return true;
    }
    void finishClones (vector<ICountProcessor<span>*>& clones)
    {
        for (size_t i=0; i<clones.size(); i++)
            if (DumpAllProcessor* c = dynamic_cast<DumpAllProcessor*> (clones[i]))  { _total += c->_total; }
    }
    u_int64_t getTotal () const { return _total; }
private:
    Sink* _sink; size_t _kmerSize, _nbParts, _key; u_int64_t _total; string _local;
};

template<size_t span>  struct MainLoop  {  void operator () (IProperties* options)
{
    typedef typename Kmer<span>::Count Count;
    const string out = options->getStr (STR_URI_OUTPUT);
    const size_t k   = options->getInt (STR_KMER_SIZE);

    /* ---- run 1: the default processor chain ---- */
    {
        SortingCountAlgorithm<span> algo (options);
        algo.execute ();
        ofstream fs ((out + ".solid.txt").c_str());
        Partition<Count>* solid = algo.getSolidCounts ();
        u_int64_t nbSolid = 0;
        for (size_t p = 0; p < solid->size(); p++)
        {
            Iterator<Count>* it = (*solid)[p].iterator ();  LOCAL (it);
            for (it->first(); !it->isDone(); it->next())  { fs << p << " " << it->item().value.toString (k) << " " << it->item().abundance << "\n";  nbSolid++; }
        }
        ofstream fh ((out + ".histo.txt").c_str());
        Iterable<gatb::core::tools::misc::IHistogram::Entry>& histo = algo.getStorage()->getGroup ("histogram").template getCollection<gatb::core::tools::misc::IHistogram::Entry> ("histogram");
        Iterator<gatb::core::tools::misc::IHistogram::Entry>* ih = histo.iterator ();  LOCAL (ih);
        for (ih->first(); !ih->isDone(); ih->next())  { fh << ih->item().index << " " << ih->item().abundance << "\n"; }
        ofstream fi ((out + ".info.txt").c_str());
        const char* keys[] = { "kmers_nb_valid", "kmers_nb_invalid", "kmers_nb_distinct", "kmers_nb_solid", "kmers_nb_weak" };
        for (size_t i = 0; i < sizeof(keys)/sizeof(keys[0]); i++)
            if (algo.getInfo()->get (keys[i]))  { fi << keys[i] << " " << algo.getInfo()->getStr (keys[i]) << "\n"; }
        fi << "nb_partitions " << algo.getConfig()._nb_partitions << "\nnb_passes " << algo.getConfig()._nb_passes << "\n";
        fi << "solid_iterated " << nbSolid << "\n";
        cout << "run 1 (default chain): " << nbSolid << " solid k-mers in " << solid->size() << " partitions" << endl;
    }
    /* ---- the .h5 file itself, re-opened from disk the way Graph::load / Graph::create -in x.h5 consume it
     *      (debruijn/impl/Graph.cpp:151-235, 778-803): dsk/solid/<p> datasets + group attributes, histogram, minimizers/minimRepart ---- */
    if (options->getStr (STR_STORAGE_TYPE) == "hdf5")
    {
        Storage* storage = StorageFactory (STORAGE_HDF5).load (out);  LOCAL (storage);
        ofstream fd ((out + ".h5dump.txt").c_str());
        Group& dsk = storage->getGroup ("dsk");
        fd << "dsk.kmer_size " << dsk.getProperty ("kmer_size") << "\n";
        Partition<Count>& solid = dsk.getPartition<Count> ("solid");
        fd << "dsk/solid partitions " << solid.size() << " items " << solid.getNbItems() << "\n";
        for (size_t p = 0; p < solid.size(); p++)
        {
            Iterator<Count>* it = solid[p].iterator ();  LOCAL (it);
            for (it->first(); !it->isDone(); it->next())  { fd << p << " " << it->item().value.toString (k) << " " << it->item().abundance << "\n"; }
        }
        Iterable<gatb::core::tools::misc::IHistogram::Entry>& histo = storage->getGroup ("histogram").getCollection<gatb::core::tools::misc::IHistogram::Entry> ("histogram");
        Iterator<gatb::core::tools::misc::IHistogram::Entry>* ih = histo.iterator ();  LOCAL (ih);
        for (ih->first(); !ih->isDone(); ih->next())  { if (ih->item().abundance) fd << "histogram " << ih->item().index << " " << ih->item().abundance << "\n"; }
        Repartitor repart;  repart.load (storage->getGroup ("minimizers"));
        u_int64_t sum = 0;  const u_int64_t nbm = (u_int64_t)1 << (2 * options->getInt (STR_MINIMIZER_SIZE));
        for (u_int64_t i = 0; i < nbm; i++)  { sum = sum * 1000003ULL + repart (i); }
        fd << "minimRepart passes " << repart.getNbPasses() << " checksum " << sum << "\n";
        cout << "re-opened " << out << ".h5: " << solid.getNbItems() << " solid k-mers in " << solid.size() << " datasets" << endl;
    }
    /* ---- run 2: a custom count processor (kmer12.cpp pattern) ---- */
    {
        Sink sink;
        SortingCountAlgorithm<span> algo (options);
        DumpAllProcessor<span>* proc = new DumpAllProcessor<span> (&sink, k);
        algo.addProcessor (proc);
        algo.execute ();
        ofstream fa ((out + ".all.txt").c_str());
        for (map<size_t,string>::iterator it = sink.text.begin(); it != sink.text.end(); ++it)  { fa << it->second; }
        cout << "run 2 (custom processor): " << proc->getTotal() << " distinct k-mers" << endl;
    }
}};

int main (int argc, char* argv[])
{
    IOptionsParser* parser = SortingCountAlgorithm<>::getOptionsParser ();  LOCAL (parser);
    if (OptionsParser* p = dynamic_cast<OptionsParser*> (parser))
    {   /* the general options a Tool adds to the DSK parser (tools/misc/impl/Tool.cpp) */
        p->push_back (new OptionOneParam (STR_NB_CORES, "number of cores", false, "0"));
        p->push_back (new OptionOneParam (STR_VERBOSE,  "verbosity level", false, "0"));
    }
    try
    {
        IProperties* options = parser->parse (argc, argv);
        Integer::apply<MainLoop,IProperties*> (options->getInt (STR_KMER_SIZE), options);
    }
    catch (OptionFailure& e)  { return e.displayErrors (cout); }
    catch (Exception& e)      { cerr << "EXCEPTION: " << e.getMessage() << endl;  return EXIT_FAILURE; }
    return EXIT_SUCCESS;
}
