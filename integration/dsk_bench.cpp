/* integration/dsk_bench.cpp -- times SortingCountAlgorithm<span>::execute() through GATB-core's public API (default processor
 * chain: histogram -> solidity -> dump into the storage), FASTA/FASTQ file in, storage out.  Linked twice like dsk_tool
 * (integration/Makefile): dsk_bench_gpu = GPU path, dsk_bench_cpu = the reference's own instantiation.  bench.py runs both on the
 * same file for its `e2e_api` and `cpu_baseline` figures.  Prints ONE line of JSON on stdout.
 *   dsk_bench -in reads.fa -kmer-size 31 -abundance-min 2 -out /tmp/x -out-tmp /tmp -storage-type file -nb-cores N */
#include <gatb/gatb_core.hpp>
#include <chrono>
using namespace std;

static string g_json;

template<size_t span>  struct MainLoop  {  void operator () (IProperties* options)
{
    SortingCountAlgorithm<span> algo (options);
    const auto t0 = chrono::steady_clock::now ();
    algo.execute ();
    const double sec = chrono::duration<double> (chrono::steady_clock::now () - t0).count ();
    IProperties* info = algo.getInfo ();
    const char* keys[] = { "kmers_nb_valid", "kmers_nb_distinct", "kmers_nb_solid", "seq_number", "fill_partitions", "fill_solid_kmers" };
    char buf[256];
    snprintf (buf, sizeof(buf), "{\"seconds\": %.6f, \"nb_partitions\": %d, \"nb_passes\": %d", sec, (int)algo.getConfig()._nb_partitions, (int)algo.getConfig()._nb_passes);
    g_json = buf;
    for (size_t i = 0; i < sizeof(keys)/sizeof(keys[0]); i++)
        if (info->get (keys[i]))  { g_json += string (", \"") + keys[i] + "\": " + info->getStr (keys[i]); }
    g_json += "}";
}};

int main (int argc, char* argv[])
{
    IOptionsParser* parser = SortingCountAlgorithm<>::getOptionsParser ();  LOCAL (parser);
    if (OptionsParser* p = dynamic_cast<OptionsParser*> (parser))
    {
        p->push_back (new OptionOneParam (STR_NB_CORES, "number of cores", false, "0"));
        p->push_back (new OptionOneParam (STR_VERBOSE,  "verbosity level", false, "0"));
    }
    try
    {
        IProperties* options = parser->parse (argc, argv);
        Integer::apply<MainLoop,IProperties*> (options->getInt (STR_KMER_SIZE), options);
    }
    catch (OptionFailure& e)  { return e.displayErrors (cerr); }
    catch (Exception& e)      { cerr << "EXCEPTION: " << e.getMessage() << endl;  return EXIT_FAILURE; }
    cout << g_json << endl;
    return EXIT_SUCCESS;
}
