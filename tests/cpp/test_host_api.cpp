// C++ tests of the host-side mirror of GATB's API (gatb_core_b200/host/gatb/gatb_core_b200.hpp).  They read like the
// reference's own cppunit tests (test/unit/src/kmer/TestKmer.cpp, TestDSK.cpp) because the API is the same.
//   ./test_host_api cpu   -- k-mer models only (no device needed)
//   ./test_host_api gpu   -- + SortingCountAlgorithm / BloomBuilder through the C ABI on cuda:0
#include <gatb/gatb_core_b200.hpp>
#include <iostream>
#include <set>

using namespace gatb::core;
using namespace gatb::core::kmer;
using namespace gatb::core::kmer::impl;
using namespace gatb::core::bank;

static int failures = 0;
#define CHECK(cond) do { if (!(cond)) { std::cerr << "FAILED " << __FILE__ << ":" << __LINE__ << "  " #cond << std::endl; failures++; } } while (0)

template<size_t span> struct ModelChecks
{
    template<class Model> struct Collect { std::vector<uint64_t>& v; Collect (std::vector<uint64_t>& v) : v(v) {} void operator() (const typename Model::Kmer& k, size_t) { v.push_back (k.value ().getVal ()); } };
    static void run ()
    {
        // TestKmer.cpp:141-190
        const char* seq = "CATTGATAGTGG";
        long direct[] = {18, 10, 43, 44, 50, 8, 35, 14, 59, 47}, both[] = {11, 2, 16, 36, 9, 8, 24, 6, 17, 20};
        { typename Kmer<span>::ModelDirect model (3); std::vector<uint64_t> v; model.iterate (seq, strlen (seq), Collect<typename Kmer<span>::ModelDirect> (v));
          CHECK (v.size () == 10); for (size_t i = 0; i < v.size (); i++) CHECK ((long)v[i] == direct[i]); }
        { typename Kmer<span>::ModelCanonical model (3); std::vector<uint64_t> v; model.iterate (seq, strlen (seq), Collect<typename Kmer<span>::ModelCanonical> (v));
          CHECK (v.size () == 10); for (size_t i = 0; i < v.size (); i++) CHECK ((long)v[i] == both[i]); }
        // TestKmer.cpp:233-261
        { typename Kmer<span>::ModelCanonical model (5); std::vector<typename Kmer<span>::ModelCanonical::Kmer> kmers;
          uint64_t check[] = {0x61, 0x187, 0x21c, 0x72, 0x1c9, 0x1c9, 0x9c, 0x9c, 0x127, 0x49, 0xb8};
          model.build ("ACTACGATCGATGTA", kmers); CHECK (kmers.size () == 11);
          for (size_t i = 0; i < kmers.size (); i++) { CHECK (kmers[i].value ().getVal () == check[i]); CHECK (model.getKmer ("ACTACGATCGATGTA", i).value ().getVal () == check[i]); } }
        // the model refuses a k-mer size that does not fit the span (Model.hpp:402-407)
        bool thrown = false; try { typename Kmer<span>::ModelCanonical m (span); } catch (system::Exception&) { thrown = true; } CHECK (thrown);
    }
};

static void minimizer3 ()
{   // TestKmer.cpp:434-502
    const char* seq = "ATGTCTGAAGTGACCTAACATTGCAGTGTGTT";
    typedef Kmer<>::ModelCanonical ModelCanonical; typedef Kmer<>::ModelMinimizer<ModelCanonical> ModelMinimizer;
    ModelMinimizer model (15, 7);
    const ModelCanonical& modelMini = model.getMmersModel ();
    struct { const char* kmer; const char* minimizer; int position; bool changed; } table[] = {
        {"ATGTCTGAAGTGACC", "AAGTGAC", 7, true }, {"AGGTCACTTCAGACA", "AAGTGAC", 6, false}, {"TAGGTCACTTCAGAC", "AAGTGAC", 5, false},
        {"TCTGAAGTGACCTAA", "AAGTGAC", 4, false}, {"CTGAAGTGACCTAAC", "AAGTGAC", 3, false}, {"TGAAGTGACCTAACA", "AAGTGAC", 2, false},
        {"ATGTTAGGTCACTTC", "AAGTGAC", 1, false}, {"AATGTTAGGTCACTT", "AATGTTA", 8, true }, {"AGTGACCTAACATTG", "AACATTG", 8, true },
        {"GCAATGTTAGGTCAC", "AACATTG", 7, false}, {"TGACCTAACATTGCA", "AACATTG", 6, false}, {"CTGCAATGTTAGGTC", "AACATTG", 5, false},
        {"ACCTAACATTGCAGT", "AACATTG", 4, false}, {"CACTGCAATGTTAGG", "AACATTG", 3, false}, {"ACACTGCAATGTTAG", "AACATTG", 2, false},
        {"CACACTGCAATGTTA", "AACATTG", 1, false}, {"AACATTGCAGTGTGT", "AACATTG", 0, false}, {"AACACACTGCAATGT", "AACACAC", 8, true } };
    ModelMinimizer::Kmer kmer = model.codeSeed (seq);
    for (size_t idx = 0; idx < sizeof(table) / sizeof(table[0]); idx++)
    {
        if (idx) kmer = model.codeSeedRight (kmer, seq[15 + idx - 1]);
        CHECK (model.toString (kmer.value ()) == table[idx].kmer);
        CHECK (modelMini.toString (kmer.minimizer ().value ()) == table[idx].minimizer);
        CHECK (kmer.position () == table[idx].position);
        CHECK (kmer.hasChanged () == table[idx].changed);
        CHECK (model.getMinimizerString (kmer.value ()) == table[idx].minimizer);
    }
}

static void badchar ()
{   // TestKmer.cpp:542-569
    typedef Kmer<>::ModelDirect ModelDirect;
    ModelDirect model (11);
    const char* seq = "ACGNCNTGCTAGCTATTTAGCTTTAGANAGTAGATGACGCNC";
    bool valid[] = {0,0,0,0,0,0, 1,1,1,1,1,1,1,1,1,1,1, 0,0,0,0,0,0,0,0,0,0,0, 1,1, 0,0};
    struct F { const ModelDirect& m; const char* seq; bool* valid; void operator() (const ModelDirect::Kmer& k, size_t idx)
               { CHECK (k.isValid () == valid[idx]); std::string s (seq + idx, 11); for (size_t i = 0; i < s.size (); i++) if (s[i] == 'N') s[i] = 'G'; CHECK (m.toString (k.value ()) == s); } };
    F f = { model, seq, valid };
    model.iterate (seq, strlen (seq), f);
}

static void repartitor_roundtrip ()
{   // PartiInfo.cpp:228-295 byte stream
    Repartitor r (7, 4, 1); std::vector<uint16_t> t (256); for (size_t i = 0; i < t.size (); i++) t[i] = (uint16_t)(i % 7); r.setTable (t);
    std::stringstream ss; r.save (ss);
    CHECK (ss.str ().size () == 2 + 8 + 2 + 256 * 2 + 1 + 4);
    Repartitor q; q.load (ss);
    CHECK (q.getNbPartitions () == 7 && q.getTable () == t && q (10) == 3);
}

// ---- GPU part --------------------------------------------------------------------------------------------------
template<size_t span> static void DSK_check2 ()
{   // TestDSK.cpp:244-341: exact solid 31-mers + checksum, for Kmer<32> and Kmer<64>
    typedef typename Kmer<span>::Type Type; typedef typename Kmer<span>::Count Count;
    Configuration config; config._kmerSize = 31; config._abundance[0] = tools::misc::CountRange (1, 0x7fffffff);
    SortingCountAlgorithm<span> sortingCount (new BankStrings ("GATCGATTCTTAGCACGTCCCCCCCTACACCCAAT", (const char*)0), config, 0);
    sortingCount.execute ();
    std::set<uint64_t> ok; ok.insert (0x1CA68D1E55561150ULL); ok.insert (0x09CA68D1E5556115ULL); ok.insert (0x2729A34795558454ULL); ok.insert (0x32729A3479555845ULL); ok.insert (0x0AFEE3FFF1ED8309ULL);
    std::vector<Count>& solid = (*sortingCount.getSolidCounts ())[0];
    uint64_t checksum = 0; std::set<uint64_t> seen;
    for (size_t i = 0; i < solid.size (); i++) { CHECK (ok.count (solid[i].value.getVal ()) == 1); CHECK (solid[i].value.hi () == 0); seen.insert (solid[i].value.getVal ()); checksum += solid[i].value.getVal (); if (i) CHECK (solid[i-1].value < solid[i].value); }
    CHECK (checksum == 0x8b0c176c3b43d207ULL); CHECK (seen.size () == ok.size ());
    CHECK (sortingCount.getInfo ().getInt ("kmers_nb_solid") == 5);
}

static void DSK_check1 ()
{   // TestDSK.cpp:147-241 (first block)
    const char* s1 = "GATCCTCCCCAGGCCCCTACACCCAAT";
    struct { int n; int k; int nks; int expected; } cases[] = { {1,27,1,1}, {1,26,1,2}, {1,27,2,0}, {1,26,2,0}, {2,27,2,1}, {2,26,2,2}, {2,27,3,0}, {3,26,3,2}, {3,27,4,0} };
    for (size_t c = 0; c < sizeof(cases) / sizeof(cases[0]); c++)
    {
        std::vector<std::string> seqs (cases[c].n, s1);
        Configuration config; config._kmerSize = cases[c].k; config._minim_size = 8; config._abundance[0] = tools::misc::CountRange (cases[c].nks, 0x7fffffff);
        SortingCountAlgorithm<> dsk (new BankStrings (seqs), config, 0);
        dsk.execute ();
        CHECK (dsk.getInfo ().getInt ("kmers_nb_solid") == cases[c].expected);
    }
}

// a custom count processor (like examples/kmer/kmer12.cpp): must see EVERY distinct k-mer, ascending inside a partition
template<size_t span> class Spy : public CountProcessorAbstract<span>
{
public:
    typedef typename Kmer<span>::Type Type;
    struct Shared { uint64_t total, sum, orderViolations, parts; Shared () : total(0), sum(0), orderViolations(0), parts(0) {} };
    Spy (Shared* s) : _s(s), _first(true) {}
    ICountProcessor<span>* clone () { return new Spy (_s); }
    void beginPart (size_t, size_t, size_t, const char*) { _first = true; _s->parts++; }
    bool process (size_t, const Type& kmer, const CountVector& count, CountNumber sum)
    { CHECK (count.size () == 1 && count[0] == sum); _s->total++; _s->sum += sum; if (!_first && !(_prev < kmer)) _s->orderViolations++; _prev = kmer; _first = false; return true; }
private:
    Shared* _s; Type _prev; bool _first;
};

static void custom_processor_and_partitions ()
{
    std::vector<std::string> seqs; uint64_t x = 88172645463325252ULL;
    for (int i = 0; i < 400; i++) { std::string s; for (int j = 0; j < 120; j++) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; s += "ACGT"[x & 3]; } seqs.push_back (s); seqs.push_back (s); }
    Configuration config; config._kmerSize = 31; config._minim_size = 8; config._nb_partitions = 4;
    Repartitor* rep = new Repartitor (4, 8, 1); std::vector<uint16_t> t (1 << 16); for (size_t i = 0; i < t.size (); i++) t[i] = (uint16_t)((i * 2654435761u >> 7) % 4); rep->setTable (t);
    Spy<32>::Shared shared;
    std::vector<ICountProcessor<32>*> procs; procs.push_back (new Spy<32> (&shared));
    SortingCountAlgorithm<32> dsk (new BankStrings (seqs), config, rep, procs);
    dsk.execute ();
    CHECK (shared.parts == 4); CHECK (shared.orderViolations == 0);
    CHECK (shared.total == (uint64_t)dsk.getInfo ().getInt ("kmers_nb_distinct"));
    CHECK (shared.sum == (uint64_t)dsk.getInfo ().getInt ("kmers_nb_valid"));          // sum of abundances = k-mer occurrences
    CHECK (dsk.getInfo ().getInt ("kmers_nb_valid") == 800 * 90);
    // default chain on the same input: every k-mer occurs an even number of times -> all distinct k-mers are solid at abundance-min 2
    SortingCountAlgorithm<32> dsk2 (new BankStrings (seqs), config, rep);
    dsk2.execute ();
    uint64_t nsolid = 0; for (size_t p = 0; p < 4; p++) nsolid += (*dsk2.getSolidCounts ())[p].size ();
    CHECK (nsolid == shared.total); CHECK ((uint64_t)dsk2.getInfo ().getInt ("kmers_nb_solid") == nsolid);
    tools::misc::Histogram* h = dsk2.getHistogram (); uint64_t hs = 0; for (size_t i = 0; i <= h->getLength (); i++) hs += h->get (i);
    CHECK (hs == shared.total); CHECK (h->get (1) == 0);
    // Bloom of the solid k-mers, sized like BloomAlgorithm
    uint64_t size; size_t nbHash; BloomBuilder<32>::sizeFor (31, nsolid, size, nbHash);
    std::vector<Kmer<32>::Count> all; for (size_t p = 0; p < 4; p++) all.insert (all.end (), (*dsk2.getSolidCounts ())[p].begin (), (*dsk2.getSolidCounts ())[p].end ());
    uint64_t bits = 0; std::vector<uint8_t> bloom = BloomBuilder<32> (size, nbHash, 31, "neighbor").build (all, &bits);
    uint64_t ones = 0; for (size_t i = 0; i < bloom.size (); i++) ones += __builtin_popcount (bloom[i]);
    CHECK (bits == size); CHECK (ones > 0 && ones <= nbHash * nsolid);
}

int main (int argc, char** argv)
{
    std::string mode = argc > 1 ? argv[1] : "cpu";
    ModelChecks<32>::run (); ModelChecks<64>::run ();
    minimizer3 (); badchar (); repartitor_roundtrip ();
    if (mode == "gpu") { DSK_check2<32> (); DSK_check2<64> (); DSK_check1 (); custom_processor_and_partitions (); }
    else
    {   // without a device the algorithm must fail loudly, never fall back
        if (!gatb_gpu_create (0))
        {
            bool thrown = false; Configuration config;
            try { SortingCountAlgorithm<> dsk (new BankStrings ("ACGTACGTACGTACGTACGTACGTACGTACGTACGT", (const char*)0), config, 0); dsk.execute (); }
            catch (system::Exception& e) { thrown = true; std::cout << "expected failure without a device: " << e.getMessage () << std::endl; }
            CHECK (thrown);
        }
    }
    std::cout << (failures ? "FAILED" : "OK") << " (" << mode << ", " << failures << " failures)" << std::endl;
    return failures ? 1 : 0;
}
