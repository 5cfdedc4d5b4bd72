// The sampling pass of the Repartitor as the device runs it (k_repart_core.cuh, compiled for the host) + the host distribution
// (repart_host.h): reads one sequence per line from argv[1], writes the u16[4^m] table to argv[6].
// usage: test_repart_core seqs.txt k m nb_partitions nb_seqs_to_see out.bin      (tests/test_repartition.py compares with the reference)
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include <fstream>
#include "k_repart_core.cuh"
#include "repart_host.h"

int main (int argc, char** argv)
{
    if (argc < 7) return 2;
    const int k = atoi (argv[2]), m = atoi (argv[3]), nparts = atoi (argv[4]);
    const unsigned long long to_see = strtoull (argv[5], 0, 10);
    std::ifstream in (argv[1]);
    std::string line;
    std::vector<unsigned long long> kx ((size_t)1 << (2 * m), 0);
    unsigned long long seen = 0, nreads = 0;
    while (std::getline (in, line))
    {
        const char* s = line.data (); const int len = (int)line.size ();
        auto nuc = [&] (int i) -> uint32_t { return ((unsigned char)s[i] >> 1) & 3u; };                       // A=0 C=1 T=2 G=3 (Data.hpp:185)
        auto bad = [&] (int i) -> bool { const char c = s[i]; return !(c=='A'||c=='C'||c=='G'||c=='T'||c=='a'||c=='c'||c=='g'||c=='t'); };
        seen += krp_scan_read (nuc, bad, len, k, m, [&] (uint32_t mini, uint32_t n) { kx[mini] += n; });
        nreads++;
        if (seen > to_see) break;                                                                             // the iteration is cancelled between two reads
    }
    std::vector<uint16_t> table (kx.size (), 0);
    repartition_distribute (kx, nparts, table.data ());
    FILE* f = fopen (argv[6], "wb");
    fwrite (table.data (), 2, table.size (), f);
    fclose (f);
    printf ("%llu reads sampled, %llu super-k-mers\n", nreads, seen);
    return 0;
}
