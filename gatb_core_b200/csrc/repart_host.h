// repart_host.h -- host arithmetic of Repartitor::computeDistrib (/root/reference/gatb-core/src/gatb/kmer/impl/PartiInfo.cpp:48-106):
// the minimizer bins sorted by decreasing estimated size (std::sort with the reference's comparator on the same sequence: ties fall
// the way libstdc++ lets them fall there), each given to the partition with the least space used so far (std::priority_queue with the
// reference's comparator).  Plain C++ so that tests/cpp/test_repart_core.cpp checks it on the CPU against the reference itself.
#pragma once
#include <stdint.h>
#include <algorithm>
#include <queue>
#include <utility>
#include <vector>

namespace repart_host
{
typedef std::pair<uint64_t, uint64_t> ipair;                                          // bin size, bin number (PartiInfo.hpp:347)
struct itriple { uint64_t first, second, third; };                                    // partition, space used, bins held (:349-358)
struct comp_bins { bool operator() (ipair l, ipair r) { return l.first > r.first; } };                  // :360-362
struct comp_space { bool operator() (itriple l, itriple r) { return l.second > r.second; } };           // :368-370
}

static inline void repartition_distribute (const std::vector<unsigned long long>& kx, int nb_partitions, uint16_t* table)
{
    using namespace repart_host;
    std::vector<ipair> bins;
    for (uint64_t i = 0; i < kx.size (); i++) bins.push_back (ipair (kx[i], i));
    std::priority_queue<itriple, std::vector<itriple>, comp_space> pq;
    for (int j = 0; j < nb_partitions; j++) { itriple t = { (uint64_t)j, 0, 0 }; pq.push (t); }
    std::sort (bins.begin (), bins.end (), comp_bins ());
    for (uint64_t c = 0; c < bins.size (); c++)
    {
        itriple smallest = pq.top (); pq.pop ();
        table[bins[c].second] = (uint16_t)smallest.first;
        smallest.second += bins[c].first; smallest.third++;
        pq.push (smallest);
    }
}
