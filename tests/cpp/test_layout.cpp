// CPU check of the coarse-bin region layout shared by the partition kernels, the fine split and the multi-GPU exchange
// (gatb_core_b200/csrc/kernels.h, coarse_index): it must be a bijection of (bin, slot) onto [0, bins * cap), and a region
// whose bins hold at most c records each must only touch its first ceil(c / COARSE_BLK) rounds (what the exchange sends).
#define __host__
#define __device__
#define __forceinline__ inline
#include <stdint.h>
#include <stdio.h>
#include <vector>
typedef int cudaError_t; typedef void* cudaStream_t;       // kernels.h only needs the names for its launcher prototypes
struct uint4 { uint32_t x, y, z, w; }; struct uint2 { uint32_t x, y; };
#define KERNELS_H_NO_CUDA_RUNTIME
#include "kernels.h"

int main ()
{
    const uint32_t nbs[] = { 1, 3, 64, 1000 }, caps[] = { COARSE_BLK, 2 * COARSE_BLK, 7 * COARSE_BLK };
    for (uint32_t nb : nbs) for (uint32_t cap : caps)
    {
        std::vector<uint8_t> seen ((size_t)nb * cap, 0);
        for (uint32_t b = 0; b < nb; b++) for (uint32_t s = 0; s < cap; s++)
        {
            const uint64_t i = coarse_index (b, s, nb);
            if (i >= seen.size () || seen[i]) { printf ("not a bijection: nb=%u cap=%u bin=%u slot=%u -> %llu\n", nb, cap, b, s, (unsigned long long)i); return 1; }
            seen[i] = 1;
        }
        for (uint32_t c = 1; c <= cap; c++)
        {
            const uint64_t used = (uint64_t)((c + COARSE_BLK - 1) / COARSE_BLK) * nb * COARSE_BLK;      // records the exchange moves
            for (uint32_t b = 0; b < nb; b += (nb > 8 ? nb / 8 : 1))
                if (coarse_index (b, c - 1, nb) >= used) { printf ("slot %u of bin %u lies beyond the used prefix\n", c - 1, b); return 1; }
        }
    }
    printf ("coarse layout ok (block of %d records)\n", (int)COARSE_BLK);
    return 0;
}
