// CPU check of the chunked record decoder of the counting kernel (gatb_core_b200/csrc/k2_decode.cuh) against a
// nucleotide-by-nucleotide canonical k-mer.
#include "k2_decode.cuh"
#include <stdio.h>
#include <vector>
static uint64_t st = 0x9E3779B97F4A7C15ULL;
static uint64_t rnd () { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; }
int main ()
{
    long checked = 0;
    for (int k = 2; k <= 31; k++)
        for (int rep = 0; rep < 400; rep++)
        {
            const int maxlen = 58 - k + 1 < 28 ? 58 - k + 1 : 28;
            const int len = 1 + (int)(rnd () % maxlen), nn = k + len - 1;
            uint8_t nt[64];
            for (int i = 0; i < 64; i++) nt[i] = (rep % 7 == 0) ? (uint8_t)((i / 2) & 1 ? 2 : 0) : (uint8_t)(rnd () & 3);
            uint32_t r[4] = {0, 0, 0, 0};
            for (int i = 0; i < nn; i++) r[i / 16] |= (uint32_t)nt[i] << (2 * (i % 16));
            for (int c = 0; 4 * c < len; c++)
            {
                K2Chunk C; k2_chunk_begin (C, r[0], r[1], r[2], r[3], c, k);
                uint32_t lo[4], hi[4];
                k2_chunk_kmer<0> (C, lo[0], hi[0]); k2_chunk_kmer<1> (C, lo[1], hi[1]);
                k2_chunk_kmer<2> (C, lo[2], hi[2]); k2_chunk_kmer<3> (C, lo[3], hi[3]);
                for (int i = 0; i < 4 && 4 * c + i < len; i++)
                {
                    const int j = 4 * c + i;
                    uint64_t f = 0, rc = 0;
                    for (int p = 0; p < k; p++) { uint64_t n = nt[j + p]; f = (f << 2) | n; rc |= (n ^ 2) << (2 * p); }
                    const uint64_t want = f < rc ? f : rc, got = ((uint64_t)hi[i] << 32) | lo[i];
                    if (want != got) { printf ("k=%d len=%d kmer %d: want %016llx got %016llx\n", k, len, j, (unsigned long long)want, (unsigned long long)got); return 1; }
                    checked++;
                }
            }
        }
    // ---- 32 <= k <= 63: 32-byte records, 128-bit values ----
    long checked2 = 0;
    for (int k = 32; k <= 63; k++)
        for (int rep = 0; rep < 300; rep++)
        {
            const int maxlen = 122 - k + 1 < 60 ? 122 - k + 1 : 60;
            const int len = 1 + (int)(rnd () % maxlen), nn = k + len - 1;
            uint8_t nt[128];
            for (int i = 0; i < 128; i++) nt[i] = (rep % 7 == 0) ? (uint8_t)((i / 2) & 1 ? 2 : 0) : (uint8_t)(rnd () & 3);
            uint32_t r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int i = 0; i < nn; i++) r[i / 16] |= (uint32_t)nt[i] << (2 * (i % 16));
            for (int c = 0; 4 * c < len; c++)
            {
                K2Chunk2 C; k2_chunk2_begin (C, r, c, k);
                uint32_t v[4][4];
                k2_chunk2_kmer<0> (C, v[0]); k2_chunk2_kmer<1> (C, v[1]); k2_chunk2_kmer<2> (C, v[2]); k2_chunk2_kmer<3> (C, v[3]);
                for (int i = 0; i < 4 && 4 * c + i < len; i++)
                {
                    const int j = 4 * c + i;
                    unsigned __int128 f = 0, rc = 0;
                    for (int p = 0; p < k; p++) { unsigned __int128 n = nt[j + p]; f = (f << 2) | n; rc |= (n ^ 2) << (2 * p); }
                    const unsigned __int128 want = f < rc ? f : rc;
                    unsigned __int128 got = 0;
                    for (int w = 3; w >= 0; w--) got = (got << 32) | v[i][w];
                    if (want != got) { printf ("k=%d len=%d kmer %d: 128-bit mismatch\n", k, len, j); return 1; }
                    checked2++;
                }
            }
        }
    printf ("k2 decode ok (%ld + %ld k-mers)\n", checked, checked2);
    return 0;
}
