// k_repart_core.cuh -- one read of the Repartitor's sampling pass (row f3 of SURVEY.md 8): super-k-mers in GATB's own minimizer order and,
// per super-k-mer, the number of kx-mers the reference charges to its minimizer.
//
// Replaces (paths relative to /root/reference/gatb-core/src/gatb/):
//   SampleRepart<span>::processSuperkmer        kmer/impl/RepartitionAlgorithm.cpp:157-243   (kx-mer accounting, _kx = 4)
//   Sequence2SuperKmer<span>::operator()        kmer/impl/Sequence2SuperKmer.hpp:81-159      (super-k-mer cuts)
//   ModelCanonical / ModelMinimizer first/next  kmer/impl/Model.hpp:857-884, 1082-1139, 1254-1287
// k <= 63 (Kmer<32> and Kmer<64>: the k-mer is kept in 128 bits, the super-k-mer capacity follows the span like Sequence2SuperKmer.hpp:147);
// m-mer key = min(m-mer, revcomp) with the "AA" rule applied arithmetically (common.cuh gatb_mmer_key).
// __host__ __device__: tests/cpp/test_repart_core.cpp runs the same code on the CPU against the oracle's restatement.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define KRP_HD __host__ __device__ __forceinline__
#else
#define KRP_HD inline
#endif

KRP_HD uint32_t krp_mmer_key (uint32_t mm, int m, uint32_t mmask, uint32_t mask_ma1)
{
    // reverse complement of an m-mer value (first nucleotide in the most significant position; complement = code ^ 2)
    uint32_t r = 0, x = mm;
    for (int i = 0; i < m; i++) { r = (r << 2) | ((x & 3u) ^ 2u); x >>= 2; }
    uint32_t cm = r < mm ? r : mm;
    uint32_t a1 = ~(cm | (cm >> 2));
    a1 = ((a1 >> 1) & a1) & mask_ma1;                      // an "AA" anywhere but at the prefix: not allowed
    return a1 ? mmask : cm;
}
typedef unsigned __int128 krp_u128;
KRP_HD krp_u128 krp_revcomp (krp_u128 x, int k)
{
    krp_u128 r = 0;
    for (int i = 0; i < k; i++) { r = (r << 2) | (krp_u128)(((uint32_t)x & 3u) ^ 2u); x >>= 2; }
    return r;
}

// Nuc(i) -> code 0..3 of nucleotide i of the read, Bad(i) -> true when it is not A/C/G/T; Sink(minimizer, kxmers) once per super-k-mer
// with a valid minimizer.  Returns the number of such super-k-mers.
template<class Nuc, class Bad, class Sink>
KRP_HD uint32_t krp_scan_read (Nuc nuc, Bad bad, int len, int k, int m, Sink sink)
{
    if (len < k) return 0;                                                               // Sequence2SuperKmer.hpp:144
    const uint32_t mmask = (uint32_t)((1ULL << (2 * m)) - 1), DEF = mmask;                // Model.hpp:1032
    const uint32_t mask_ma1 = (uint32_t)(0x5555555555555555ULL & ((1ULL << ((m - 2) * 2)) - 1));
    const krp_u128 kmask = (((krp_u128)1) << (2 * k)) - 1;                                // k <= 63
    const int nbm = k - m + 1, maxs = k < 32 ? 28 : 60, KX = 4;                           // (8 * sizeof (Type) - 8) / 2, Sequence2SuperKmer.hpp:147
    krp_u128 fwd = 0; int badidx = -1;
    for (int i = 0; i < k; i++) { fwd = (fwd << 2) + nuc (i); if (bad (i)) badidx = i; }
    krp_u128 rev = krp_revcomp (fwd, k);
    bool valid = badidx < 0;
    uint32_t mini = DEF; int pos = -1;
    auto new_minimizer = [&] ()
    {   // Model.hpp:1254-1287: the m-mers of the k-mer from the last to the first, strictly smaller wins
        mini = DEF; pos = -1;
        krp_u128 v = fwd;
        for (int idx = nbm - 1; idx >= 0; idx--)
        {
            const uint32_t c = krp_mmer_key ((uint32_t)v & mmask, m, mmask, mask_ma1);
            if (c < mini) { mini = c; pos = idx; }
            v >>= 2;
        }
    };
    new_minimizer ();
    // "no super-k-mer open" is its own sentinel (SuperKmer::DEFAULT_MINIMIZER, Model.hpp:1372): a k-mer without any allowed m-mer has
    // the minimizer 4^m-1 (Model.hpp:1032), an ordinary value that gets its super-k-mers like any other
    const uint32_t NONE = 1000000000u;
    uint32_t sk_min = NONE; int sk_size = 0; bool prev = false; int kxs = 0; uint32_t cnt = 0, nsk = 0;
    auto flush = [&] ()
    {   // SampleRepart::processSuperkmer (nb_passes = 1): a super-k-mer with a valid minimizer counts once, its kx-mers cnt + 1
        if (sk_min != NONE && sk_size > 0) { sink (sk_min, cnt + 1); nsk++; }
    };
    for (int idx = k; ; idx++)
    {
        if (!valid) { flush (); sk_size = 0; sk_min = NONE; }
        else
        {
            const uint32_t h = mini;
            if (sk_min == NONE) sk_min = h;
            if (h != sk_min || sk_size >= maxs) { flush (); sk_size = 0; }
            sk_min = h;
            const bool w = fwd < rev;                                                     // which(): the forward strand is the canonical one
            if (sk_size == 0) { prev = w; kxs = 0; cnt = 0; }
            else
            {
                if (w != prev || kxs >= KX) { cnt++; kxs = 0; } else kxs++;
                prev = w;
            }
            sk_size++;
        }
        if (idx >= len) break;
        const uint32_t c = nuc (idx);
        if (bad (idx)) badidx = k - 1; else badidx--;                                      // Model.hpp:753-754
        fwd = ((fwd << 2) + c) & kmask;
        rev = ((rev >> 2) + ((krp_u128)(c ^ 2u) << (2 * (k - 1)))) & kmask;
        valid = badidx < 0;
        const uint32_t mmer = krp_mmer_key ((uint32_t)fwd & mmask, m, mmask, mask_ma1);
        pos--;
        if (mmer < mini) { mini = mmer; pos = nbm - 1; }
        else if (pos < 0) new_minimizer ();
    }
    flush ();
    return nsk;
}
