// CPU check of the register minimizer scanner of the partition kernel (gatb_core_b200/csrc/k1_scan.cuh, compiled
// here as plain C++) against a direct restatement: every window minimum recomputed from the nucleotides.
#include "k1_scan.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <string>
#include <algorithm>

struct Ev { uint32_t key; int start, len; bool amb; };
struct Collect
{
    std::vector<Ev>* v;
    void operator() (uint32_t key, int start, int len) { v->push_back (Ev{key, start, len, false}); }
    void operator() (uint32_t key, int start, int len, bool amb) { v->push_back (Ev{key, start, len, amb}); }
};

static uint64_t rng_state = 88172645463325252ULL;
static uint64_t rnd () { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

static long g_kmers = 0, g_amb = 0, g_flagged = 0;
static const char* NT4 = "ACTG";
static std::string rc_string (const std::string& s)
{
    std::string r (s.rbegin (), s.rend ());
    for (auto& c : r) c = c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'C' ? 'G' : 'C';
    return r;
}

template<int WIN, bool HAS_N = false, bool ORI = false> struct Driver
{
    typedef K1Scanner<WIN, 0, HAS_N, ORI> Scanner;
    template<int PH> static void phases (Scanner& sc, Collect& c, int& j0, int nm)
    {
        if (j0 < nm)
        {
            if (j0 + 16 <= nm) sc.template step16<PH, false> (c); else sc.template step16<PH, true> (c);
            if (j0 + 16 >= nm) sc.finish (c);
            j0 += 16;
        }
        if constexpr (PH + 1 < Scanner::PHASES) phases<PH + 1> (sc, c, j0, nm);
    }
    static int run (int m, int n_reads, int mode)
    {
        const int k = m + WIN - 1;
        for (int r = 0; r < n_reads; r++)
        {
            const int len = k + (int)(rnd () % 300);
            const uint64_t roff = rnd () % 97;
            std::vector<uint8_t> nt (roff + len + 64);
            for (size_t i = 0; i < nt.size (); i++)
                nt[i] = mode == 0 ? (uint8_t)(rnd () & 3) : mode == 1 ? (uint8_t)((i / 3) & 1 ? 0 : (rnd () & 3) * ((rnd () & 7) == 0)) : (uint8_t)((i % 5) & 3);
            if (mode == 3)
            {   // hairpins and palindromic m-mers: the minimal rank shows up on both strands
                for (size_t i = 0; i < nt.size (); i++) nt[i] = (uint8_t)(rnd () & 3);
                for (size_t i = roff + 40; i + 40 < nt.size (); i += 61)
                    for (int q = 0; q < 24; q++) nt[i + q] = nt[i - 1 - q] ^ 2;                 // reverse complement of the 24 nt before
            }
            std::vector<uint32_t> words ((nt.size () + 15) / 16 + 8, 0);
            for (size_t i = 0; i < nt.size (); i++) words[i / 16] |= (uint32_t)nt[i] << (2 * (i % 16));
            // invalid nucleotides (HAS_N): sparse, in bursts, or none at all for some reads; bits beyond the read are noise
            std::vector<uint8_t> isbad (nt.size (), 0);
            std::vector<uint32_t> nmask (nt.size () / 32 + 8, 0);
            if (HAS_N)
            {
                const int flavour = r % 4;
                for (size_t i = 0; i < nt.size (); i++)
                    isbad[i] = flavour == 0 ? 0 : flavour == 1 ? (rnd () % 40 == 0) : flavour == 2 ? ((i / 50) % 3 == 1 && rnd () % 3 == 0) : (rnd () % 200 == 0);
                for (size_t i = 0; i < nt.size (); i++) if (isbad[i]) nmask[i / 32] |= 1u << (i % 32);
            }
            // direct keys
            const int nm = len - m + 1, nk = len - k + 1;
            std::vector<uint32_t> key (nm), keyr (nm);
            for (int j = 0; j < nm; j++)
            {
                uint64_t f = 0, rc = 0;
                for (int p = 0; p < m; p++) { uint64_t c = nt[roff + j + p]; f = (f << 2) | c; rc |= (c ^ 2) << (2 * p); }
                key[j] = k1s_key ((uint32_t)(f < rc ? f : rc));
                if (ORI) { key[j] = k1s_key_fwd ((uint32_t)f); keyr[j] = k1s_key_rc ((uint32_t)rc); }
            }
            std::vector<uint32_t> wmin (nk);
            std::vector<uint8_t> isamb (nk, 0);
            std::vector<uint8_t> valid (nk, 1);
            int want_inv = 0;
            for (int i = 0; i < nk; i++)
            {
                uint32_t v = 0xFFFFFFFFu; for (int j = i; j < i + WIN; j++) if (key[j] < v) v = key[j]; wmin[i] = v;
                if (ORI)
                {
                    uint32_t b = 0xFFFFFFFFu; for (int j = i; j < i + WIN; j++) if (keyr[j] < b) b = keyr[j];
                    wmin[i] = v < b ? v : b; isamb[i] = ((v ^ b) == 1u);
                    auto word = [&] (int t) { uint32_t lo = words[(2 * roff) / 32 + t], hi = words[(2 * roff) / 32 + t + 1]; int sh = (int)((2 * roff) & 31); return sh ? ((lo >> sh) | (hi << (32 - sh))) : lo; };
                    const int cls = k1s_classify_kmer (word, i, WIN, m);
                    if (cls != (isamb[i] ? 2 : (v < b ? 0 : 1))) { printf ("classify mismatch\n"); return 1; }
                    if ((cls == 1) != ((wmin[i] & 1u) == 1u) && cls != 2) { printf ("class/parity mismatch\n"); return 1; }
                }
                for (int q = 0; q < k; q++) if (isbad[roff + i + q]) valid[i] = 0;
                want_inv += !valid[i];
            }
            // scanner
            std::vector<Ev> ev; Collect c{&ev};
            Scanner sc;
            sc.begin (words.data (), roff, len, m, 0, nmask.data ());
            if (sc.j >= sc.nm) sc.finish (c);
            int j0 = WIN;
            while (j0 < nm) phases<0> (sc, c, j0, nm);
            // the events tile the VALID k-mers of [0, nk) in order; keys are the window minima; inside a run of valid
            // k-mers a key repeats only after a forced split
            int pos = 0, nbad = 0;
            for (size_t e = 0; e < ev.size (); e++)
            {
                while (pos < nk && !valid[pos]) pos++;                       // invalid k-mers belong to no event
                if (ev[e].start != pos || ev[e].len < 1 || ev[e].len >= 64) { nbad++; break; }
                for (int i = pos; i < pos + ev[e].len && i < nk; i++) if (wmin[i] != ev[e].key || !valid[i]) { nbad++; break; }
                if (ORI)
                {
                    bool any = false;
                    for (int i = pos; i < pos + ev[e].len && i < nk; i++) any |= isamb[i] != 0;
                    if (any && !ev[e].amb) { nbad++; break; }                 // a flag may be conservative, never missing
                    g_kmers += ev[e].len; g_flagged += ev[e].amb ? ev[e].len : 0;
                    for (int i = pos; i < pos + ev[e].len && i < nk; i++) g_amb += isamb[i];
                }
                if (e && ev[e].key == ev[e-1].key && ev[e-1].start + ev[e-1].len == ev[e].start && ev[e-1].len < Scanner::MAXRUN - 1) { nbad++; break; }
                pos += ev[e].len;
            }
            while (pos < nk && !valid[pos]) pos++;
            if (pos != nk) nbad++;
            if (HAS_N && (int)sc.ninv != want_inv) nbad++;
            if (nbad) { printf ("WIN=%d m=%d read %d (len %d roff %d): MISMATCH (%zu events, pos %d, nk %d)\n", WIN, m, r, len, (int)roff, ev.size (), pos, nk); return 1; }
        }
        return 0;
    }
};

int main ()
{
    int bad = 0;
    // ---- oriented scan (k <= 31 windows): same tiling, strand-tagged keys, ambiguity flags; plus the strand symmetry of
    //      the representative: a k-mer and its reverse complement get the same class-resolved sequence ----
    for (int mode = 0; mode < 4; mode++)
        for (int m = 8; m <= 16; m++)
        {
            bad += Driver<8, false, true>::run (m, 60, mode); bad += Driver<16, false, true>::run (m, 60, mode);
            bad += Driver<8, true, true>::run (m, 60, mode);  bad += Driver<16, true, true>::run (m, 60, mode);
        }
    printf ("oriented: %ld k-mers, %ld ambiguous, %ld in flagged super-k-mers\n", g_kmers, g_amb, g_flagged);
    for (int m : { 8, 11, 16 })
        for (int win : { 8, 16 })
        {
            const int k = m + win - 1;
            for (int rep = 0; rep < 400; rep++)
            {
                std::string K (k, 'A');
                for (auto& c : K) c = NT4[rnd () & 3];
                if (rep % 3 == 1) { const int h = m / 2; for (int q = 0; q < h; q++) K[h + q + (m & 1)] = rc_string (K.substr (0, h))[q]; }   // palindromic first m-mer (even m)
                if (rep % 3 == 2) { const std::string r = rc_string (K.substr (0, m)); if (2 * m <= k) K.replace (k - m, m, r); }             // the same m-mer on both strands
                std::string rep_of[2];
                for (int strand = 0; strand < 2; strand++)
                {
                    const std::string S = strand ? rc_string (K) : K;
                    std::vector<uint32_t> w ((k + 15) / 16 + 3, 0);
                    for (int i = 0; i < k; i++) { const char c = S[i]; const uint32_t v = c == 'A' ? 0 : c == 'C' ? 1 : c == 'T' ? 2 : 3; w[i / 16] |= v << (2 * (i % 16)); }
                    auto word = [&] (int t) { return w[t]; };
                    const int cls = k1s_classify_kmer (word, 0, win, m);
                    rep_of[strand] = cls == 0 ? S : cls == 1 ? rc_string (S) : std::min (S, rc_string (S));
                    // k1s_revcomp_span against the string reverse complement
                    uint64_t lo = (uint64_t)w[0] | ((uint64_t)w[1] << 32), hi = (uint64_t)w[2] | ((uint64_t)w[3] << 32);
                    k1s_revcomp_span (lo, hi, k);
                    const std::string R = rc_string (S);
                    for (int i = 0; i < k; i++)
                    {
                        const uint32_t v = (uint32_t)((i < 32 ? lo >> (2 * i) : hi >> (2 * (i - 32))) & 3);
                        if (NT4[v] != R[i]) { printf ("k1s_revcomp_span mismatch (k=%d)\n", k); return 1; }
                    }
                }
                if (rep_of[0] != rep_of[1]) { printf ("orientation rule is not strand-symmetric (m=%d win=%d)\n", m, win); return 1; }
            }
        }
    for (int mode = 0; mode < 3; mode++)
    {
        for (int m = 8; m <= 16; m++)
        {
            bad += Driver<8>::run (m, 60, mode);  bad += Driver<16>::run (m, 60, mode); bad += Driver<24>::run (m, 40, mode);
            bad += Driver<32>::run (m, 40, mode); bad += Driver<40>::run (m, 40, mode); bad += Driver<48>::run (m, 40, mode);
            bad += Driver<8, true>::run (m, 80, mode);  bad += Driver<16, true>::run (m, 80, mode); bad += Driver<24, true>::run (m, 40, mode);
            bad += Driver<32, true>::run (m, 40, mode); bad += Driver<40, true>::run (m, 40, mode); bad += Driver<48, true>::run (m, 40, mode);
        }
    }
    if (bad) { printf ("FAILED\n"); return 1; }
    printf ("k1 scanner ok\n");
    return 0;
}
