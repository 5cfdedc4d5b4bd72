// CPU check of the register minimizer scanner of the partition kernel (gatb_core_b200/csrc/k1_scan.cuh, compiled
// here as plain C++) against a direct restatement: every window minimum recomputed from the nucleotides.
#include "k1_scan.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <vector>

struct Ev { uint32_t key; int start, len; };
struct Collect { std::vector<Ev>* v; void operator() (uint32_t key, int start, int len) { v->push_back (Ev{key, start, len}); } };

static uint64_t rng_state = 88172645463325252ULL;
static uint64_t rnd () { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

template<int WIN, bool HAS_N = false> struct Driver
{
    typedef K1Scanner<WIN, 0, HAS_N> Scanner;
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
            std::vector<uint32_t> key (nm);
            for (int j = 0; j < nm; j++)
            {
                uint64_t f = 0, rc = 0;
                for (int p = 0; p < m; p++) { uint64_t c = nt[roff + j + p]; f = (f << 2) | c; rc |= (c ^ 2) << (2 * p); }
                key[j] = k1s_key ((uint32_t)(f < rc ? f : rc));
            }
            std::vector<uint32_t> wmin (nk);
            std::vector<uint8_t> valid (nk, 1);
            int want_inv = 0;
            for (int i = 0; i < nk; i++)
            {
                uint32_t v = 0xFFFFFFFFu; for (int j = i; j < i + WIN; j++) if (key[j] < v) v = key[j]; wmin[i] = v;
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
