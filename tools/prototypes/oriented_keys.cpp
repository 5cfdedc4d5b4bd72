// Prototype (CPU only, not part of the product): "canonical by minimizer strand" keys for the counting tables.
//
// Today k2b rebuilds min(forward, reverse complement) for every k-mer occurrence (about 40 % of its window work,
// DESIGN.md section 8).  Idea for the next round: orient a whole super-k-mer by the strand of its minimizer, so that the
// table key of a k-mer is just a slice of the (already oriented) record, and convert to GATB's canonical value only when
// a distinct k-mer is emitted.  The orientation must be a function of the k-mer alone and identical for a k-mer and its
// reverse complement.  Rule checked here:
//   * every window position p carries (P_p, s_p): P_p = upper 31 bits of the rank key of the canonical m-mer,
//     s_p = 0 when the forward m-mer is the canonical one, 1 when the reverse complement is; palindromes count as BOTH;
//   * a = min_p (P_p << 1 | s_p), b = min_p (P_p << 1 | !s_p) over the k-mer's window (two sliding minima instead of one);
//   * a == b  <=>  the minimal P occurs with both strands (or as a palindrome): the k-mer is AMBIGUOUS and keeps the
//     classic key min(forward, revcomp) (tagged so that the two key kinds never collide);
//   * otherwise the k-mer is oriented by the low bit of a: representative R(K) = K if it is 0, revcomp(K) if it is 1.
// The program verifies on random, repetitive and hairpin-rich sequences that R(K) == R(revcomp K) and
// ambiguous(K) == ambiguous(revcomp K) for every k-mer, and reports how rare the ambiguous ones are.
//   g++ -O2 -std=c++17 tools/prototypes/oriented_keys.cpp -o /tmp/oriented_keys && /tmp/oriented_keys
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <algorithm>

static uint64_t st = 0x2545F4914F6CDD1DULL;
static uint64_t rnd () { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; }
static uint32_t rank_key (uint32_t cm) { return cm * 0x9E3779B1u + 0x7F4A7C15u; }       // k1s_key

struct KInfo { bool ambiguous; std::string repr; };

static std::string revcomp (const std::string& s)
{
    std::string r (s.rbegin (), s.rend ());
    for (auto& c : r) c = c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'C' ? 'G' : 'C';
    return r;
}
static int code (char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'T' ? 2 : 3; }       // GATB: A=0 C=1 T=2 G=3

static KInfo classify (const std::string& K, int m)
{
    const int w = (int)K.size () - m + 1;
    uint64_t a = ~0ULL, b = ~0ULL;
    for (int p = 0; p < w; p++)
    {
        uint64_t f = 0, rc = 0;
        for (int q = 0; q < m; q++) { uint64_t c = code (K[p + q]); f = (f << 2) | c; rc |= (c ^ 2) << (2 * q); }
        const uint32_t cm = (uint32_t)std::min (f, rc);
        const uint64_t P = rank_key (cm) >> 1;
        const int s = f < rc ? 0 : 1;
        const bool pal = f == rc;
        a = std::min (a, (P << 1) | (uint64_t)(pal ? 0 : s));
        b = std::min (b, (P << 1) | (uint64_t)(pal ? 0 : !s));
    }
    KInfo r;
    r.ambiguous = (a == b);
    if (r.ambiguous) { const std::string rcK = revcomp (K); r.repr = std::min (K, rcK); }        // any symmetric choice
    else r.repr = (a & 1) ? revcomp (K) : K;
    return r;
}

int main ()
{
    const char* nt = "ACTG";
    long total = 0, ambiguous = 0, bad = 0;
    for (int flavour = 0; flavour < 4; flavour++)
        for (int m : { 8, 12, 16 })
            for (int k : { m + 7, m + 15 })
            {
                std::string g (flavour == 3 ? 20000 : 60000, 'A');
                for (auto& c : g) c = nt[rnd () & 3];
                if (flavour == 1) for (size_t i = 100; i + 100 < g.size (); i += 97) g.replace (i, 40, g.substr (i - 60, 40));                 // tandem copies
                if (flavour == 2) for (size_t i = 100; i + 100 < g.size (); i += 71) g.replace (i, 30, revcomp (g.substr (i - 45, 30)));      // hairpins
                if (flavour == 3) for (size_t i = 0; i + 2 * m < g.size (); i += 53) { std::string h = g.substr (i, m / 2); g.replace (i + m / 2, m / 2, revcomp (h)); }   // palindromic m-mers
                for (size_t i = 0; i + k <= g.size (); i++)
                {
                    const std::string K = g.substr (i, k), R = revcomp (K);
                    const KInfo x = classify (K, m), y = classify (R, m);
                    total++; ambiguous += x.ambiguous;
                    if (x.ambiguous != y.ambiguous || x.repr != y.repr) bad++;
                }
            }
    printf ("%ld k-mers, %ld ambiguous (%.4f %%), %ld inconsistent\n", total, ambiguous, 100.0 * ambiguous / total, bad);
    return bad != 0;
}
