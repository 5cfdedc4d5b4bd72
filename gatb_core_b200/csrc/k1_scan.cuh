// k1_scan.cuh -- register-resident minimizer scan of one read (the per-thread part of k1_superkmer_fast).
//
// Replaces, for the device binning order (see k1_partition.cu, K1_MODE_DEVICE), the per-k-mer work of
//   ModelCanonical/ModelMinimizer::next       kmer/impl/Model.hpp:857-884, 1106-1139
//   Sequence2SuperKmer::operator()            kmer/impl/Sequence2SuperKmer.hpp:81-159
// (paths relative to /root/reference/gatb-core/src/gatb/).
//
// Everything is compile-time indexed so that the whole state lives in registers:
//   * the read is seen as a stream of NORMALISED 32-bit words (word t = stream bits [32t, 32t+32) counted from the
//     read's first nucleotide); position j = 16t+u has its m-mer at bit 2u of (word t+1 : word t).  One funnel shift
//     gives the stream-order m-mer x (reverse-complement value = x ^ 0b1010.., common.cuh), one funnel shift on the
//     pair-reversed words gives the forward value; no per-nucleotide rolling state;
//   * rank key = min(fwd, rc) * odd + odd (one IMAD): a pseudo-random minimizer order, m <= 16;
//   * sliding minimum over WIN = k-m+1 keys by blocks of WIN (prefix minima of the current block, suffix minima of
//     the previous one): 3 min per position, WIN registers, no data-dependent rescans;
//   * every normalised word is also parked in a 16-deep shared-memory ring (one column per thread);
//   * a super-k-mer ends when the window minimum changes; the scanner only hands (key, first k-mer, length) to the
//     emitter.  Lengths are bounded by 47 (forced split), the emitter cuts them to the record capacity.
// The code is __host__ __device__ so that tests/cpp/test_k1_scan.cu can check it on the CPU against a direct
// restatement (every window minimum recomputed from scratch).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define K1S_HD __host__ __device__ __forceinline__
#else
#define K1S_HD inline
#endif

K1S_HD uint32_t k1s_fshr (uint32_t lo, uint32_t hi, int s)      // low 32 bits of (hi:lo) >> (s & 31)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r (lo, hi, s);
#else
    s &= 31; return s ? ((lo >> s) | (hi << (32 - s))) : lo;
#endif
}
K1S_HD uint32_t k1s_fshl (uint32_t lo, uint32_t hi, int s)      // high 32 bits of (hi:lo) << (s & 31)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_l (lo, hi, s);
#else
    s &= 31; return s ? ((hi << s) | (lo >> (32 - s))) : hi;
#endif
}
K1S_HD uint32_t k1s_brev (uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __brev (x);
#else
    uint32_t r = 0; for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i); return r;
#endif
}
K1S_HD uint32_t k1s_ldg (const uint32_t* p)
{
#if defined(__CUDA_ARCH__)
    return __ldg (p);
#else
    return *p;
#endif
}
#define K1S_FSHR(lo, hi, s) k1s_fshr ((lo), (hi), (s))
#define K1S_FSHL(lo, hi, s) k1s_fshl ((lo), (hi), (s))
#define K1S_BREV(x)         k1s_brev (x)
#define K1S_LDG(p)          k1s_ldg (p)

#define K1S_MUL 0x9E3779B1u
#define K1S_ADD 0x7F4A7C15u

// rank key of a canonical m-mer (device binning order)
K1S_HD uint32_t k1s_key (uint32_t cm) { return cm * K1S_MUL + K1S_ADD; }

// ---- ORIENTED scan (ORI = true): strand-tagged rank keys -----------------------------------------------------------
// The forward m-mer f of a position gets an EVEN key, its reverse complement rc an ODD one, with the same 31-bit rank
// function in the upper bits:  hF = 2*(f*MUL + C) , hR = 2*(rc*MUL + C) + 1  (mod 2^32).
// For a k-mer K let A = min hF, B = min hR over its window; for revcomp(K) the two swap roles (B-1, A+1).  Rule:
//   t = min(A, B);  t even and B != A+1 : K is taken as read            (class F)
//                   t odd               : revcomp(K) is taken            (class R; then B>>1 < A>>1 strictly)
//                   B == A+1            : the minimal rank occurs on both strands (or in a palindrome): AMBIGUOUS,
//                                         the k-mer is stored as min(K, revcomp K) like GATB's canonical form.
// The representative is a function of the k-mer alone and the same for K and revcomp(K) whatever the rank function
// (tools/prototypes/oriented_keys.cpp checks the rule; tests/cpp/test_k1_scan.cpp checks this implementation), and
// t >> 1 (the rank of the minimizer) is the same for both strands: it selects the bin.  All k-mers of a super-k-mer
// (constant t) therefore share one orientation relative to the read: the whole record is stored in that orientation
// and the counting kernel uses plain slices of the record as table keys -- no reverse complement, no min() per k-mer
// occurrence.
#define K1S_MUL2 ((K1S_MUL << 1) & 0xFFFFFFFFu)
#define K1S_ADDF (K1S_ADD & ~1u)
#define K1S_ADDR (K1S_ADD | 1u)
K1S_HD uint32_t k1s_key_fwd (uint32_t f)  { return f  * K1S_MUL2 + K1S_ADDF; }
K1S_HD uint32_t k1s_key_rc  (uint32_t rc) { return rc * K1S_MUL2 + K1S_ADDR; }

K1S_HD uint32_t k1s_pair_reverse (uint32_t x)
{
    uint32_t r = K1S_BREV (x);
    return ((r >> 1) & 0x55555555u) | ((r << 1) & 0xAAAAAAAAu);
}

constexpr int k1s_gcd (int a, int b) { return b ? k1s_gcd (b, a % b) : a; }

// RS = stride (in words) of the optional shared-memory word ring: every normalised word is also stored at
// ring[(word index % K1S_RING) * RS], so that whoever builds the record of a closed super-k-mer reads the nucleotides
// from shared memory instead of going back to global memory (0 = no ring, used by the host test).
#define K1S_RING 16
// HAS_N: the read may hold invalid nucleotides (1 bit per nucleotide in 'nmask', include/gatb_gpu.h): a k-mer that
// overlaps one is invalid, it is dropped and it ends the open super-k-mer (Sequence2SuperKmer.hpp:95-108).  The scanner
// keeps the number of nucleotides since the last invalid one; the keys are computed as usual (an invalid nucleotide is
// encoded like G), only the k-mers are masked.
// ORI: oriented scan (see above); emit gets a fourth argument: true when some k-mer of the super-k-mer is ambiguous.
// SM: the packed words of the read were staged in shared memory (k1_partition.cu stages a warp's reads with one TMA bulk
// copy per 32 reads); the scanner then takes its words with ld.shared from the 32-bit shared address 'wps'.
template<int WIN, int RS = 0, bool HAS_N = false, bool ORI = false, bool SM = false>
struct K1Scanner
{
    static constexpr int LCM    = WIN / k1s_gcd (WIN, 16) * 16;   // positions after which (window slot, word phase) repeat
    static constexpr int PHASES = LCM / 16;
    static constexpr int MAXRUN = 32;                              // forced split threshold (lengths stay < 32+16)

    uint32_t suf[WIN];            // keys of the current block / suffix minima of the previous one (ORI: forward-strand keys)
    uint32_t sur[ORI ? WIN : 1];  // ORI: the same for the reverse-complement keys
    uint32_t pr, amb;             // ORI: prefix minimum of the reverse keys; min over the open super-k-mer of A ^ B (1 = ambiguous)
    uint32_t na, nb, ra, rb;      // normalised words t, t+1 and their pair-reversed images
    uint32_t raw;                 // last raw word consumed
    uint32_t ahead;               // raw word loaded one step early (its latency hides behind 16 positions of work)
    const uint32_t* wp;           // next raw word
    uint32_t wps;                 // SM: shared-memory address of the next raw word
    uint32_t* ring;               // this thread's column of the word ring (RS != 0)
    uint32_t nw;                  // normalised words produced so far
    uint32_t sh;                  // bit offset of the read inside its first raw word
    uint32_t mmask, aam, fsh;     // 4^m-1, 0xAAAAAAAA & mmask, 32-2m
    uint32_t p, cur;              // prefix minimum of the current block, key of the open super-k-mer
    int start;                    // first k-mer of the open super-k-mer
    int j;                        // next m-mer position
    int nm;                       // m-mer positions of the read (len-m+1)
    // HAS_N only
    const uint32_t* nmask; uint64_t roff_nt;
    uint32_t bb;                  // invalid bits of the nucleotides [16t, 16t+32) of the read, t = current word
    int bg;                       // next 16-nucleotide group to fetch
    int since_bad, kk, m1;        // nucleotides since the last invalid one; k; m-1
    bool open;                    // a super-k-mer is open
    uint32_t ninv;                // invalid k-mers met

    K1S_HD uint32_t bad_group (int g) const
    {
        const uint64_t a = roff_nt + 16 * (uint64_t)g;
        const uint32_t w0 = K1S_LDG (nmask + (a >> 5)), w1 = K1S_LDG (nmask + (a >> 5) + 1);
        return K1S_FSHR (w0, w1, (int)(a & 31)) & 0xFFFFu;
    }
    // the nucleotide entering the window at word position u is nucleotide 16t+u+m-1 of the read
    K1S_HD void note_nucleotide (int u) { since_bad = ((bb >> (u + m1)) & 1u) ? 0 : since_bad + 1; }

    K1S_HD uint32_t load_raw ()
    {
#if defined(__CUDA_ARCH__)
        if constexpr (SM) { uint32_t v; asm volatile ("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(wps)); wps += 4; return v; }
#endif
        const uint32_t v = K1S_LDG (wp); wp++; return v;
    }
    K1S_HD uint32_t next_word ()
    {
        const uint32_t r = ahead;
        ahead = load_raw ();
        const uint32_t n = K1S_FSHR (raw, r, sh);
        raw = r;
        if (RS) { ring[(nw & (K1S_RING - 1)) * RS] = n; nw++; }
        return n;
    }
    K1S_HD void advance_word ()
    {
        na = nb; ra = rb; nb = next_word (); rb = k1s_pair_reverse (nb);
        if (HAS_N) { bb = (bb >> 16) | (bad_group (bg) << 16); bg++; }
    }

    template<int U> K1S_HD uint32_t key_at () const
    {
        const uint32_t x  = U ? K1S_FSHR (na, nb, 2 * U) : na;
        const uint32_t rc = (x & mmask) ^ aam;
        const uint32_t g  = U ? K1S_FSHL (rb, ra, 2 * U) : ra;
        const uint32_t f  = g >> fsh;
        return k1s_key (f < rc ? f : rc);
    }

    template<int U> K1S_HD void keys_at (uint32_t& hf, uint32_t& hr) const
    {
        const uint32_t x  = U ? K1S_FSHR (na, nb, 2 * U) : na;
        const uint32_t rc = (x & mmask) ^ aam;
        const uint32_t g  = U ? K1S_FSHL (rb, ra, 2 * U) : ra;
        hf = k1s_key_fwd (g >> fsh); hr = k1s_key_rc (rc);
    }
    K1S_HD static uint32_t min2 (uint32_t a, uint32_t b) { return a < b ? a : b; }
    template<class Emit> K1S_HD void do_emit (Emit& emit, uint32_t key, int s, int l)
    {
        if constexpr (ORI) emit (key, s, l, amb == 1u); else emit (key, s, l);
    }

    // one position: T = slot in the window block, U = nucleotide inside the normalised word; i = k-mer index j-(WIN-1)
    template<int T, int U, bool TAIL, class Emit> K1S_HD void position (int q, Emit& emit)
    {
        if (TAIL && j + q >= nm) return;
        uint32_t wmin, x = 0;
        if constexpr (ORI)
        {
            uint32_t hf, hr; keys_at<U> (hf, hr);
            const uint32_t sf = (T + 1 < WIN) ? suf[(T + 1) % WIN] : 0xFFFFFFFFu;
            const uint32_t sr = (T + 1 < WIN) ? sur[(T + 1) % WIN] : 0xFFFFFFFFu;
            suf[T] = hf; sur[T] = hr;
            p  = (T == 0) ? hf : min2 (hf, p);
            pr = (T == 0) ? hr : min2 (hr, pr);
            const uint32_t wf = (T + 1 < WIN) ? min2 (sf, p) : p, wr = (T + 1 < WIN) ? min2 (sr, pr) : pr;
            wmin = min2 (wf, wr); x = wf ^ wr;
        }
        else
        {
            const uint32_t key = key_at<U> ();
            const uint32_t s = (T + 1 < WIN) ? suf[(T + 1) % WIN] : 0xFFFFFFFFu;
            suf[T] = key;
            p = (T == 0) ? key : (key < p ? key : p);
            wmin = (T + 1 < WIN) ? (s < p ? s : p) : p;
        }
        if (HAS_N)
        {
            note_nucleotide (U);
            const int i = j + q - (WIN - 1);
            if (since_bad >= kk)
            {
                if (!open || wmin != cur)
                {
                    if (open) do_emit (emit, cur, start, i - start);
                    cur = wmin; start = i; open = true; amb = x;
                }
                else if (ORI) amb = min2 (amb, x);
            }
            else
            {
                ninv++;
                if (open) { do_emit (emit, cur, start, i - start); open = false; }
            }
        }
        else if (wmin != cur)
        {
            const int i = j + q - (WIN - 1);
            do_emit (emit, cur, start, i - start);
            cur = wmin; start = i; amb = x;
        }
        else if (ORI) amb = min2 (amb, x);
        if (T == WIN - 1)
        {
            #pragma unroll
            for (int u = WIN - 2; u >= 1; u--) suf[u] = suf[u] < suf[u + 1] ? suf[u] : suf[u + 1];
            if constexpr (ORI)
            {
                #pragma unroll
                for (int u = WIN - 2; u >= 1; u--) sur[u] = sur[u] < sur[u + 1] ? sur[u] : sur[u + 1];
            }
        }
        if (U == 15) advance_word ();
    }

    // ---- the read starts at nucleotide 'roff' of the packed stream 'words32'; len >= k = m+WIN-1 is required -------
    // Processes the first block (m-mer positions 0..WIN-1): afterwards the super-k-mer of k-mer 0 is open.
    uintptr_t ring_shared_word0;  // SM: shared address of the raw word that holds the read's first nucleotide
    K1S_HD void begin_shared (uint32_t word0_saddr, uint64_t roff, int len, int m, uint32_t* ring_column, const uint32_t* n_mask)
    { ring_shared_word0 = word0_saddr; begin ((const uint32_t*)0, roff, len, m, ring_column, n_mask); }
    K1S_HD void begin (const uint32_t* words32, uint64_t roff, int len, int m, uint32_t* ring_column = 0, const uint32_t* n_mask = 0)
    {
        ring = ring_column; nw = 0;
        if (HAS_N)
        {
            nmask = n_mask; roff_nt = roff; kk = m + WIN - 1; m1 = m - 1; ninv = 0; open = false;
            bb = bad_group (0) | (bad_group (1) << 16); bg = 2;
            const uint32_t lowbits = bb & ((1u << m1) - 1);             // the m-1 nucleotides before the first entering one
            int msb = -1; for (int b = 0; b < 16; b++) if (lowbits & (1u << b)) msb = b;
            since_bad = m1 - 1 - msb;
        }
        mmask = (m >= 16) ? 0xFFFFFFFFu : ((1u << (2 * m)) - 1);
        aam = 0xAAAAAAAAu & mmask; fsh = 32 - 2 * m;
        nm = len - m + 1;
        const uint64_t b0 = 2 * roff;
        wp = words32 + (b0 >> 5); sh = (uint32_t)(b0 & 31);
        if (SM) wps = (uint32_t)(uintptr_t)ring_shared_word0;       // set by begin_shared
        raw = load_raw ();
        ahead = load_raw ();
        na = next_word (); ra = k1s_pair_reverse (na);
        nb = next_word (); rb = k1s_pair_reverse (nb);
        first_block (Int<0> ());
        cur = p; start = 0; j = WIN; amb = 0xFFFFFFFFu; pr = ORI ? pr : 0u;
        if constexpr (ORI) { cur = min2 (p, pr); amb = p ^ pr; }
        if (HAS_N) { open = since_bad >= kk; if (!open) ninv++; }
        #pragma unroll
        for (int u = WIN - 2; u >= 1; u--) suf[u] = suf[u] < suf[u + 1] ? suf[u] : suf[u + 1];
        if constexpr (ORI)
        {
            #pragma unroll
            for (int u = WIN - 2; u >= 1; u--) sur[u] = sur[u] < sur[u + 1] ? sur[u] : sur[u + 1];
        }
    }
    template<int N> struct Int {};
    template<int T> K1S_HD void first_block (Int<T>)
    {
        if constexpr (ORI)
        {
            uint32_t hf, hr; keys_at<T % 16> (hf, hr);
            suf[T] = hf; sur[T] = hr;
            p = (T == 0) ? hf : min2 (hf, p); pr = (T == 0) ? hr : min2 (hr, pr);
        }
        else
        {
            const uint32_t key = key_at<T % 16> ();
            suf[T] = key;
            p = (T == 0) ? key : (key < p ? key : p);
        }
        if (HAS_N) note_nucleotide (T % 16);
        if (T % 16 == 15) advance_word ();
        first_block (Int<T + 1> ());
    }
    K1S_HD void first_block (Int<WIN>) {}

    // ---- 16 positions starting at j (j = WIN + 16*n, PH = n % PHASES) -----------------------------------------------
    template<int PH, bool TAIL, class Emit> K1S_HD void step16 (Emit& emit)
    {
        // forced split of very long runs (tandem repeats): everything but the last k-mer seen leaves, so that a change
        // of the minimum at the next position still closes a non-empty super-k-mer
        if ((!HAS_N || open) && j - (WIN - 1) - start >= MAXRUN) { const int i = j - (WIN - 1) - 1; do_emit (emit, cur, start, i - start); start = i; }
        run16<PH, 0, TAIL> (emit, Int<0> ());
        j += 16;
    }
    template<int PH, int Q, bool TAIL, class Emit, int QQ> K1S_HD void run16 (Emit& emit, Int<QQ>)
    {
        position<(WIN + 16 * PH + QQ) % WIN, (WIN + QQ) % 16, TAIL> (QQ, emit);
        run16<PH, Q, TAIL> (emit, Int<QQ + 1> ());
    }
    template<int PH, int Q, bool TAIL, class Emit> K1S_HD void run16 (Emit&, Int<16>) {}

    // the read is over: close the open super-k-mer
    template<class Emit> K1S_HD void finish (Emit& emit)
    {
        const int nk = nm - (WIN - 1);
        if (!HAS_N || open) do_emit (emit, cur, start, nk - start);
    }
};

// ---- exact class of ONE k-mer under the orientation rule (slow path of the emitter: super-k-mers flagged ambiguous) ---
// word(t) returns the normalised 32-bit word t of the read (stream bits [32t, 32t+32)); i = index of the k-mer.
// Returns 0 = class F (stored as read), 1 = class R (stored reverse-complemented), 2 = ambiguous.
template<class WordAt> K1S_HD int k1s_classify_kmer (WordAt word, int i, int win, int m)
{
    const uint32_t mmask = (m >= 16) ? 0xFFFFFFFFu : ((1u << (2 * m)) - 1);
    const uint32_t aam = 0xAAAAAAAAu & mmask;
    uint32_t A = 0xFFFFFFFFu, B = 0xFFFFFFFFu;
    for (int j = i; j < i + win; j++)
    {
        const uint32_t x  = K1S_FSHR (word (j >> 4), word ((j >> 4) + 1), 2 * (j & 15)) & mmask;
        const uint32_t hf = k1s_key_fwd (k1s_pair_reverse (x) >> (32 - 2 * m)), hr = k1s_key_rc (x ^ aam);
        A = hf < A ? hf : A; B = hr < B ? hr : B;
    }
    return ((A ^ B) == 1u) ? 2 : (A < B ? 0 : 1);
}

// ---- stream bits of a span of nn <= 64 nucleotides, reverse-complemented (still in stream order) ---------------------
// s = the span's 2nn stream bits (lo, hi), masked.  revcomp in stream order = pair-reversal of the whole span with
// every nucleotide complemented (XOR 0b10).
K1S_HD void k1s_revcomp_span (uint64_t& lo, uint64_t& hi, int nn)
{
    const uint32_t w0 = (uint32_t)lo, w1 = (uint32_t)(lo >> 32), w2 = (uint32_t)hi, w3 = (uint32_t)(hi >> 32);
    const uint64_t ylo = ((uint64_t)k1s_pair_reverse (w2) << 32) | k1s_pair_reverse (w3);      // pair-reversed 128-bit value
    const uint64_t yhi = ((uint64_t)k1s_pair_reverse (w0) << 32) | k1s_pair_reverse (w1);
    const int s = 128 - 2 * nn;                                                                  // 0..126, even
    uint64_t rl, rh;
    if (s == 0)      { rl = ylo; rh = yhi; }
    else if (s < 64) { rl = (ylo >> s) | (yhi << (64 - s)); rh = yhi >> s; }
    else if (s == 64){ rl = yhi; rh = 0; }
    else             { rl = yhi >> (s - 64); rh = 0; }
    const uint64_t ml = nn >= 32 ? ~0ULL : ((1ULL << (2 * nn)) - 1), mh = nn <= 32 ? 0ULL : (nn >= 64 ? ~0ULL : ((1ULL << (2 * nn - 64)) - 1));
    lo = rl ^ (0xAAAAAAAAAAAAAAAAULL & ml); hi = rh ^ (0xAAAAAAAAAAAAAAAAULL & mh);
}
