"""The C restatement (oracle/kmer_oracle.c) against the golden vectors of the reference's own unit tests."""
import numpy as np

from golden import reference_vectors as G

NT = {"A": 0, "C": 1, "T": 2, "G": 3}


def enc(s):
    v = 0
    for c in s:
        v = v * 4 + NT[c]
    return v


def rc(s):
    return "".join({"A": "T", "C": "G", "G": "C", "T": "A"}[c] for c in reversed(s))


def identity_repart(m, nparts=1):
    return np.zeros(4 ** m, np.uint16)


def test_direct_and_canonical_k3(oracle):
    # TestKmer.cpp:141-190
    assert [enc(G.KMER3_SEQ[i:i + 3]) for i in range(10)] == G.KMER3_DIRECT      # the encoding itself
    lo, hi, mi, va, st = oracle.kmers(G.KMER3_SEQ, 3, 2)
    assert lo.tolist() == G.KMER3_CANONICAL and not hi.any() and va.all()


def test_canonical_k5(oracle):
    # TestKmer.cpp:233-261
    lo, _, _, va, _ = oracle.kmers(G.KMER5_SEQ, 5, 3)
    assert lo.tolist() == G.KMER5_CANONICAL and va.all()


def test_minimizer3_table(oracle):
    # TestKmer.cpp:434-502
    k, m = G.MINIMIZER3_K, G.MINIMIZER3_M
    lo, _, mi, va, st = oracle.kmers(G.MINIMIZER3_SEQ, k, m)
    assert len(lo) == len(G.MINIMIZER3_TABLE)
    for i, (kmer, minim, pos, changed) in enumerate(G.MINIMIZER3_TABLE):
        assert lo[i] == enc(kmer)
        assert mi[i] == enc(minim)
        assert mi[i] == G.MINIMIZER3_VALUES[i]
        # the canonical k-mer is whichever strand is smaller
        fwd = G.MINIMIZER3_SEQ[i:i + k]
        assert enc(kmer) == min(enc(fwd), enc(rc(fwd)))
        assert bool(st[i]) == (enc(fwd) < enc(rc(fwd)))


def test_minimizer_bruteforce_with_AA_rule(oracle):
    # the brute-force oracle of TestKmer.cpp:265-374: minimum canonical m-mer among those without an inner "AA"
    rng = np.random.default_rng(7)
    for k, m in ((15, 7), (21, 8), (31, 10), (27, 5), (63, 10), (41, 8)):
        seq = "".join("ACGT"[i] for i in rng.integers(0, 4, 300))
        lo, hi, mi, va, st = oracle.kmers(seq, k, m)
        for i in range(len(seq) - k + 1):
            best = 4 ** m - 1
            for j in range(k - m + 1):
                mm = seq[i + j:i + j + m]
                c = min(mm, rc(mm), key=enc)
                if "AA" in c[1:]:
                    continue
                best = min(best, enc(c))
            assert mi[i] == best, (k, m, i)
            v = min(enc(seq[i:i + k]), enc(rc(seq[i:i + k])))
            assert (int(hi[i]) << 64 | int(lo[i])) == v


def test_badchar(oracle):
    # TestKmer.cpp:542-569: k-mers overlapping an N are invalid, N is encoded as G
    k = G.BADCHAR_K
    lo, _, _, va, _ = oracle.kmers(G.BADCHAR_SEQ, k, 5)
    assert va.astype(bool).tolist() == G.BADCHAR_VALID
    for i in range(len(lo)):
        f = G.BADCHAR_SEQ[i:i + k].replace("N", "G")
        assert lo[i] == min(enc(f), enc(rc(f)))


def test_dsk_check1_solid_counts(oracle):
    # TestDSK.cpp:147-241
    for seqs, k, nks, expected in G.DSK1_CASES:
        m = min(8, k - 1)
        res = oracle.dsk([s.encode() for s in seqs], k, m, identity_repart(m), 1, abundance_min=nks)
        assert res["stats"][3] == expected, (k, nks)
        assert len(res["solid"][0][0]) == expected


def test_dsk_check2_exact_set_and_checksum(oracle):
    # TestDSK.cpp:244-341
    res = oracle.dsk([G.DSK2_SEQ.encode()], 31, 10, identity_repart(10), 1, abundance_min=1)
    lo, hi, cn = res["solid"][0]
    assert sorted(lo.tolist()) == sorted(G.DSK2_SOLID) and not hi.any()
    assert sum(lo.tolist()) % 2 ** 64 == G.DSK2_CHECKSUM
    assert lo.tolist() == sorted(lo.tolist())               # ascending emission order


def test_dsk_all_kmers_invariant(oracle):
    # TestDSK.cpp:615-678 (DSK_perBankKmer): a bank holding all 4^k k-mers -> 4^k/2 canonical k-mers (k odd), abundance 2
    k, m = 5, 3
    seqs = []
    for v in range(4 ** k):
        seqs.append("".join("ACTG"[(v >> (2 * (k - 1 - i))) & 3] for i in range(k)).encode())
    res = oracle.dsk(seqs, k, m, identity_repart(m), 1, abundance_min=1)
    lo, _, cn = res["parts"][0]
    assert len(lo) == 4 ** k // 2 and (cn == 2).all()


def test_superkmer_roundtrip_and_split_rules(oracle):
    rng = np.random.default_rng(3)
    k, m = 31, 10
    seqs = ["".join("ACGT"[i] for i in rng.integers(0, 4, 150)).encode() for _ in range(50)]
    seqs.append(b"A" * 200)                                     # one constant minimizer: only the maxs=28 rule splits
    seqs.append(b"ACGT" * 5)                                    # shorter than k: skipped
    nparts = 4
    repart = rng.integers(0, nparts, 4 ** m).astype(np.uint16)
    streams, stats = oracle.superkmers(seqs, k, m, repart, nparts)
    total = 0
    multiset = []
    for p in range(nparts):
        lo, hi = oracle.decode_superkmers(streams[p], k)
        total += len(lo)
        multiset += lo.tolist()
        # every k-mer of partition p has a minimizer mapped to p
    expect = []
    for s in seqs:
        lo, hi, mi, va, st = oracle.kmers(s, k, m)
        expect += lo[va.astype(bool)].tolist()
    assert total == stats[1] == stats[2] == len(expect)
    assert sorted(multiset) == sorted(expect)
    # poly-A read: 170 k-mers, one minimizer => ceil(170/28) = 7 super-k-mers of <= 28 k-mers
    s2, st2 = oracle.superkmers([b"A" * 200], k, m, repart, nparts)
    assert st2[0] == 7 and st2[1] == 170


def test_bloom_sizing(oracle):
    # SURVEY.md 8(a) F + BASELINE.md: 49 972 solid 21-mers -> 293 528 bits, 4 hashes
    assert oracle.bloom_params(21, 49972) == (293528, 4)
    assert abs(oracle.L.orc_nbits_per_kmer(31) - 6.03437) < 1e-6


def test_bloom_no_false_negative(oracle):
    # TestContainer.cpp:63-130 property
    rng = np.random.default_rng(5)
    for words, k in ((1, 31), (2, 63)):
        lo = rng.integers(0, 2 ** 62, 2000, dtype=np.uint64)
        hi = rng.integers(0, 2 ** 62, 2000, dtype=np.uint64) if words == 2 else None
        for kind in ("basic", "cache", "neighbor"):
            bits, bitsize = oracle.bloom(kind, 20000, 4, k, words, lo, hi)
            again, _ = oracle.bloom(kind, 20000, 4, k, words, np.concatenate([lo, lo]), None if hi is None else np.concatenate([hi, hi]))
            assert (bits == again).all()                         # idempotent
            assert 0 < np.unpackbits(bits).sum() <= 4 * 2000
