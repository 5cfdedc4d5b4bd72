"""The C restatement against the UNMODIFIED reference compiled from /root/reference (oracle/_ref/libgatbref.so).
Skipped when that library is absent (it is built by oracle/Makefile only where /root/reference exists, and travels
to the GPU box as a prebuilt file)."""
import os
import tempfile

import numpy as np
import pytest

import oracle_lib


def rand_seq(rng, n, p_n=0.0):
    s = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)].copy()
    if p_n:
        s[rng.random(n) < p_n] = ord("N")
    return s.tobytes()


def write_fasta(path, seqs):
    with open(path, "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b">r%d\n" % i + s + b"\n")


@pytest.mark.parametrize("k,m", [(21, 8), (31, 10), (15, 7), (63, 10), (47, 9), (31, 5)])
def test_kmers_and_minimizers(oracle, reference, k, m):
    rng = np.random.default_rng(k * 100 + m)
    for n, p_n in ((400, 0.0), (400, 0.02), (k, 0.0), (k + 1, 0.3), (k - 1, 0.0)):
        s = rand_seq(rng, n, p_n)
        a, b = oracle.kmers(s, k, m), reference.kmers(s, k, m)
        for x, y, name in zip(a, b, ("lo", "hi", "minimizer", "valid", "strand")):
            if name == "minimizer":        # only defined by the algorithm for valid k-mers... but state is shared: compare all
                pass
            assert (x == y).all(), (name, n, p_n)


def test_integer_helpers(oracle, reference):
    rng = np.random.default_rng(1)
    for _ in range(300):
        lo, hi = int(rng.integers(0, 2 ** 63)) * 2 + int(rng.integers(0, 2)), int(rng.integers(0, 2 ** 62))
        for k in (9, 21, 31):
            v = lo & (4 ** k - 1)
            assert oracle.revcomp(v, 0, k, 1) == reference.revcomp(v, 0, k, 1)
        for k in (33, 47, 63, 31, 32):
            full = ((hi << 64) | lo) & (4 ** k - 1)
            assert oracle.revcomp(full & (2 ** 64 - 1), full >> 64, k, 2) == reference.revcomp(full & (2 ** 64 - 1), full >> 64, k, 2)
        seed = int(rng.integers(0, 2 ** 63))
        for words in (1, 2):
            assert oracle.hash1(lo, hi, words, seed) == reference.hash1(lo, hi, words, seed)
            for shift in range(0, 5):
                assert oracle.simplehash16(lo, hi, words, shift) == reference.simplehash16(lo, hi, words, shift)
    for k in range(129):
        ref = reference.nbits_per_kmer(k)
        assert oracle.L.orc_nbits_per_kmer(k) == (ref if ref != 0 else 1.0)


@pytest.mark.parametrize("k,m,nparts,npass", [(31, 10, 7, 1), (21, 8, 3, 2), (63, 10, 5, 1), (12, 6, 2, 1), (40, 8, 4, 3)])
def test_superkmer_streams_byte_identical(oracle, reference, k, m, nparts, npass):
    rng = np.random.default_rng(k + nparts)
    seqs = [rand_seq(rng, int(rng.integers(20, 400)), 0.01 if i % 3 == 0 else 0.0) for i in range(200)]
    seqs += [b"A" * 300, b"ACGTN" * 30, b"", b"N" * 100, b"ACGT"]
    repart = rng.integers(0, nparts, 4 ** m).astype(np.uint16)
    for pass_ in range(npass):
        a, sa = oracle.superkmers(seqs, k, m, repart, nparts, npass, pass_)
        b, sb = reference.superkmers(seqs, k, m, repart, nparts, npass, pass_)
        assert sa[0] == sb[0] and sa[1] == sb[1]
        for p in range(nparts):
            # the reference writes one file per partition in sequence order; with a single thread that order is deterministic
            assert a[p] == b[p], (pass_, p)


@pytest.mark.parametrize("k,m,nreads,L,cores", [(21, 10, 3000, 100, 1), (31, 10, 4000, 150, 4), (63, 10, 1500, 250, 2), (31, 8, 2000, 150, 1)])
def test_full_dsk_against_reference(oracle, reference, k, m, nreads, L, cores):
    codes = oracle.synth_reads(11 + k, nreads * L // 30, 0, nreads, L).reshape(nreads, L)
    seqs = [oracle.codes_to_ascii(r) for r in codes]
    seqs[5] = seqs[5][:40] + b"N" + seqs[5][41:]          # an invalid character somewhere
    with tempfile.TemporaryDirectory() as tmp:
        fa = os.path.join(tmp, "reads.fa")
        write_fasta(fa, seqs)
        ref = reference.dsk(fa, k, m, abundance_min=2, nb_cores=cores)
    nparts, npass = ref["nb_partitions"], ref["nb_passes"]
    res = oracle.dsk(seqs, k, m, ref["repart"], nparts, abundance_min=2, nb_passes=npass)
    assert res["stats"][0] == ref["kmers_nb_valid"] and res["stats"][1] == ref["kmers_nb_invalid"]
    assert res["stats"][2] == ref["nb_distinct"]
    abund = []
    for key in range(nparts * npass):
        lo, hi, cn = res["parts"][key]
        rlo, rhi, rcn = ref["parts"][key]
        assert (lo == rlo).all() and (hi == rhi).all() and (cn == rcn).all(), key      # same order too (ascending)
        abund.append(rcn)
    table, cutoff, nbsolids, peak = reference.histogram(np.concatenate(abund))
    assert (table == res["histogram"]).all()
    c2, n2, p2 = oracle.histogram_cutoff(res["histogram"])
    assert (c2, n2) == (cutoff, nbsolids)


def test_histogram_cutoff_shapes(oracle, reference):
    rng = np.random.default_rng(9)
    for trial in range(20):
        lam = float(rng.uniform(5, 60))
        ab = np.concatenate([rng.poisson(lam, 20000) + 1, rng.geometric(0.6, int(rng.integers(1000, 60000)))]).astype(np.int32)
        if trial % 5 == 0:
            ab = np.concatenate([ab, np.full(10, 20000, np.int32)])     # beyond histo_max -> clamped
        table, cutoff, nbsolids, peak = reference.histogram(ab)
        hist = np.bincount(np.minimum(ab, 10000), minlength=10001).astype(np.uint64)
        assert (hist == table).all()
        c, n, p = oracle.histogram_cutoff(hist)
        assert (c, n, p) == (cutoff, nbsolids, peak)


@pytest.mark.parametrize("kind", ["basic", "cache", "neighbor"])
@pytest.mark.parametrize("words,k", [(1, 21), (1, 31), (2, 63), (2, 33)])
def test_bloom_bytes_identical(oracle, reference, kind, words, k):
    rng = np.random.default_rng(words * 1000 + k)
    n = 5000
    full = [int(rng.integers(0, 2 ** 63)) * (2 ** 63) * 4 + int(rng.integers(0, 2 ** 63)) for _ in range(n)]
    full = [v & (4 ** k - 1) for v in full]
    lo = np.array([v & (2 ** 64 - 1) for v in full], np.uint64)
    hi = np.array([v >> 64 for v in full], np.uint64) if words == 2 else None
    for bit_size in (30000, 2 ** 15, 2 ** 15 - 8192, 12345):      # includes the power-of-two 'tai--' quirk for each kind
        a, abits = oracle.bloom(kind, bit_size, 4, k, words, lo, hi)
        b, bbits = reference.bloom(kind, bit_size, 4, k, words, lo, hi)
        assert abits == bbits and len(a) == len(b)
        assert (a == b).all(), (kind, bit_size)


def test_configuration_arithmetic_matches_the_reference(reference, oracle, tmp_path):
    # bench.py sizes its partitions with oracle_lib.Reference.configuration: it must give what ConfigurationAlgorithm gives
    import oracle_lib
    n, L, k = 30000, 100, 31
    codes = oracle.synth_reads(9, n * L // 30, 0, n, L).reshape(n, L)
    fa = tmp_path / "r.fa"
    with open(fa, "wb") as f:
        for i, r in enumerate(codes):
            f.write(b">r%d\n" % i + oracle.codes_to_ascii(r) + b"\n")
    for cores, mem in ((1, 5000), (4, 5000), (3, 1)):
        res = reference.dsk(str(fa), k, 10, nb_cores=cores, max_memory_mb=mem)
        want = (res["nb_passes"], res["nb_partitions"])
        got = oracle_lib.Reference.configuration(n * (L - k + 1), 8, cores, max_memory_mb=mem, max_disk_mb=10 ** 7, max_open_files=10 ** 6)
        assert got == want, (cores, mem, got, want)
    table = reference.repartition(str(fa), k, 10, 24, 1, 2)
    assert table.shape == (4 ** 10,) and int(table.max()) == 23 and len(np.unique(table)) == 24
