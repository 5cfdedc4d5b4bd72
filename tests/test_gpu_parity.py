"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle / committed reference fixtures /
the reference's golden vectors.  Bit-exact everywhere (integer work)."""
import numpy as np
import pytest

import fixtures
from golden import reference_vectors as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import gatb_core_b200
    g = gatb_core_b200.GatbGpu(0)
    yield g
    g.close()


def pad(a, n=64):
    return np.concatenate([a, np.zeros(n, a.dtype)])


def pack_seqs(oracle, seqs):
    """ASCII sequences -> (packed stream, offsets, n_mask or None) in the C-ABI layout."""
    blob = b"".join(seqs)
    codes = np.frombuffer(blob, np.uint8)
    bad = ~np.isin(codes, np.frombuffer(b"ACGTacgt", np.uint8))
    c2 = (codes >> 1) & 3
    packed = pad(oracle.pack_2bit(c2.astype(np.uint8)))
    offs = np.zeros(len(seqs) + 1, np.uint64)
    offs[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    mask = None
    if bad.any():
        bits = np.zeros(((len(codes) + 31) // 32 + 2) * 32, np.uint8)
        bits[:len(codes)] = bad
        mask = np.packbits(bits.reshape(-1, 8), axis=1, bitorder="little").reshape(-1).view(np.uint32).copy()
    return packed, offs, mask


def rand_seq(rng, n, p_n=0.0):
    s = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)].copy()
    if p_n:
        s[rng.random(n) < p_n] = ord("N")
    return s.tobytes()


def check_parts(got, want_parts, nkeys, words):
    for key in range(nkeys):
        lo, hi, cn = want_parts[key]
        glo, ghi, gcn = got["parts"][key]
        assert len(glo) == len(lo), (key, len(glo), len(lo))
        assert (glo == lo).all() and (gcn == cn).all(), key
        if words == 2:
            assert (ghi == hi).all(), key


# ------------------------------------------------------------------------------------------------------------------
def test_synth_generator_matches_oracle(gpu, oracle):
    n, L = 3000, 150
    for seed, first in ((1, 0), (42, 12345)):
        codes = oracle.synth_reads(seed, 100000, first, n, L)
        want = oracle.pack_2bit(codes)
        d = gpu.malloc(len(want) + 64)
        gpu.synth_reads_dev(seed, 100000, first, n, L, d)
        got = np.zeros((len(want) + 3) // 4 * 4, np.uint8)
        gpu.d2h(got, d)
        gpu.free(d)
        assert (got[:len(want)] == want).all()


def test_pack_ascii_matches_reference_encoding(gpu, oracle):
    rng = np.random.default_rng(0)
    s = rand_seq(rng, 10007, 0.01) + b"acgtnRYK"
    packed, mask, bad = gpu.pack_ascii(s)
    want, offs, wmask = pack_seqs(oracle, [s])
    n = (len(s) + 3) // 4
    assert (packed[:n] == want[:n]).all()
    assert bad == sum(c not in b"ACGTacgt" for c in s)
    assert (mask[:len(wmask) - 2] == wmask[:len(wmask) - 2]).all()


@pytest.mark.parametrize("name", fixtures.NAMES)
@pytest.mark.parametrize("emit_all", [False, True])
def test_dsk_against_reference_fixtures(gpu, oracle, name, emit_all):
    fx = fixtures.Fixture(name, oracle)
    packed, offs, mask = pack_seqs(oracle, fx.seqs)
    params = gpu.make_params(fx.k, fx.m, nb_partitions=fx.nb_partitions, nb_passes=fx.nb_passes,
                             abundance_min=fx.abundance_min, emit_all=emit_all)
    got = gpu.count(packed, offs, len(fx.seqs), params, repart=fx.repart, n_mask=mask)
    nkeys = fx.nb_partitions * fx.nb_passes
    want = {key: (fx.part(key) if emit_all else fx.solid(key)) for key in range(nkeys)}
    check_parts(got, want, nkeys, fx.words)
    assert (got["histogram"] == fx.z["histogram"]).all()
    st = got["stats"]
    assert [st["kmers_nb_valid"], st["kmers_nb_invalid"], st["kmers_nb_distinct"], st["kmers_nb_solid"]] == fx.z["stats"].tolist()
    assert list(gpu.histogram_cutoff(got["histogram"])) == fx.z["cutoff"].tolist()


@pytest.mark.parametrize("name", fixtures.NAMES)
def test_fixed_length_fast_path_equals_offsets_path(gpu, oracle, name):
    fx = fixtures.Fixture(name, oracle)
    if len(fx.z["n_positions"]):
        pytest.skip("fixture has invalid nucleotides")
    packed, offs, _ = pack_seqs(oracle, fx.seqs)
    p1 = gpu.make_params(fx.k, fx.m, nb_partitions=fx.nb_partitions, abundance_min=2)
    p2 = gpu.make_params(fx.k, fx.m, nb_partitions=fx.nb_partitions, abundance_min=2, read_len=fx.L)
    a = gpu.count(packed, offs, len(fx.seqs), p1, repart=fx.repart)
    b = gpu.count(packed, None, len(fx.seqs), p2, repart=fx.repart)
    check_parts(a, b["parts"], fx.nb_partitions, fx.words)


@pytest.mark.parametrize("table_log2", [5, 7, 10])
def test_small_tables_force_the_global_fallback(gpu, oracle, table_log2):
    fx = fixtures.Fixture("dsk_k31_parts", oracle)
    packed, offs, mask = pack_seqs(oracle, fx.seqs)
    params = gpu.make_params(fx.k, fx.m, nb_partitions=fx.nb_partitions, abundance_min=2, table_log2=table_log2)
    got = gpu.count(packed, offs, len(fx.seqs), params, repart=fx.repart, n_mask=mask)
    check_parts(got, {key: fx.solid(key) for key in range(fx.nb_partitions)}, fx.nb_partitions, 1)
    assert (got["histogram"] == fx.z["histogram"]).all()


# every alternate code path behind gatb_gpu_params.path_flags (one process, no environment variables): general partition
# kernel, CTA-per-bin counting kernels, one k-mer per lane, no second tier, exact two-pass bucket scatter, a directory too
# small for the pooled scatter, the register scanner without orientation, and other bin loads
def path_variants():
    import gatb_core_b200 as g
    F = g.PATH_FUSED
    return [("general_k1", dict(path_flags=g.PATH_K1_GENERAL)), ("general_k1_fused", dict(path_flags=g.PATH_K1_GENERAL | F)),
            ("k1_tma_staging", dict(path_flags=g.PATH_K1_STAGING)), ("fused", dict(path_flags=F)), ("fused_no_dedup", dict(path_flags=F | g.PATH_NO_DEDUP)),
            ("fused_canonical", dict(path_flags=F | g.PATH_CANONICAL)),
            ("fused_overflow_to_tiers", dict(path_flags=F, table_log2=7)), ("fused_overflow_to_global", dict(table_log2=5, path_flags=F | g.PATH_NO_TIER2)),
            ("fused_dense", dict(path_flags=F, bin_load_pct=350)), ("cta128", dict(path_flags=g.PATH_K2B_CTA128)),
            ("cta256", dict(path_flags=g.PATH_K2B_CTA256)), ("lane", dict(path_flags=g.PATH_K2B_LANE)),
            ("no_tier2_small_table", dict(path_flags=g.PATH_NO_TIER2, table_log2=6)), ("k3_no_pool", dict(path_flags=g.PATH_K3_NO_POOL)),
            ("k3_tiny_directory", dict(k3_dir_rounds=1)), ("canonical_records", dict(path_flags=g.PATH_CANONICAL)),
            ("canonical_cta128", dict(path_flags=g.PATH_CANONICAL | g.PATH_K2B_CTA128)),
            ("no_dedup", dict(path_flags=g.PATH_NO_DEDUP)), ("no_dedup_canonical", dict(path_flags=g.PATH_NO_DEDUP | g.PATH_CANONICAL)),
            ("fine_bits_5", dict(fine_bits=5)), ("fine_bits_9_dense", dict(fine_bits=9, bin_load_pct=120)), ("fine_bits_12", dict(fine_bits=12)),
            ("bin_target_tiny", dict(bin_target_pct=3)), ("bin_target_beyond_table", dict(bin_target_pct=300)),
            ("k2a_presplit", dict(path_flags=g.PATH_K2A_PRESPLIT)), ("k2a_presplit_small_stage", dict(path_flags=g.PATH_K2A_PRESPLIT | g.PATH_K2A_SMALL_STAGE, bin_load_pct=30)),
            ("k2a_several_passes", dict(path_flags=g.PATH_K2A_SMALL_STAGE, bin_load_pct=30)),
            ("k2a_passes_then_fallback", dict(path_flags=g.PATH_K2A_SMALL_STAGE, bin_load_pct=50, fine_bits=3)),
            ("k2a_two_pass_fallback", dict(path_flags=g.PATH_K2A_SMALL_STAGE, bin_load_pct=300)),
            ("dense_bins", dict(bin_load_pct=150)), ("sparse_bins", dict(bin_load_pct=10)), ("tier_tables", dict(table_log2=6))]


@pytest.mark.parametrize("variant", [v[0] for v in path_variants()])
@pytest.mark.parametrize("name", ["dsk_k31_parts", "dsk_k21_cfg1"])
def test_alternate_code_paths(gpu, oracle, name, variant):
    fx = fixtures.Fixture(name, oracle)
    packed, offs, mask = pack_seqs(oracle, fx.seqs)
    kw = dict(path_variants())[variant]
    params = gpu.make_params(fx.k, fx.m, nb_partitions=fx.nb_partitions, nb_passes=fx.nb_passes, abundance_min=fx.abundance_min, **kw)
    got = gpu.count(packed, offs, len(fx.seqs), params, repart=fx.repart, n_mask=mask)
    nkeys = fx.nb_partitions * fx.nb_passes
    check_parts(got, {key: fx.solid(key) for key in range(nkeys)}, nkeys, fx.words)
    assert (got["histogram"] == fx.z["histogram"]).all()


def test_k63_warp_per_bin_kernel(gpu, oracle):
    import gatb_core_b200 as g
    fx = fixtures.Fixture("dsk_k63_w16", oracle)
    packed, offs, mask = pack_seqs(oracle, fx.seqs)
    params = gpu.make_params(fx.k, fx.m, nb_partitions=fx.nb_partitions, nb_passes=fx.nb_passes, abundance_min=fx.abundance_min,
                             path_flags=g.PATH_K2B_W2_WARP)
    got = gpu.count(packed, offs, len(fx.seqs), params, repart=fx.repart, n_mask=mask)
    nkeys = fx.nb_partitions * fx.nb_passes
    check_parts(got, {key: fx.solid(key) for key in range(nkeys)}, nkeys, 2)
    assert (got["histogram"] == fx.z["histogram"]).all()


def revcomp_ascii(s):
    return s[::-1].translate(bytes.maketrans(b"ACGT", b"TGCA"))


@pytest.mark.parametrize("k,m", [(31, 10), (23, 8), (27, 9), (15, 7)])
def test_oriented_records_on_both_strands_hairpins_and_palindromes(gpu, oracle, k, m):
    # the oriented partition (k1_scan.cuh) must give every k-mer ONE representative whatever the strand it is read from:
    # reads from both strands of a genome rich in hairpins (a minimizer and its reverse complement inside one k-mer),
    # palindromic m-mers, tandem repeats and homopolymers; ambiguous k-mers take the exact slow path
    rng = np.random.default_rng(k * 100 + m)
    g = bytearray(rand_seq(rng, 6000))
    for i in range(100, 5800, 97):
        kind = (i // 97) % 4
        if kind == 0:
            g[i:i + 24] = revcomp_ascii(bytes(g[i - 24:i]))                  # hairpin
        elif kind == 1:
            h = bytes(g[i:i + 8]); g[i + 8:i + 16] = revcomp_ascii(h)        # palindromic 16-mer
        elif kind == 2:
            g[i:i + 40] = bytes(g[i - 5:i]) * 8                              # tandem repeat
        else:
            g[i:i + 40] = b"AT" * 20 if (i // 97) % 8 == 3 else b"A" * 40
    g = bytes(g)
    seqs = []
    for i in range(1500):
        a = int(rng.integers(0, len(g) - 200)); n = int(rng.integers(k, 200))
        sq = g[a:a + n]
        seqs.append(revcomp_ascii(sq) if i % 2 else sq)
    packed, offs, mask = pack_seqs(oracle, seqs)
    want = oracle.dsk(seqs, k, m, np.zeros(4 ** m, np.uint16), 1, abundance_min=1)
    for kw in (dict(), dict(table_log2=6), dict(bin_load_pct=200)):
        got = gpu.count(packed, offs, len(seqs), gpu.make_params(k, m, abundance_min=1, **kw), n_mask=mask)
        check_parts(got, want["solid"], 1, 1)
        assert (got["histogram"] == want["histogram"]).all()
        assert got["stats"]["kmers_nb_distinct"] == int(want["stats"][2])


@pytest.mark.parametrize("name", ["dsk_k31_parts", "dsk_k63_w16"])
def test_streaming_input_equals_one_shot(gpu, oracle, name):
    # gatb_gpu_reads_begin / _push_ascii / _count: ragged batches of ASCII sequences (with N), packed on the device at
    # arbitrary nucleotide offsets, must give the same result as the one-shot host call
    fx = fixtures.Fixture(name, oracle)
    rng = np.random.default_rng(3)
    seqs = list(fx.seqs) + [b"", b"ACGTN" * 7, b"acgtacgtacgtacgtacgtacgtacgtacgtacgtacgtacgtacgtacgtacgtacgtacgtacgt", b"A"]
    params = gpu.make_params(fx.k, fx.m, nb_partitions=fx.nb_partitions, nb_passes=fx.nb_passes, abundance_min=fx.abundance_min)
    packed, offs, mask = pack_seqs(oracle, seqs)
    want = gpu.count(packed, offs, len(seqs), params, repart=fx.repart, n_mask=mask)
    cuts = sorted(set([0, len(seqs)] + [int(c) for c in rng.integers(0, len(seqs), 9)]))
    batches = [seqs[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
    got = gpu.count_pushed(batches, params, repart=fx.repart)
    nkeys = fx.nb_partitions * fx.nb_passes
    check_parts(got, want["parts"], nkeys, fx.words)
    assert (got["histogram"] == want["histogram"]).all()
    assert got["stats"] == want["stats"] or all(got["stats"][k] == want["stats"][k] for k in ("kmers_nb_valid", "kmers_nb_invalid", "kmers_nb_distinct", "kmers_nb_solid", "sequences", "nucleotides"))


@pytest.mark.parametrize("fmt", ["fasta", "fastq"])
def test_text_parsed_on_the_device_equals_packed_input(gpu, oracle, fmt):
    # gatb_gpu_reads_push_text: FASTA (multi-line records, CRLF, lower case, N, empty records, blank lines) and four-line FASTQ,
    # in several batches cut at record boundaries, against the same sequences packed on the host
    fx = fixtures.Fixture("dsk_k31_parts", oracle)
    rng = np.random.default_rng(11)
    seqs = [s if i % 5 else s.lower() for i, s in enumerate(fx.seqs)] + [b"", b"ACGTNNNNACGT" * 9, b"a", b"ACGT" * 40]
    recs = []
    for i, sq in enumerate(seqs):
        if fmt == "fasta":
            w = int(rng.integers(20, 90)) if i % 3 == 0 else 10 ** 6
            eol = b"\r\n" if i % 7 == 0 else b"\n"
            body = b"".join(sq[a:a + w] + eol for a in range(0, len(sq), w)) if sq else b""
            recs.append(b">read_%d some text > with @ signs" % i + eol + body + (b"\n" if i % 13 == 0 else b""))
        else:
            recs.append(b"@read_%d\n" % i + sq + b"\n+\n" + bytes(rng.integers(33, 74, len(sq), dtype=np.uint8)) + b"\n")
    if fmt == "fasta":
        recs[-1] = recs[-1].rstrip(b"\n")                     # a file that does not end with a newline
    cuts = sorted(set([0, len(recs)] + [int(c) for c in rng.integers(0, len(recs), 5)]))
    batches = [b"".join(recs[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
    params = gpu.make_params(fx.k, fx.m, nb_partitions=fx.nb_partitions, nb_passes=fx.nb_passes, abundance_min=fx.abundance_min)
    got, info = gpu.count_text(batches, 0 if fmt == "fasta" else 1, params, repart=fx.repart)
    up = [s.upper() for s in seqs]
    packed, offs, mask = pack_seqs(oracle, up)
    want = gpu.count(packed, offs, len(up), params, repart=fx.repart, n_mask=mask)
    nkeys = fx.nb_partitions * fx.nb_passes
    check_parts(got, want["parts"], nkeys, fx.words)
    assert (got["histogram"] == want["histogram"]).all()
    lens = [len(s) for s in seqs]
    assert info[:4] == [len(seqs), sum(lens), min(lens), max(lens)] and info[4] == float(sum(l * l for l in lens))
    assert info[5] == sum(c not in b"ACGTacgt" for s in seqs for c in s)
    assert got["stats"]["sequences"] == len(seqs) and got["stats"]["kmers_nb_invalid"] == want["stats"]["kmers_nb_invalid"]


def test_reference_golden_vectors_dsk(gpu, oracle):
    # TestDSK.cpp:147-241 (solid counts) and :244-341 (exact set + checksum)
    for seqs, k, nks, expected in G.DSK1_CASES:
        m = min(8, k - 1)
        packed, offs, mask = pack_seqs(oracle, [s.encode() for s in seqs])
        got = gpu.count(packed, offs, len(seqs), gpu.make_params(k, m, abundance_min=nks))
        assert got["stats"]["kmers_nb_solid"] == expected and got["n_items"] == expected, (k, nks)
    packed, offs, mask = pack_seqs(oracle, [G.DSK2_SEQ.encode()])
    got = gpu.count(packed, offs, 1, gpu.make_params(31, 10, abundance_min=1))
    lo = got["parts"][0][0]
    assert lo.tolist() == sorted(G.DSK2_SOLID)
    assert sum(lo.tolist()) % 2 ** 64 == G.DSK2_CHECKSUM


# every window size the register scanner is compiled for (k-m+1 = 8, 16, ..., 48), their edges, and the k values that fall
# back to the general partition kernel (k < 15); ragged reads, some with invalid nucleotides, several partitions
@pytest.mark.parametrize("k", [5, 14, 15, 16, 22, 23, 24, 27, 32, 33, 39, 40, 47, 48, 55, 56, 62, 63])
def test_dsk_every_kmer_size_class(gpu, oracle, k):
    seqs, m, nparts, repart = kmer_class_case(k)
    packed, offs, mask = pack_seqs(oracle, seqs)
    want = oracle.dsk(seqs, k, m, repart, nparts, abundance_min=2)
    got = gpu.count(packed, offs, len(seqs), gpu.make_params(k, m, nb_partitions=nparts, abundance_min=2), repart=repart, n_mask=mask)
    check_parts(got, want["solid"], nparts, 1 if k < 32 else 2)
    assert (got["histogram"] == want["histogram"]).all()
    assert got["stats"]["kmers_nb_valid"] == int(want["stats"][0]) and got["stats"]["kmers_nb_invalid"] == int(want["stats"][1])
    assert got["stats"]["kmers_nb_distinct"] == int(want["stats"][2])


def kmer_class_case(k):
    rng = np.random.default_rng(1000 + k)
    m = min(8, k - 1)
    nparts = 3
    base = rand_seq(rng, 3000)                                         # reads drawn from one short genome: real multiplicities
    seqs = []
    for i in range(260):
        a = int(rng.integers(0, 2600)); n = int(rng.integers(1, 400))
        sq = bytearray(base[a:a + n])
        if i % 3 == 0 and len(sq) > 3:
            sq[int(rng.integers(0, len(sq)))] = ord("N")
        seqs.append(bytes(sq))
    repart = (np.arange(4 ** m) * 2654435761 % nparts).astype(np.uint16)
    return seqs, m, nparts, repart


def test_edge_cases(gpu, oracle):
    rng = np.random.default_rng(5)
    k, m = 31, 10
    cases = {
        "empty": [],
        "all_short": [b"ACGT" * 5, b"A" * 30, b""],
        "exactly_k": [rand_seq(rng, k)],
        "only_N": [b"N" * 100, b"ACGTN" * 40],
        "identical_reads": [rand_seq(np.random.default_rng(1), 150)] * 3000,      # extreme skew: every record in a few bins
        "poly_A": [b"A" * 500, b"T" * 500, b"AC" * 300],
        "ragged": [rand_seq(rng, int(n), 0.02) for n in rng.integers(1, 700, 400)],
        "long_read": [rand_seq(rng, 200000)],
    }
    for name, seqs in cases.items():
        packed, offs, mask = pack_seqs(oracle, seqs) if seqs else (np.zeros(64, np.uint8), np.zeros(1, np.uint64), None)
        want = oracle.dsk(seqs, k, m, np.zeros(4 ** m, np.uint16), 1, abundance_min=1) if seqs else None
        got = gpu.count(packed, offs, len(seqs), gpu.make_params(k, m, abundance_min=1), n_mask=mask)
        if want is None:
            assert got["n_items"] == 0 and not got["histogram"].any()
            continue
        check_parts(got, want["solid"], 1, 1)
        assert (got["histogram"] == want["histogram"]).all(), name
        assert got["stats"]["kmers_nb_valid"] == int(want["stats"][0]) and got["stats"]["kmers_nb_invalid"] == int(want["stats"][1]), name


@pytest.mark.parametrize("k,m,nparts,npass", [(31, 10, 7, 1), (21, 8, 3, 2), (63, 10, 5, 1), (12, 6, 2, 1), (40, 8, 4, 3)])
def test_gatb_exact_superkmers(gpu, oracle, k, m, nparts, npass):
    # rows A3-A6: same super-k-mers, same partition, same serialised bytes as the reference (multiset per key)
    import oracle_lib
    rng = np.random.default_rng(k + nparts)
    seqs = [rand_seq(rng, int(rng.integers(20, 400)), 0.01 if i % 3 == 0 else 0.0) for i in range(300)]
    seqs += [b"A" * 300, b"ACGTN" * 30, b"", b"N" * 100, b"ACGT"]
    repart = rng.integers(0, nparts, 4 ** m).astype(np.uint16)
    packed, offs, mask = pack_seqs(oracle, seqs)
    params = gpu.make_params(k, m, nb_partitions=nparts, nb_passes=npass)
    streams, stats = gpu.superkmers(packed, offs, len(seqs), params, repart=repart, n_mask=mask)
    tot_sk = 0
    for pass_ in range(npass):
        want, wst = oracle.superkmers(seqs, k, m, repart, nparts, npass, pass_)
        tot_sk += int(wst[0])
        for p in range(nparts):
            assert oracle_lib.split_records(streams[pass_ * nparts + p], k) == oracle_lib.split_records(want[p], k), (pass_, p)
    assert int(stats[0]) == tot_sk


@pytest.mark.parametrize("kind", ["basic", "cache", "neighbor"])
@pytest.mark.parametrize("words,k", [(1, 21), (1, 31), (2, 63), (2, 33)])
def test_bloom_bytes(gpu, oracle, kind, words, k):
    rng = np.random.default_rng(words * 1000 + k)
    n = 20000
    full = [(int(rng.integers(0, 2 ** 63)) * (2 ** 65) + int(rng.integers(0, 2 ** 63)) * 4 + int(rng.integers(0, 4))) & (4 ** k - 1) for _ in range(n)]
    lo = np.array([v & (2 ** 64 - 1) for v in full], np.uint64)
    hi = np.array([v >> 64 for v in full], np.uint64) if words == 2 else None
    for bit_size in (120000, 2 ** 17, 2 ** 17 - 8192, 12345):
        got, gbits = gpu.bloom(kind, bit_size, 4, k, lo, hi)
        want, wbits = oracle.bloom(kind, bit_size, 4, k, words, lo, hi)
        assert gbits == wbits and len(got) == len(want)
        assert (got == want).all(), (kind, bit_size)


@pytest.mark.parametrize("name", fixtures.NAMES)
def test_bloom_of_fixture_solid_kmers(gpu, oracle, name):
    fx = fixtures.Fixture(name, oracle)
    nkeys = fx.nb_partitions * fx.nb_passes
    lo = np.concatenate([fx.solid(key)[0] for key in range(nkeys)])
    hi = np.concatenate([fx.solid(key)[1] for key in range(nkeys)])
    size, nh = gpu.bloom_params(fx.k, len(lo))
    assert [size, nh] == fx.z["bloom_size"].tolist()
    for kind in ("basic", "cache", "neighbor"):
        got, bits = gpu.bloom(kind, size, nh, fx.k, lo, hi if fx.words == 2 else None)
        assert bits == int(fx.z["bloom_%s_bitsize" % kind][0])
        assert (got == fx.z["bloom_" + kind]).all(), kind


def test_device_resident_api_and_size_independent_properties(gpu, oracle):
    # 2e5 reads against the oracle, then 2e6 reads through invariants only
    k, m, L = 31, 10, 150
    for n, check_oracle in ((200000, True), (2000000, False)):
        genome = n * L // 30
        nbytes = (n * L + 3) // 4
        d = gpu.malloc(nbytes + 64)
        gpu.synth_reads_dev(42, genome, 0, n, L, d)
        params = gpu.make_params(k, m, abundance_min=2, read_len=L)
        res = gpu.count_dev(d, None, n, params)
        got = gpu.result_to_host(res, params)
        gpu.result_free(res)
        gpu.free(d)
        lo, hi, cn = got["parts"][0]
        h = got["histogram"].astype(object)
        assert (np.diff(lo.astype(object)) > 0).all() if len(lo) < 10 ** 6 else (lo[1:] > lo[:-1]).all()     # strictly ascending
        assert sum(h) == got["stats"]["kmers_nb_distinct"]
        assert sum(int(i) * int(v) for i, v in enumerate(h) if i < 10000) + 0 == got["stats"]["kmers_nb_valid"] or h[10000] > 0
        assert got["stats"]["kmers_nb_valid"] == n * (L - k + 1)
        assert (cn >= 2).all() and len(lo) == got["stats"]["kmers_nb_solid"]
        if check_oracle:
            codes = oracle.synth_reads(42, genome, 0, n, L).reshape(n, L)
            want = oracle.dsk([oracle.codes_to_ascii(r) for r in codes], k, m, np.zeros(4 ** m, np.uint16), 1, abundance_min=2)
            check_parts(got, want["solid"], 1, 1)
            assert (got["histogram"] == want["histogram"]).all()


def test_smoke_entry(gpu):
    import __graft_entry__
    __graft_entry__.smoke()


def test_staged_multi_gpu_path_on_one_rank(gpu, oracle):
    # the staged entry points (plan / partition_into / exchange / count_bins) with world = 1 must equal the one-shot call
    import torch
    from gatb_core_b200 import multigpu
    fx = fixtures.Fixture("dsk_k31_parts", oracle)
    packed, offs, mask = pack_seqs(oracle, fx.seqs)
    if mask is not None:                                      # staged path without the N mask: use clean reads
        seqs = [s.replace(b"N", b"A") for s in fx.seqs]
        packed, offs, mask = pack_seqs(oracle, seqs)
    else:
        seqs = fx.seqs
    params = gpu.make_params(fx.k, fx.m, nb_partitions=fx.nb_partitions, abundance_min=2)
    want = gpu.count(packed, offs, len(seqs), params, repart=fx.repart)
    d_reads = torch.from_numpy(packed).cuda()
    d_offs = torch.from_numpy(offs.astype(np.int64)).cuda()
    total_kmers = sum(max(len(s) - fx.k + 1, 0) for s in seqs)
    res, stats = multigpu.count_distributed(gpu, params, d_reads.data_ptr(), len(seqs), len(seqs), total_kmers, 0, 1,
                                            repart=fx.repart, d_offsets=d_offs.data_ptr())
    got = gpu.result_to_host(res, params)
    gpu.result_free(res)
    check_parts(got, want["parts"], fx.nb_partitions, 1)
    assert (got["histogram"] == want["histogram"]).all()
    assert stats["kmers_nb_distinct"] == want["stats"]["kmers_nb_distinct"]


@pytest.mark.parametrize("name", ["dsk_k31_parts", "dsk_k63_w16", "dsk_k21_cfg1"])
def test_several_devices_in_one_process(gpu, oracle, name):
    # gatb_gpu_count_multi: one host thread per device, peer copies of the bin regions, host merge == the single-device count
    import torch
    import gatb_core_b200
    n_dev = min(torch.cuda.device_count(), 4)
    if n_dev < 2:
        pytest.skip("needs at least two GPUs")
    fx = fixtures.Fixture(name, oracle)
    packed, offs, mask = pack_seqs(oracle, fx.seqs)
    W = 1 if fx.k < 32 else 2
    params = gpu.make_params(fx.k, fx.m, nb_partitions=fx.nb_partitions, nb_passes=fx.nb_passes, abundance_min=fx.abundance_min)
    want = gpu.count(packed, offs, len(fx.seqs), params, repart=fx.repart, n_mask=mask)
    others = [gatb_core_b200.GatbGpu(d) for d in range(1, n_dev)]
    try:
        got = gpu.count_multi(others, packed, offs, len(fx.seqs), params, repart=fx.repart, n_mask=mask)
    finally:
        for o in others:
            o.close()
    check_parts(got, want["parts"], fx.nb_partitions * fx.nb_passes, W)
    assert (got["histogram"] == want["histogram"]).all()
    for key in ("kmers_nb_valid", "kmers_nb_invalid", "kmers_nb_distinct", "kmers_nb_solid"):
        assert got["stats"][key] == want["stats"][key], key


@pytest.mark.parametrize("name,world", [("dsk_k31_parts", 3), ("dsk_k63_w16", 2), ("dsk_k21_cfg1", 8)])
def test_routed_result_on_one_gpu(gpu, oracle, name, world):
    # second exchange of a multi-GPU run (gatb_gpu_count_bins_routed + gatb_gpu_sort_routed), all ranks played by one GPU: the items
    # routed to destination r, sorted, must be exactly the partitions key % world == r of the one-shot count, every other key empty
    import torch
    from gatb_core_b200.multigpu import _as_tensor
    fx = fixtures.Fixture(name, oracle)
    seqs = [s.replace(b"N", b"A") for s in fx.seqs]
    packed, offs, _ = pack_seqs(oracle, seqs)
    W = 1 if fx.k < 32 else 2
    n_keys = fx.nb_partitions * fx.nb_passes
    params = gpu.make_params(fx.k, fx.m, nb_partitions=fx.nb_partitions, nb_passes=fx.nb_passes, abundance_min=2)
    want = gpu.count(packed, offs, len(seqs), params, repart=fx.repart)
    dev = torch.device("cuda", 0)
    d_reads = torch.from_numpy(packed).cuda()
    d_offs = torch.from_numpy(offs.astype(np.int64)).cuda()
    total_kmers = sum(max(len(s) - fx.k + 1, 0) for s in seqs)
    geom = gpu.plan(params, total_kmers, len(seqs), 1)
    bins = torch.empty(geom.nb1 * geom.cap * geom.record_bytes, dtype=torch.uint8, device=dev)
    cursors = torch.zeros(geom.nb1, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    st = gpu.partition_into(params, geom, d_reads.data_ptr(), d_offs.data_ptr(), len(seqs), bins.data_ptr(), cursors.data_ptr())
    assert st[3] == 0
    res1, d_keys, send, cap = gpu.count_bins_routed(params, geom, [bins.data_ptr()], [cursors.data_ptr()], geom.nb1, total_kmers, world, repart=fx.repart)
    assert sum(send) == int(res1.n_items) == want["stats"]["kmers_nb_solid"]
    lo = _as_tensor(res1.kmers_lo, cap * world, torch.int64, dev).clone()
    hi = _as_tensor(res1.kmers_hi, cap * world, torch.int64, dev).clone() if W == 2 else None
    cn = _as_tensor(res1.counts, cap * world, torch.int32, dev).clone()
    ky = _as_tensor(d_keys, cap * world * 2, torch.uint8, dev).clone()
    seen = 0
    for r in range(world):
        a, b = r * cap, r * cap + send[r]
        r_lo, r_cn, r_ky = lo[a:b].clone(), cn[a:b].clone(), ky[2 * a:2 * b].clone()
        r_hi = hi[a:b].clone() if W == 2 else None
        torch.cuda.synchronize()
        res = gpu.sort_routed(params, r_lo.data_ptr(), r_hi.data_ptr() if W == 2 else None, r_cn.data_ptr(), r_ky.data_ptr(), send[r], world, r)
        got = gpu.result_to_host(res, params)
        for key in range(n_keys):
            glo, ghi, gcn = got["parts"][key]
            if key % world != r:
                assert len(glo) == 0, (r, key)
                continue
            wlo, whi, wcn = want["parts"][key]
            assert len(glo) == len(wlo) and (glo == wlo).all() and (gcn == wcn).all(), (r, key)
            if W == 2:
                assert (ghi == whi).all(), (r, key)
            seen += len(glo)
    assert seen == want["stats"]["kmers_nb_solid"]


# ---- BASELINE.json configs 3 and 4 at a down-sampled size, against the REFERENCE itself (oracle/_ref: SortingCountAlgorithm on all
#      host cores with its own configuration and Repartitor table): same generator and seeds as SURVEY.md 8d, n / 1000 ----
@pytest.mark.parametrize("name,k,L,n,seed", [("cfg3_k31", 31, 150, 1000000, 43), ("cfg4_k63", 63, 250, 500000, 44)])
def test_downsampled_configs_against_the_reference(gpu, oracle, reference, tmp_path, name, k, L, n, seed):
    codes = oracle.synth_reads(seed, n * L // 30, 0, n, L)
    fa = tmp_path / "reads.fa"
    with open(fa, "wb") as f:
        block = np.empty((n, 3 + L + 1), np.uint8)
        block[:, :3] = np.frombuffer(b">r\n", np.uint8)
        block[:, 3:3 + L] = np.frombuffer(b"ACTG", np.uint8)[codes].reshape(n, L)
        block[:, -1] = ord("\n")
        f.write(block.tobytes())
    ref = reference.dsk(str(fa), k, 10, abundance_min=2, nb_cores=4)
    nparts, npass = ref["nb_partitions"], ref["nb_passes"]
    packed = pad(oracle.pack_2bit(codes))
    params = gpu.make_params(k, 10, nb_partitions=nparts, nb_passes=npass, abundance_min=2, read_len=L, emit_all=True)
    got = gpu.count(packed, None, n, params, repart=ref["repart"])
    assert got["stats"]["kmers_nb_distinct"] == ref["nb_distinct"]
    for key in range(nparts * npass):
        lo, hi, cn = ref["parts"][key]
        glo, ghi, gcn = got["parts"][key]
        assert len(lo) == len(glo) and (lo == glo).all() and (cn == gcn).all(), (name, key)
        if k > 31:
            assert (hi == ghi).all(), (name, key)


def test_multi_gpu_equals_single_gpu():
    """torchrun of tools/check_multigpu.py on every visible GPU (at most 8): the sharded count, merged per partition, must equal
    the single-GPU count bit for bit, k = 31 and k = 63.  Skipped on a one-GPU box."""
    import os
    import subprocess
    import sys
    import torch
    n = min(torch.cuda.device_count(), 8)
    if n < 2:
        pytest.skip("one GPU visible")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
                        "--master-port", "29733", os.path.join(root, "tools", "check_multigpu.py")], capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("GPUs == 1 GPU") == 2, r.stdout[-2000:]


# ---- BASELINE.json config 5 (metagenome-like Zipf skew) at a reduced size: generator parity, counts against the oracle, overflow
#      statistics, and the Bloom filter of the solid k-mers ----
def test_zipf_metagenome_generator_and_counts(gpu, oracle):
    n_species, n, L, k, m = 300, 60000, 150, 31, 10
    cdf, off = oracle.zipf_tables(45, n_species, 1.1)
    codes = oracle.synth_reads_zipf(45, cdf, off, 1234, n, L)
    want_packed = oracle.pack_2bit(codes)
    d = gpu.malloc(len(want_packed) + 64)
    gpu.synth_zipf_dev(45, cdf, off, 1234, n, L, d)
    got_packed = np.zeros((len(want_packed) + 3) // 4 * 4, np.uint8)
    gpu.d2h(got_packed, d)
    assert (got_packed[:len(want_packed)] == want_packed).all()
    # the species sizes follow the table: the first species holds the Zipf share of the reads
    share0 = float(cdf[0]) / 2.0 ** 64
    assert 0.1 < share0 < 0.3
    params = gpu.make_params(k, m, nb_partitions=4, abundance_min=2, read_len=L)
    repart = (np.arange(4 ** m) * 2654435761 % 4).astype(np.uint16)
    res = gpu.count_dev(d, None, n, params, repart=repart)
    got = gpu.result_to_host(res, params)
    gpu.result_free(res)
    gpu.free(d)
    seqs = [oracle.codes_to_ascii(r) for r in codes.reshape(n, L)]
    want = oracle.dsk(seqs, k, m, repart, 4, abundance_min=2)
    check_parts(got, want["solid"], 4, 1)
    assert (got["histogram"] == want["histogram"]).all()
    # the dominant species is covered several times, most others less than once
    assert max(int(c.max()) for _, _, c in got["parts"].values() if len(c)) > 5
    lo = np.concatenate([got["parts"][key][0] for key in range(4)])
    size, nh = gpu.bloom_params(k, len(lo))
    b, _ = gpu.bloom("neighbor", size, nh, k, lo)
    ob, _ = oracle.bloom("neighbor", size, nh, k, 1, lo)
    assert (b == ob).all()
