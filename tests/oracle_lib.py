"""ctypes bindings for the TEST-ONLY libraries under oracle/:

  liboracle.so          plain-C restatement (oracle/kmer_oracle.c)
  _ref/libgatbref.so    the unmodified reference compiled from /root/reference (oracle/ref_harness.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

u8p, u16p, u32p, u64p, i32p = (np.ctypeslib.ndpointer(dtype=t, flags="C_CONTIGUOUS")
                               for t in (np.uint8, np.uint16, np.uint32, np.uint64, np.int32))
VP = C.c_void_p


def build_oracle():
    """(Re)build liboracle.so, and libgatbref.so when the reference tree is present."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "-j8", "all"], check=True)


def _opt(a):
    return None if a is None else a.ctypes.data_as(VP)


class Oracle:
    def __init__(self):
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        L = self.L = C.CDLL(path)
        L.orc_revcomp.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.orc_hash1.restype = C.c_uint64
        L.orc_hash1.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_uint64]
        L.orc_simplehash16.restype = C.c_uint64
        L.orc_simplehash16.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int]
        L.orc_mmer_lut.argtypes = [C.c_int, u32p]
        L.orc_kmers.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_int, u64p, u64p, u32p, u8p, u8p]
        L.orc_superkmers.argtypes = [C.c_char_p, u64p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, u16p, C.c_int,
                                     C.POINTER(VP), u64p, u64p]
        L.orc_free.argtypes = [VP]
        L.orc_decode_superkmers.restype = C.c_uint64
        L.orc_decode_superkmers.argtypes = [VP, C.c_uint64, C.c_int, VP, VP]
        L.orc_dsk_run.restype = VP
        L.orc_dsk_run.argtypes = [C.c_char_p, u64p, C.c_uint64, C.c_int, C.c_int, C.c_int, u16p, C.c_int, C.c_int,
                                  C.c_int64, C.c_int, C.c_int]
        for f in ("orc_dsk_part_size", "orc_dsk_solid_size"):
            getattr(L, f).restype = C.c_uint64
            getattr(L, f).argtypes = [VP, C.c_uint32]
        for f in ("orc_dsk_get_part", "orc_dsk_get_solid"):
            getattr(L, f).argtypes = [VP, C.c_uint32, VP, VP, VP]
        L.orc_dsk_histogram.restype = C.POINTER(C.c_uint64)
        L.orc_dsk_histogram.argtypes = [VP]
        L.orc_dsk_stats.restype = C.POINTER(C.c_uint64)
        L.orc_dsk_stats.argtypes = [VP]
        L.orc_dsk_free.argtypes = [VP]
        L.orc_histogram_cutoff.argtypes = [u64p, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
        L.orc_nbits_per_kmer.restype = C.c_float
        L.orc_nbits_per_kmer.argtypes = [C.c_int]
        L.orc_bloom_params.argtypes = [C.c_int, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
        L.orc_bloom.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_int, C.c_int, VP, VP, C.c_uint64, VP,
                                C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.orc_splitmix64.restype = C.c_uint64
        L.orc_splitmix64.argtypes = [C.c_uint64]
        L.orc_synth_reads.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, u8p]
        L.orc_pack_2bit.argtypes = [u8p, C.c_uint64, u8p]
        L.orc_zipf_tables.argtypes = [C.c_uint64, C.c_uint64, C.c_double, u64p, u64p]
        L.orc_synth_reads_zipf.argtypes = [C.c_uint64, C.c_uint64, u64p, u64p, C.c_uint64, C.c_uint64, C.c_int, u8p]
        L.orc_codes_to_ascii.argtypes = [u8p, C.c_uint64, VP]

    # ---- helpers -------------------------------------------------------------------------------------
    def revcomp(self, lo, hi, k, words):
        a, b = C.c_uint64(), C.c_uint64()
        self.L.orc_revcomp(lo, hi, k, words, C.byref(a), C.byref(b))
        return a.value, b.value

    def hash1(self, lo, hi, words, seed):
        return self.L.orc_hash1(lo, hi, words, seed)

    def simplehash16(self, lo, hi, words, shift):
        return self.L.orc_simplehash16(lo, hi, words, shift)

    def mmer_lut(self, m):
        out = np.zeros(4 ** m, np.uint32)
        self.L.orc_mmer_lut(m, out)
        return out

    def kmers(self, seq, k, m):
        s = seq.encode() if isinstance(seq, str) else bytes(seq)
        n = max(len(s) - k + 1, 0)
        lo, hi = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        mi, va, st = np.zeros(n, np.uint32), np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        if n:
            self.L.orc_kmers(s, len(s), k, m, lo, hi, mi, va, st)
        return lo, hi, mi, va, st

    def superkmers(self, seqs, k, m, repart, nb_partitions, nb_passes=1, pass_=0):
        """seqs: list of ASCII byte strings.  Returns (list of per-partition record-stream bytes, stats[4])."""
        return _superkmers(self.L.orc_superkmers, self.L.orc_free, seqs, k, m, repart, nb_partitions, nb_passes, pass_, None)

    def decode_superkmers(self, stream, k):
        buf = np.frombuffer(stream, np.uint8)
        n = self.L.orc_decode_superkmers(_opt(buf) if len(buf) else None, len(buf), k, None, None)
        lo, hi = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        if n:
            self.L.orc_decode_superkmers(_opt(buf), len(buf), k, _opt(lo), _opt(hi))
        return lo, hi

    def dsk(self, seqs, k, m, repart, nb_partitions, abundance_min=2, abundance_max=2**31 - 1, histo_max=10000,
            nb_passes=1, nthreads=0):
        """Full DSK.  Returns dict(parts={key:(lo,hi,counts)}, solid={...}, histogram, stats)."""
        blob, offs = concat(seqs)
        h = self.L.orc_dsk_run(blob, offs, len(seqs), k, m, nb_passes, repart, nb_partitions, abundance_min,
                               abundance_max, histo_max, nthreads)
        res = {"parts": {}, "solid": {}}
        for key in range(nb_passes * nb_partitions):
            for name, fsz, fget in (("parts", self.L.orc_dsk_part_size, self.L.orc_dsk_get_part),
                                    ("solid", self.L.orc_dsk_solid_size, self.L.orc_dsk_get_solid)):
                n = fsz(h, key)
                lo, hi, cn = np.zeros(n, np.uint64), np.zeros(n, np.uint64), np.zeros(n, np.int32)
                if n:
                    fget(h, key, _opt(lo), _opt(hi), _opt(cn))
                res[name][key] = (lo, hi, cn)
        res["histogram"] = np.ctypeslib.as_array(self.L.orc_dsk_histogram(h), (histo_max + 1,)).copy()
        res["stats"] = np.ctypeslib.as_array(self.L.orc_dsk_stats(h), (8,)).copy()
        self.L.orc_dsk_free(h)
        return res

    def histogram_cutoff(self, table, min_auto_threshold=3):
        t = np.ascontiguousarray(table, np.uint64)
        c, n, p = C.c_uint32(), C.c_uint64(), C.c_uint32()
        self.L.orc_histogram_cutoff(t, len(t) - 1, min_auto_threshold, C.byref(c), C.byref(n), C.byref(p))
        return c.value, n.value, p.value

    def bloom_params(self, k, nb_solid):
        s, h = C.c_uint64(), C.c_int()
        self.L.orc_bloom_params(k, nb_solid, C.byref(s), C.byref(h))
        return s.value, h.value

    def bloom(self, kind, bit_size, nb_hash, k, words, lo, hi=None):
        return _bloom(self.L.orc_bloom, kind, bit_size, nb_hash, k, words, lo, hi)

    def synth_reads(self, seed, genome_len, first_read, n_reads, L):
        codes = np.zeros(n_reads * L, np.uint8)
        self.L.orc_synth_reads(seed, genome_len, first_read, n_reads, L, codes)
        return codes

    def zipf_tables(self, seed, n_species, exponent=1.1):
        """(cdf thresholds u64[n_species], genome offsets u64[n_species+1]) of the metagenome-like generator (config 5)."""
        cdf, off = np.zeros(n_species, np.uint64), np.zeros(n_species + 1, np.uint64)
        self.L.orc_zipf_tables(seed, n_species, exponent, cdf, off)
        return cdf, off

    def synth_reads_zipf(self, seed, cdf, off, first_read, n_reads, L):
        codes = np.zeros(n_reads * L, np.uint8)
        self.L.orc_synth_reads_zipf(seed, len(cdf), cdf, off, first_read, n_reads, L, codes)
        return codes

    def pack_2bit(self, codes):
        out = np.zeros((len(codes) + 3) // 4, np.uint8)
        self.L.orc_pack_2bit(np.ascontiguousarray(codes, np.uint8), len(codes), out)
        return out

    def codes_to_ascii(self, codes):
        out = np.zeros(len(codes), np.uint8)
        self.L.orc_codes_to_ascii(np.ascontiguousarray(codes, np.uint8), len(codes), _opt(out))
        return out.tobytes()


def concat(seqs):
    offs = np.zeros(len(seqs) + 1, np.uint64)
    offs[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    return b"".join(seqs), offs


def _superkmers(fn, free, seqs, k, m, repart, nb_partitions, nb_passes, pass_, tmpdir):
    blob, offs = concat(seqs)
    ptrs = (VP * nb_partitions)()
    sizes = np.zeros(nb_partitions, np.uint64)
    stats = np.zeros(4, np.uint64)
    rp = np.ascontiguousarray(repart, np.uint16)
    if tmpdir is None:
        rc = fn(blob, offs, len(seqs), k, m, nb_passes, pass_, rp, nb_partitions, ptrs, sizes, stats)
    else:
        rc = fn(blob, offs, len(seqs), k, m, nb_passes, pass_, rp, nb_partitions, tmpdir.encode(), ptrs, sizes, stats)
    assert rc == 0
    out = []
    for p in range(nb_partitions):
        out.append(C.string_at(ptrs[p], int(sizes[p])))
        free(ptrs[p])
    return out, stats


def _bloom(fn, kind, bit_size, nb_hash, k, words, lo, hi):
    lo = np.ascontiguousarray(lo, np.uint64)
    hi = None if hi is None else np.ascontiguousarray(hi, np.uint64)
    nb, bs = C.c_uint64(), C.c_uint64()
    assert fn(kind.encode(), bit_size, nb_hash, k, words, _opt(lo), _opt(hi), len(lo), None, C.byref(nb), C.byref(bs)) == 0
    out = np.zeros(nb.value, np.uint8)
    assert fn(kind.encode(), bit_size, nb_hash, k, words, _opt(lo), _opt(hi), len(lo), _opt(out), C.byref(nb), C.byref(bs)) == 0
    return out, bs.value


def split_records(stream, k):
    """Splits a reference-format record stream ([u8 nbK][ceil((k+nbK-1)/4) bytes]...) into a sorted list of records."""
    out, i = [], 0
    while i < len(stream):
        n = 1 + (k + stream[i] - 1 + 3) // 4
        out.append(stream[i:i + n])
        i += n
    assert i == len(stream)
    return sorted(out)


class Reference:
    """The unmodified reference (oracle/_ref/libgatbref.so).  `available` is False when it was not built."""

    def __init__(self):
        path = os.path.join(ORACLE_DIR, "_ref", "libgatbref.so")
        self.available = os.path.exists(path)
        if not self.available:
            return
        L = self.L = C.CDLL(path)
        L.ref_last_error.restype = C.c_char_p
        L.ref_dsk_run.restype = VP
        L.ref_dsk_run.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int]
        for f, t in (("ref_dsk_nb_partitions", C.c_int), ("ref_dsk_nb_passes", C.c_int), ("ref_dsk_seconds", C.c_double),
                     ("ref_dsk_fill_partitions_seconds", C.c_double), ("ref_dsk_fill_solid_seconds", C.c_double),
                     ("ref_dsk_kmers_nb_valid", C.c_uint64), ("ref_dsk_kmers_nb_invalid", C.c_uint64),
                     ("ref_dsk_nb_distinct", C.c_uint64), ("ref_dsk_info_xml", C.c_char_p)):
            getattr(L, f).restype = t
            getattr(L, f).argtypes = [VP]
        L.ref_dsk_part_size.restype = C.c_uint64
        L.ref_dsk_part_size.argtypes = [VP, C.c_uint32]
        L.ref_dsk_get_part.argtypes = [VP, C.c_uint32, VP, VP, VP]
        L.ref_dsk_get_repart.argtypes = [VP, u16p]
        L.ref_dsk_free.argtypes = [VP]
        L.ref_repartition.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, u16p]
        L.ref_histogram.argtypes = [i32p, C.c_uint64, C.c_int, C.c_int, u64p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
        L.ref_kmers.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_int, u64p, u64p, u32p, u8p, u8p]
        L.ref_superkmers.argtypes = [C.c_char_p, u64p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, u16p, C.c_int,
                                     C.c_char_p, C.POINTER(VP), u64p, u64p]
        L.ref_free.argtypes = [VP]
        L.ref_revcomp.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.ref_hash1.restype = C.c_uint64
        L.ref_hash1.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_uint64]
        L.ref_simplehash16.restype = C.c_uint64
        L.ref_simplehash16.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int]
        L.ref_nbits_per_kmer.restype = C.c_float
        L.ref_nbits_per_kmer.argtypes = [C.c_int]
        L.ref_bloom.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_int, C.c_int, VP, VP, C.c_uint64, VP,
                                C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]

    def repartition(self, fasta_path, k, m, nb_partitions, nb_passes=1, nb_cores=1):
        """The reference's Repartitor table (u16[4^m]) for a given partition count, sampled from the bank by RepartitorAlgorithm."""
        table = np.zeros(4 ** m, np.uint16)
        with tempfile.TemporaryDirectory() as tmp:
            cwd = os.getcwd()
            os.chdir(tmp)
            try:
                rc = self.L.ref_repartition(os.fsencode(fasta_path), k, m, nb_partitions, nb_passes, nb_cores,
                                            os.fsencode(os.path.join(tmp, "repart")), table)
            finally:
                os.chdir(cwd)
        if rc:
            raise RuntimeError(self.L.ref_last_error().decode())
        return table

    @staticmethod
    def configuration(n_kmers, kmer_bytes=8, nb_cores=1, max_memory_mb=5000, max_disk_mb=None, max_open_files=None):
        """nb_passes / nb_partitions as ConfigurationAlgorithm computes them (kmer/impl/ConfigurationAlgorithm.cpp:316-417):
        volume = kmers * sizeof(Type) / MB; volume_minim = volume * 0.5 * 1.2; nb_passes = (volume/4)/max_disk + 1;
        nb_partitions = (volume_per_pass * nb_partitions_in_parallel) / max_memory + 1, halving the partitions in parallel
        (then adding passes) while the count reaches the open-file limit; finally rounded up to a multiple of the partitions
        in parallel.  Defaults: 5000 MB, all cores in parallel."""
        import resource
        volume = max(n_kmers * kmer_bytes // (1 << 20), 1)
        volume_minim = max(int(volume * 0.5 * 1.2), 1)
        if max_disk_mb is None:
            st = os.statvfs(os.getcwd())
            avail = st.f_bavail * st.f_frsize // (1 << 20)
            max_disk_mb = max(75 * avail // 100, avail - 2000) or 10000
        if max_open_files is None:
            max_open_files = resource.getrlimit(resource.RLIMIT_NOFILE)[0] // 2
        nb_passes = (volume // 4) // max_disk_mb + 1
        in_parallel = max(nb_cores, 1)
        while True:
            per_pass = volume_minim // nb_passes
            nb_partitions = (per_pass * in_parallel) // max_memory_mb + 1
            if nb_partitions >= max_open_files and in_parallel > 1:
                in_parallel //= 2
            elif nb_partitions >= max_open_files:
                nb_passes += 1
            else:
                break
        inc = (in_parallel - nb_partitions % in_parallel) % in_parallel        # rounded up to a multiple of the partitions in parallel (:422-425)
        if max_open_files - nb_partitions > inc:
            nb_partitions += inc
        return nb_passes, nb_partitions

    def dsk(self, fasta_path, k, m, abundance_min=2, nb_cores=1, max_memory_mb=5000, minimizer_type=0, repartition_type=0):
        """Runs SortingCountAlgorithm on a FASTA/FASTQ file.  Returns dict(parts, repart, nb_partitions, ...)."""
        with tempfile.TemporaryDirectory() as tmp:
            cwd = os.getcwd()
            os.chdir(tmp)          # the reference drops temp files in the current directory
            try:
                h = self.L.ref_dsk_run(os.path.abspath(fasta_path).encode() if not os.path.isabs(fasta_path) else fasta_path.encode(),
                                       k, m, abundance_min, nb_cores, max_memory_mb, os.path.join(tmp, "out").encode(),
                                       minimizer_type, repartition_type)
            finally:
                os.chdir(cwd)
            if not h:
                raise RuntimeError("reference DSK failed: %s" % self.L.ref_last_error().decode())
            res = {"nb_partitions": self.L.ref_dsk_nb_partitions(h), "nb_passes": self.L.ref_dsk_nb_passes(h),
                   "seconds": self.L.ref_dsk_seconds(h), "kmers_nb_valid": self.L.ref_dsk_kmers_nb_valid(h),
                   "kmers_nb_invalid": self.L.ref_dsk_kmers_nb_invalid(h), "nb_distinct": self.L.ref_dsk_nb_distinct(h),
                   "fill_partitions_s": self.L.ref_dsk_fill_partitions_seconds(h),
                   "fill_solid_s": self.L.ref_dsk_fill_solid_seconds(h),
                   "info_xml": self.L.ref_dsk_info_xml(h).decode(), "parts": {}}
            rp = np.zeros(4 ** m, np.uint16)
            self.L.ref_dsk_get_repart(h, rp)
            res["repart"] = rp
            for key in range(res["nb_partitions"] * res["nb_passes"]):
                n = self.L.ref_dsk_part_size(h, key)
                lo, hi, cn = np.zeros(n, np.uint64), np.zeros(n, np.uint64), np.zeros(n, np.int32)
                if n:
                    self.L.ref_dsk_get_part(h, key, _opt(lo), _opt(hi), _opt(cn))
                res["parts"][key] = (lo, hi, cn)
            self.L.ref_dsk_free(h)
            return res

    def histogram(self, abundances, histo_max=10000, min_auto_threshold=3):
        a = np.ascontiguousarray(abundances, np.int32)
        table = np.zeros(histo_max + 1, np.uint64)
        c, n, p = C.c_uint32(), C.c_uint64(), C.c_uint32()
        self.L.ref_histogram(a, len(a), histo_max, min_auto_threshold, table, C.byref(c), C.byref(n), C.byref(p))
        return table, c.value, n.value, p.value

    def kmers(self, seq, k, m):
        s = seq.encode() if isinstance(seq, str) else bytes(seq)
        n = max(len(s) - k + 1, 0)
        lo, hi = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        mi, va, st = np.zeros(n, np.uint32), np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        if n:
            assert self.L.ref_kmers(s, len(s), k, m, lo, hi, mi, va, st) >= 0
        return lo, hi, mi, va, st

    def superkmers(self, seqs, k, m, repart, nb_partitions, nb_passes=1, pass_=0):
        with tempfile.TemporaryDirectory() as tmp:
            return _superkmers(self.L.ref_superkmers, self.L.ref_free, seqs, k, m, repart, nb_partitions, nb_passes, pass_, tmp)

    def revcomp(self, lo, hi, k, words):
        a, b = C.c_uint64(), C.c_uint64()
        self.L.ref_revcomp(lo, hi, k, words, C.byref(a), C.byref(b))
        return a.value, b.value

    def hash1(self, lo, hi, words, seed):
        return self.L.ref_hash1(lo, hi, words, seed)

    def simplehash16(self, lo, hi, words, shift):
        return self.L.ref_simplehash16(lo, hi, words, shift)

    def nbits_per_kmer(self, k):
        return self.L.ref_nbits_per_kmer(k)

    def bloom(self, kind, bit_size, nb_hash, k, words, lo, hi=None):
        return _bloom(self.L.ref_bloom, kind, bit_size, nb_hash, k, words, lo, hi)
