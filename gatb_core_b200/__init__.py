"""gatb_core_b200 -- B200-native k-mer counting behind GATB-core's DSK interface.

This Python module is only a thin ctypes caller of the C ABI in include/gatb_gpu.h (libgatb_b200.so, built in-tree
from gatb_core_b200/csrc/*.cu for sm_100a by build.py).  There is NO CPU fallback: importing works without a GPU
(so the symbol table can be checked), but every compute call needs a CUDA device and raises GatbGpuError otherwise.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgatb_b200.so")

NSTATS = 16
STAT_NAMES = ["kmers_nb_valid", "kmers_nb_invalid", "kmers_nb_distinct", "kmers_nb_solid", "records", "sequences",
              "nucleotides", "bins", "overflow_bins", "retries", "record_bytes"]
BLOOM_KINDS = {"basic": 0, "cache": 1, "neighbor": 2}

# every symbol include/gatb_gpu.h declares (tests check the library exports exactly these)
EXPORTS = ["gatb_gpu_create", "gatb_gpu_destroy", "gatb_gpu_last_error", "gatb_gpu_stream", "gatb_gpu_kernel_launches",
           "gatb_gpu_sm_count", "gatb_gpu_count", "gatb_gpu_count_dev", "gatb_gpu_result_free", "gatb_gpu_superkmers",
           "gatb_gpu_free_host", "gatb_gpu_bloom_params", "gatb_gpu_bloom_layout", "gatb_gpu_bloom", "gatb_gpu_bloom_dev",
           "gatb_gpu_histogram_cutoff", "gatb_gpu_malloc", "gatb_gpu_free", "gatb_gpu_memcpy_h2d", "gatb_gpu_memcpy_d2h",
           "gatb_gpu_synchronize", "gatb_gpu_synth_reads_dev", "gatb_gpu_pack_ascii", "gatb_gpu_plan",
           "gatb_gpu_partition_into", "gatb_gpu_partition_range_into", "gatb_gpu_count_bins", "gatb_gpu_reads_begin",
           "gatb_gpu_reads_push_ascii", "gatb_gpu_reads_count", "gatb_gpu_reads_push_text", "gatb_gpu_reads_info", "gatb_gpu_synth_zipf_dev",
           "gatb_gpu_count_bins_routed", "gatb_gpu_sort_routed", "gatb_gpu_repartition", "gatb_gpu_count_multi"]


class GatbGpuError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("kmer_size", C.c_int32), ("minimizer_size", C.c_int32), ("nb_partitions", C.c_int32),
                ("nb_passes", C.c_int32), ("abundance_min", C.c_int32), ("abundance_max", C.c_int32),
                ("histo_max", C.c_int32), ("minimizer_type", C.c_int32), ("emit_all", C.c_int32),
                ("read_len", C.c_int32), ("table_log2", C.c_int32), ("path_flags", C.c_int32),
                ("k3_dir_rounds", C.c_int32), ("bin_load_pct", C.c_int32), ("fine_bits", C.c_int32), ("bin_target_pct", C.c_int32)]


# gatb_gpu_params.path_flags (include/gatb_gpu.h): selectors of the alternate code paths, 0 = the product path
PATH_K1_GENERAL, PATH_K2B_CTA128, PATH_K2B_CTA256, PATH_K2B_LANE = 1, 2, 4, 6
PATH_K2B_W2_WARP, PATH_NO_TIER2, PATH_K3_NO_POOL, PATH_CANONICAL, PATH_NO_DEDUP, PATH_FUSED, PATH_K1_STAGING = 8, 16, 32, 64, 128, 256, 512
PATH_K2A_SMALL_STAGE, PATH_K2A_PRESPLIT = 1024, 2048


class Result(C.Structure):
    _fields_ = [("n_keys", C.c_uint64), ("n_items", C.c_uint64), ("part_offsets", C.c_void_p), ("kmers_lo", C.c_void_p),
                ("kmers_hi", C.c_void_p), ("counts", C.c_void_p), ("histogram", C.c_void_p),
                ("stats", C.c_uint64 * NSTATS), ("seconds", C.c_double * 8), ("kernel_seconds", C.c_double * 8), ("on_device", C.c_int32), ("pad", C.c_int32),
                ("owner", C.c_void_p)]


class Geometry(C.Structure):
    _fields_ = [("total_kmers", C.c_uint64), ("nb1", C.c_uint32), ("cap", C.c_uint32), ("fine_bits", C.c_int32),
                ("table_log2", C.c_int32), ("m_device", C.c_int32), ("w", C.c_int32), ("maxlen", C.c_int32),
                ("words", C.c_int32), ("n_ranks", C.c_uint32), ("bins_per_rank", C.c_uint32), ("record_bytes", C.c_uint32),
                ("coarse_blk", C.c_uint32)]


def load_library():
    if not os.path.exists(LIB_PATH):
        raise GatbGpuError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
                           "There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    VP, U64, I32 = C.c_void_p, C.c_uint64, C.c_int
    L.gatb_gpu_create.restype = VP
    L.gatb_gpu_create.argtypes = [I32]
    L.gatb_gpu_destroy.argtypes = [VP]
    L.gatb_gpu_last_error.restype = C.c_char_p
    L.gatb_gpu_last_error.argtypes = [VP]
    L.gatb_gpu_stream.restype = VP
    L.gatb_gpu_stream.argtypes = [VP]
    L.gatb_gpu_kernel_launches.restype = U64
    L.gatb_gpu_kernel_launches.argtypes = [VP]
    L.gatb_gpu_sm_count.argtypes = [VP]
    for f in ("gatb_gpu_count", "gatb_gpu_count_dev"):
        getattr(L, f).argtypes = [VP, C.POINTER(Params), VP, VP, VP, VP, U64, VP, C.POINTER(Result)]
    L.gatb_gpu_result_free.argtypes = [VP, C.POINTER(Result)]
    L.gatb_gpu_reads_begin.argtypes = [VP, U64]
    L.gatb_gpu_reads_push_ascii.argtypes = [VP, C.c_char_p, VP, U64]
    L.gatb_gpu_reads_count.argtypes = [VP, C.POINTER(Params), VP, VP, C.POINTER(Result)]
    L.gatb_gpu_reads_push_text.argtypes = [VP, C.c_char_p, U64, I32]
    L.gatb_gpu_reads_info.argtypes = [VP, VP]
    L.gatb_gpu_superkmers.argtypes = [VP, C.POINTER(Params), VP, VP, VP, U64, VP, C.POINTER(VP), VP, VP]
    L.gatb_gpu_free_host.argtypes = [VP]
    L.gatb_gpu_bloom_params.argtypes = [I32, U64, C.POINTER(U64), C.POINTER(C.c_int32)]
    L.gatb_gpu_bloom_layout.argtypes = [I32, U64, C.POINTER(U64), C.POINTER(U64)]
    L.gatb_gpu_bloom.argtypes = [VP, I32, U64, I32, I32, VP, VP, U64, VP]
    L.gatb_gpu_bloom_dev.argtypes = [VP, I32, U64, I32, I32, VP, VP, U64, VP]
    L.gatb_gpu_histogram_cutoff.argtypes = [VP, I32, I32, C.POINTER(C.c_uint32), C.POINTER(U64), C.POINTER(C.c_uint32)]
    L.gatb_gpu_malloc.restype = VP
    L.gatb_gpu_malloc.argtypes = [VP, U64]
    L.gatb_gpu_free.argtypes = [VP, VP]
    L.gatb_gpu_memcpy_h2d.argtypes = [VP, VP, VP, U64]
    L.gatb_gpu_memcpy_d2h.argtypes = [VP, VP, VP, U64]
    L.gatb_gpu_synchronize.argtypes = [VP]
    L.gatb_gpu_synth_reads_dev.argtypes = [VP, U64, U64, U64, U64, I32, VP]
    L.gatb_gpu_synth_zipf_dev.argtypes = [VP, U64, U64, VP, VP, U64, U64, I32, VP]
    L.gatb_gpu_pack_ascii.argtypes = [VP, C.c_char_p, U64, VP, VP, C.POINTER(U64)]
    L.gatb_gpu_plan.argtypes = [VP, C.POINTER(Params), U64, U64, I32, C.POINTER(Geometry)]
    L.gatb_gpu_partition_into.argtypes = [VP, C.POINTER(Params), C.POINTER(Geometry), VP, VP, U64, VP, VP, VP, VP]
    L.gatb_gpu_partition_range_into.argtypes = [VP, C.POINTER(Params), C.POINTER(Geometry), VP, VP, U64, U64, VP, VP, VP, VP]
    L.gatb_gpu_count_bins.argtypes = [VP, C.POINTER(Params), C.POINTER(Geometry), I32, C.POINTER(VP), C.POINTER(VP),
                                      C.c_uint32, VP, U64, C.POINTER(Result)]
    L.gatb_gpu_count_bins_routed.argtypes = [VP, C.POINTER(Params), C.POINTER(Geometry), I32, C.POINTER(VP), C.POINTER(VP),
                                             C.c_uint32, VP, U64, I32, VP, C.POINTER(VP), C.POINTER(Result)]
    L.gatb_gpu_count_multi.argtypes = [C.POINTER(VP), I32, C.POINTER(Params), VP, VP, VP, U64, VP, C.POINTER(Result)]
    L.gatb_gpu_repartition.argtypes = [VP, C.POINTER(Params), VP, VP, U64, VP, U64, VP, VP]
    L.gatb_gpu_sort_routed.argtypes = [VP, C.POINTER(Params), VP, VP, VP, VP, U64, I32, I32, C.POINTER(Result)]
    return L


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


class GatbGpu:
    """One context per GPU (one process per GPU for multi-GPU runs)."""

    def __init__(self, device=0):
        self.L = load_library()
        self.ctx = self.L.gatb_gpu_create(device)
        if not self.ctx:
            raise GatbGpuError(self.L.gatb_gpu_last_error(None).decode())
        self.device = device

    def close(self):
        if getattr(self, "ctx", None):
            self.L.gatb_gpu_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise GatbGpuError(self.L.gatb_gpu_last_error(self.ctx).decode())

    @property
    def kernel_launches(self):
        return self.L.gatb_gpu_kernel_launches(self.ctx)

    @property
    def stream(self):
        return self.L.gatb_gpu_stream(self.ctx)

    @property
    def sm_count(self):
        return self.L.gatb_gpu_sm_count(self.ctx)

    # ---- DSK ---------------------------------------------------------------------------------------------------
    @staticmethod
    def make_params(k, m, nb_partitions=1, nb_passes=1, abundance_min=2, abundance_max=2**31 - 1, histo_max=10000,
                    emit_all=False, read_len=0, table_log2=0, path_flags=0, k3_dir_rounds=0, bin_load_pct=0, fine_bits=0, bin_target_pct=0):
        p = Params()
        p.path_flags, p.k3_dir_rounds, p.bin_load_pct, p.fine_bits = path_flags, k3_dir_rounds, bin_load_pct, fine_bits
        p.bin_target_pct = bin_target_pct
        p.kmer_size, p.minimizer_size, p.nb_partitions, p.nb_passes = k, m, nb_partitions, nb_passes
        p.abundance_min, p.abundance_max, p.histo_max, p.minimizer_type = abundance_min, abundance_max, histo_max, 0
        p.emit_all, p.read_len, p.table_log2 = int(emit_all), read_len, table_log2
        return p

    def count(self, packed, offsets, n_reads, params, repart=None, n_mask=None):
        """Host buffers in, host arrays out.  Returns dict(parts={key:(lo,hi,counts)}, histogram, stats, seconds)."""
        res = Result()
        rp = None if repart is None else np.ascontiguousarray(repart, np.uint16)
        self._check(self.L.gatb_gpu_count(self.ctx, C.byref(params), _ptr(rp), None, _ptr(packed), _ptr(offsets), n_reads,
                                          _ptr(n_mask), C.byref(res)))
        try:
            return self._unpack_host(res, params)
        finally:
            self.L.gatb_gpu_result_free(self.ctx, C.byref(res))

    def count_pushed(self, batches, params, repart=None):
        """Streaming input: `batches` is an iterable of lists of ASCII sequences (bytes); each batch is pushed with
        gatb_gpu_reads_push_ascii and packed on the device; one count at the end (host arrays out, like count())."""
        self._check(self.L.gatb_gpu_reads_begin(self.ctx, 0))
        for seqs in batches:
            blob = b"".join(seqs)
            offs = np.zeros(len(seqs) + 1, np.uint64)
            offs[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
            self._check(self.L.gatb_gpu_reads_push_ascii(self.ctx, blob, _ptr(offs), len(seqs)))
        res = Result()
        rp = None if repart is None else np.ascontiguousarray(repart, np.uint16)
        self._check(self.L.gatb_gpu_reads_count(self.ctx, C.byref(params), _ptr(rp), None, C.byref(res)))
        try:
            return self._unpack_host(res, params)
        finally:
            self.L.gatb_gpu_result_free(self.ctx, C.byref(res))

    def count_text(self, batches, fmt, params, repart=None):
        """FASTA (fmt=0) / FASTQ (fmt=1) text batches, cut at record boundaries, parsed and packed on the device; returns
        (result like count(), info = [sequences, nucleotides, shortest, longest, sum of squares (float), invalid])."""
        self._check(self.L.gatb_gpu_reads_begin(self.ctx, 0))
        for text in batches:
            self._check(self.L.gatb_gpu_reads_push_text(self.ctx, text, len(text), fmt))
        info = np.zeros(6, np.uint64)
        self._check(self.L.gatb_gpu_reads_info(self.ctx, _ptr(info)))
        res = Result()
        rp = None if repart is None else np.ascontiguousarray(repart, np.uint16)
        self._check(self.L.gatb_gpu_reads_count(self.ctx, C.byref(params), _ptr(rp), None, C.byref(res)))
        try:
            out = self._unpack_host(res, params)
        finally:
            self.L.gatb_gpu_result_free(self.ctx, C.byref(res))
        return out, [int(info[0]), int(info[1]), int(info[2]), int(info[3]), float(info[4:5].view(np.float64)[0]), int(info[5])]

    def count_dev(self, d_packed, d_offsets, n_reads, params, repart=None, d_n_mask=None):
        """Device pointers (ints) in; returns the raw Result (device arrays) -- free with result_free()."""
        res = Result()
        rp = None if repart is None else np.ascontiguousarray(repart, np.uint16)
        self._check(self.L.gatb_gpu_count_dev(self.ctx, C.byref(params), _ptr(rp), None, _ptr(d_packed), _ptr(d_offsets),
                                              n_reads, _ptr(d_n_mask), C.byref(res)))
        return res

    def result_free(self, res):
        self.L.gatb_gpu_result_free(self.ctx, C.byref(res))

    def result_to_host(self, res, params):
        """Copies a device Result to numpy arrays (same layout as count())."""
        W = 1 if params.kmer_size < 32 else 2
        n, nk = int(res.n_items), int(res.n_keys)
        offs = np.zeros(nk + 1, np.uint64)
        lo, hi, cn = np.zeros(n, np.uint64), np.zeros(n if W == 2 else 0, np.uint64), np.zeros(n, np.int32)
        hist = np.zeros(params.histo_max + 1, np.uint64)
        self.d2h(offs, res.part_offsets)
        self.d2h(hist, res.histogram)
        if n:
            self.d2h(lo, res.kmers_lo)
            self.d2h(cn, res.counts)
            if W == 2:
                self.d2h(hi, res.kmers_hi)
        return self._assemble(res, offs, lo, hi if W == 2 else np.zeros(n, np.uint64), cn, hist)

    def _unpack_host(self, res, params):
        W = 1 if params.kmer_size < 32 else 2
        n, nk = int(res.n_items), int(res.n_keys)

        def arr(ptr, count, dtype):
            if count == 0 or not ptr:
                return np.zeros(0, dtype)
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), (count,)).copy()
        offs = arr(res.part_offsets, nk + 1, np.uint64)
        lo, cn = arr(res.kmers_lo, n, np.uint64), arr(res.counts, n, np.int32)
        hi = arr(res.kmers_hi, n, np.uint64) if W == 2 else np.zeros(n, np.uint64)
        hist = arr(res.histogram, params.histo_max + 1, np.uint64)
        return self._assemble(res, offs, lo, hi, cn, hist)

    @staticmethod
    def _assemble(res, offs, lo, hi, cn, hist):
        parts = {}
        for key in range(int(res.n_keys)):
            a, b = int(offs[key]), int(offs[key + 1])
            parts[key] = (lo[a:b], hi[a:b], cn[a:b])
        stats = {name: int(res.stats[i]) for i, name in enumerate(STAT_NAMES)}
        return {"parts": parts, "part_offsets": offs, "histogram": hist, "stats": stats,
                "seconds": [float(x) for x in res.seconds], "kernel_seconds": [float(x) for x in res.kernel_seconds], "n_items": int(res.n_items)}

    # ---- staged path (multi-GPU) --------------------------------------------------------------------------------
    def plan(self, params, total_kmers, n_reads, n_ranks):
        g = Geometry()
        self._check(self.L.gatb_gpu_plan(self.ctx, C.byref(params), total_kmers, n_reads, n_ranks, C.byref(g)))
        return g

    def partition_into(self, params, geom, d_packed, d_offsets, n_reads, d_bins, d_cursors, d_n_mask=None, first_read=0):
        """k1 over the reads [first_read, first_read + n_reads) into caller-provided device buffers (ints).
        Returns [valid, invalid, stored, dropped]."""
        st = np.zeros(4, np.uint64)
        self._check(self.L.gatb_gpu_partition_range_into(self.ctx, C.byref(params), C.byref(geom), _ptr(d_packed), _ptr(d_offsets),
                                                         first_read, n_reads, _ptr(d_n_mask), _ptr(d_bins), _ptr(d_cursors), _ptr(st)))
        return [int(x) for x in st]

    def count_bins(self, params, geom, src_bins, src_cursors, nb1_local, kmers_bound, repart=None):
        """Counts nb1_local coarse bins gathered from len(src_bins) sources (device pointers).  Returns a device Result."""
        n = len(src_bins)
        a = (C.c_void_p * n)(*src_bins)
        b = (C.c_void_p * n)(*src_cursors)
        res = Result()
        rp = None if repart is None else np.ascontiguousarray(repart, np.uint16)
        self._check(self.L.gatb_gpu_count_bins(self.ctx, C.byref(params), C.byref(geom), n, a, b, nb1_local,
                                               _ptr(rp), kmers_bound, C.byref(res)))
        return res

    def count_bins_routed(self, params, geom, src_bins, src_cursors, nb1_local, kmers_bound, n_ranks, repart=None):
        """Like count_bins, but instead of sorting, the emitted k-mers are grouped by the rank that owns their partition key.
        Returns (device Result with kmers_lo / kmers_hi / counts in one region per destination rank, device pointer of the 16-bit keys,
        items per destination, items a region can hold)."""
        n = len(src_bins)
        a = (C.c_void_p * n)(*src_bins)
        b = (C.c_void_p * n)(*src_cursors)
        res = Result()
        rp = None if repart is None else np.ascontiguousarray(repart, np.uint16)
        counts = np.zeros(n_ranks + 1, np.uint64)
        keys = C.c_void_p()
        self._check(self.L.gatb_gpu_count_bins_routed(self.ctx, C.byref(params), C.byref(geom), n, a, b, nb1_local, _ptr(rp), kmers_bound,
                                                      n_ranks, _ptr(counts), C.byref(keys), C.byref(res)))
        return res, keys.value, [int(x) for x in counts[:n_ranks]], int(counts[n_ranks])

    def sort_routed(self, params, d_lo, d_hi, d_counts, d_keys, n_items, n_ranks, rank):
        """Partition key + ascending order of routed items (device pointers).  Returns a device Result."""
        res = Result()
        self._check(self.L.gatb_gpu_sort_routed(self.ctx, C.byref(params), _ptr(d_lo), _ptr(d_hi), _ptr(d_counts), _ptr(d_keys), n_items, n_ranks, rank, C.byref(res)))
        return res

    def count_multi(self, others, packed, offsets, n_reads, params, repart=None, n_mask=None):
        """One process, several devices (this context first, then `others`: GatbGpu objects of other devices): gatb_gpu_count_multi.
        Host arrays in, the same dict as count() out."""
        ctxs = (C.c_void_p * (1 + len(others)))(self.ctx, *[o.ctx for o in others])
        res = Result()
        rp = None if repart is None else np.ascontiguousarray(repart, np.uint16)
        self._check(self.L.gatb_gpu_count_multi(ctxs, 1 + len(others), C.byref(params), _ptr(rp), _ptr(packed),
                                                _ptr(None if offsets is None else np.ascontiguousarray(offsets, np.uint64)), n_reads, _ptr(n_mask), C.byref(res)))
        return self._unpack_host(res, params)

    def repartition(self, packed, offsets, n_reads, params, nb_seqs_to_see, n_mask=None):
        """The Repartitor table (u16[4^m]) of these reads: sampling pass on the device, distribution on the host.
        Returns (table, [reads sampled, super-k-mers seen, kx-mers charged])."""
        table = np.zeros(4 ** params.minimizer_size, np.uint16)
        info = np.zeros(3, np.uint64)
        self._check(self.L.gatb_gpu_repartition(self.ctx, C.byref(params), _ptr(packed), _ptr(None if offsets is None else np.ascontiguousarray(offsets, np.uint64)),
                                                n_reads, _ptr(n_mask), nb_seqs_to_see, _ptr(table), _ptr(info)))
        return table, [int(x) for x in info]

    # ---- GATB-exact super-k-mers ---------------------------------------------------------------------------------
    def superkmers(self, packed, offsets, n_reads, params, repart=None, n_mask=None):
        nk = params.nb_partitions * params.nb_passes
        ptrs = (C.c_void_p * nk)()
        sizes = np.zeros(nk, np.uint64)
        stats = np.zeros(4, np.uint64)
        rp = None if repart is None else np.ascontiguousarray(repart, np.uint16)
        self._check(self.L.gatb_gpu_superkmers(self.ctx, C.byref(params), _ptr(rp), _ptr(packed), _ptr(offsets), n_reads,
                                               _ptr(n_mask), ptrs, _ptr(sizes), _ptr(stats)))
        out = []
        for key in range(nk):
            out.append(C.string_at(ptrs[key], int(sizes[key])))
            self.L.gatb_gpu_free_host(ptrs[key])
        return out, stats

    # ---- Bloom ---------------------------------------------------------------------------------------------------
    def bloom_params(self, k, nb_solid):
        s, h = C.c_uint64(), C.c_int32()
        if self.L.gatb_gpu_bloom_params(k, nb_solid, C.byref(s), C.byref(h)):
            raise GatbGpuError("bad kmer size")
        return s.value, h.value

    def bloom_layout(self, kind, bloom_size):
        nb, bits = C.c_uint64(), C.c_uint64()
        if self.L.gatb_gpu_bloom_layout(BLOOM_KINDS[kind], bloom_size, C.byref(nb), C.byref(bits)):
            raise GatbGpuError("bad Bloom kind")
        return nb.value, bits.value

    def bloom(self, kind, bloom_size, nb_hash, k, lo, hi=None):
        nbytes, bits = self.bloom_layout(kind, bloom_size)
        out = np.zeros(nbytes, np.uint8)
        lo = np.ascontiguousarray(lo, np.uint64)
        hi = None if hi is None else np.ascontiguousarray(hi, np.uint64)
        self._check(self.L.gatb_gpu_bloom(self.ctx, BLOOM_KINDS[kind], bloom_size, nb_hash, k, _ptr(lo), _ptr(hi), len(lo), _ptr(out)))
        return out, bits

    def histogram_cutoff(self, histogram, min_auto_threshold=3):
        h = np.ascontiguousarray(histogram, np.uint64)
        c, n, p = C.c_uint32(), C.c_uint64(), C.c_uint32()
        if self.L.gatb_gpu_histogram_cutoff(_ptr(h), len(h) - 1, min_auto_threshold, C.byref(c), C.byref(n), C.byref(p)):
            raise GatbGpuError("bad histogram")
        return c.value, n.value, p.value

    # ---- device utilities ----------------------------------------------------------------------------------------
    def malloc(self, nbytes):
        p = self.L.gatb_gpu_malloc(self.ctx, nbytes)
        if not p:
            raise GatbGpuError(self.L.gatb_gpu_last_error(self.ctx).decode())
        return p

    def free(self, p):
        self.L.gatb_gpu_free(self.ctx, p)

    def h2d(self, dptr, arr):
        self._check(self.L.gatb_gpu_memcpy_h2d(self.ctx, _ptr(dptr), _ptr(arr), arr.nbytes))

    def d2h(self, arr, dptr):
        self._check(self.L.gatb_gpu_memcpy_d2h(self.ctx, _ptr(arr), _ptr(dptr), arr.nbytes))

    def synchronize(self):
        self._check(self.L.gatb_gpu_synchronize(self.ctx))

    def synth_zipf_dev(self, seed, cdf, genome_off, first_read, n_reads, L, d_packed):
        self._check(self.L.gatb_gpu_synth_zipf_dev(self.ctx, seed, len(cdf), _ptr(cdf), _ptr(genome_off), first_read, n_reads, L, C.c_void_p(d_packed)))

    def synth_reads_dev(self, seed, genome_len, first_read, n_reads, L, d_packed):
        self._check(self.L.gatb_gpu_synth_reads_dev(self.ctx, seed, genome_len, first_read, n_reads, L, _ptr(d_packed)))

    def pack_ascii(self, ascii_bytes):
        n = len(ascii_bytes)
        packed = np.zeros((n + 3) // 4 + 16, np.uint8)
        mask = np.zeros((n + 31) // 32 + 4, np.uint32)
        bad = C.c_uint64()
        self._check(self.L.gatb_gpu_pack_ascii(self.ctx, ascii_bytes, n, _ptr(packed), _ptr(mask), C.byref(bad)))
        return packed, mask, bad.value
