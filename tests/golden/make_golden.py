#!/usr/bin/env python
"""Generates tests/golden/dsk_*.npz from the UNMODIFIED reference (oracle/_ref/libgatbref.so, built by oracle/Makefile
from /root/reference).  Run here (the container with /root/reference); the .npz files are committed so that the
parity tests do not need the reference at run time.

Each fixture holds, for one synthetic input (generator: oracle/kmer_oracle.c orc_synth_reads, seeds below):
  params           k, m, abundance_min, nb_partitions, nb_passes, n_reads, L, seed, genome_len
  repart           u16[4^m]   the reference's Repartitor table (an INPUT of the GPU path)
  part_offsets     u64[nb_parts*nb_passes+1]   into the arrays below (all distinct k-mers, ascending per partition)
  kmers_lo/hi, counts          what ICountProcessor::process received
  histogram        u64[10001]  from the reference Histogram class, cutoff / nbsolids from compute_threshold
  bloom_<kind>     the reference Bloom bytes for the solid k-mers (abundance >= abundance_min), sized by BloomAlgorithm's rule
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib  # noqa: E402

CASES = [
    # name,            k,  m, n_reads, L,  seed, cores, with_N
    ("dsk_k21_cfg1",   21, 8, 10000, 100, 1,    1,     False),   # BASELINE.json configs[0] shape (m=8 keeps the table small)
    ("dsk_k31_parts",  31, 8, 6000,  150, 42,   4,     True),    # several partitions, an N in some reads
    ("dsk_k63_w16",    63, 8, 2000,  250, 44,   2,     False),   # Kmer<64> two-word path
]


def main():
    orc, ref = oracle_lib.Oracle(), oracle_lib.Reference()
    assert ref.available, "build oracle/_ref first (make -C oracle)"
    for name, k, m, n, L, seed, cores, with_n in CASES:
        genome_len = n * L // 30
        codes = orc.synth_reads(seed, genome_len, 0, n, L).reshape(n, L)
        seqs = [orc.codes_to_ascii(r) for r in codes]
        n_pos = []
        if with_n:
            for i in range(0, n, 97):
                j = (i * 7) % L
                seqs[i] = seqs[i][:j] + b"N" + seqs[i][j + 1:]
                n_pos.append((i, j))
        with tempfile.TemporaryDirectory() as tmp:
            fa = os.path.join(tmp, "reads.fa")
            with open(fa, "wb") as f:
                for i, s in enumerate(seqs):
                    f.write(b">r%d\n" % i + s + b"\n")
            res = ref.dsk(fa, k, m, abundance_min=2, nb_cores=cores)
        nkeys = res["nb_partitions"] * res["nb_passes"]
        offs = np.zeros(nkeys + 1, np.uint64)
        for key in range(nkeys):
            offs[key + 1] = offs[key] + len(res["parts"][key][0])
        lo = np.concatenate([res["parts"][key][0] for key in range(nkeys)])
        hi = np.concatenate([res["parts"][key][1] for key in range(nkeys)])
        cn = np.concatenate([res["parts"][key][2] for key in range(nkeys)])
        table, cutoff, nbsolids, peak = ref.histogram(cn)
        solid = cn >= 2
        words = 1 if k < 32 else 2
        nb_solid = int(solid.sum())
        bits = ref.nbits_per_kmer(k)
        bloom_size = int(np.uint64(np.float32(nb_solid) * np.float32(bits)))
        nb_hash = int(np.floor(np.float32(0.7) * np.float32(bits)))
        out = dict(params=np.array([k, m, 2, res["nb_partitions"], res["nb_passes"], n, L, seed, genome_len], np.int64),
                   n_positions=np.array(n_pos, np.int64).reshape(-1, 2), repart=res["repart"], part_offsets=offs,
                   kmers_lo=lo, kmers_hi=hi if words == 2 else np.zeros(0, np.uint64), counts=cn,
                   histogram=table, cutoff=np.array([cutoff, nbsolids, peak], np.uint64),
                   stats=np.array([res["kmers_nb_valid"], res["kmers_nb_invalid"], res["nb_distinct"], nb_solid], np.uint64),
                   bloom_size=np.array([bloom_size, nb_hash], np.uint64))
        for kind in ("basic", "cache", "neighbor"):
            b, bitsize = ref.bloom(kind, bloom_size, nb_hash, k, words, lo[solid], hi[solid] if words == 2 else None)
            out["bloom_" + kind] = b
            out["bloom_" + kind + "_bitsize"] = np.array([bitsize], np.uint64)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "parts", res["nb_partitions"], "passes", res["nb_passes"], "distinct", len(lo), "solid", nb_solid,
              "bloom", bloom_size, nb_hash, "->", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
