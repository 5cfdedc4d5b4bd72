"""Loader for the committed reference fixtures (tests/golden/dsk_*.npz, made by tests/golden/make_golden.py)."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "dsk_*.npz")))


class Fixture:
    def __init__(self, name, oracle):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.z = z
        (self.k, self.m, self.abundance_min, self.nb_partitions, self.nb_passes, self.n_reads, self.L, self.seed,
         self.genome_len) = (int(v) for v in z["params"])
        self.words = 1 if self.k < 32 else 2
        self.repart = z["repart"]
        # regenerate the reads exactly as make_golden.py did
        codes = oracle.synth_reads(self.seed, self.genome_len, 0, self.n_reads, self.L).reshape(self.n_reads, self.L)
        self.codes = codes
        self.invalid = np.zeros_like(codes)
        for i, j in z["n_positions"]:
            self.invalid[i, j] = 1
        self.seqs = [oracle.codes_to_ascii(r) for r in codes]
        for i, j in z["n_positions"]:
            self.seqs[i] = self.seqs[i][:j] + b"N" + self.seqs[i][j + 1:]
            self.codes[i, j] = 3          # the reference encodes a bad character as G (value 3)

    def part(self, key):
        a, b = int(self.z["part_offsets"][key]), int(self.z["part_offsets"][key + 1])
        hi = self.z["kmers_hi"][a:b] if self.words == 2 else np.zeros(b - a, np.uint64)
        return self.z["kmers_lo"][a:b], hi, self.z["counts"][a:b]

    def solid(self, key):
        lo, hi, cn = self.part(key)
        s = cn >= self.abundance_min
        return lo[s], hi[s], cn[s]
