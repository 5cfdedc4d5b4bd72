#!/usr/bin/env python
"""BASELINE.json config 5 on ONE GPU at a chosen size: metagenome-like Zipf-skew reads (10^4 species of 10^5..10^6 nt, exponent
1.1, seed 45), k = 31 DSK with the reference-style partitioning, then the Bloom filter (kind neighbor) of the solid
k-mers.  Prints one JSON line: step time, kernel split, overflow statistics (how the skew lands in the tiers), Bloom time.
   python tools/bench_zipf.py [reads] [species]"""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import gatb_core_b200
import oracle_lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
K, M, L = 31, 10, 150
orc = oracle_lib.Oracle()
cdf, off = orc.zipf_tables(45, ns, 1.1)
gpu = gatb_core_b200.GatbGpu(0)
nbytes = (n * L + 3) // 4
d = gpu.malloc(nbytes + 64)
gpu.synth_zipf_dev(45, cdf, off, 0, n, L, d)
nparts = 64
repart = ((np.arange(4 ** M, dtype=np.uint64) * np.uint64(2654435761) >> np.uint64(7)) % np.uint64(nparts)).astype(np.uint16)
params = gpu.make_params(K, M, nb_partitions=nparts, abundance_min=2, read_len=L)
out = None
for i in range(3):
    t0 = time.time()
    res = gpu.count_dev(d, None, n, params, repart=repart)
    dt = time.time() - t0
    solid = int(res.stats[3])
    size, nh = gpu.bloom_params(K, solid)
    nbytes_b, _bits = gpu.bloom_layout("neighbor", size)
    db = gpu.malloc((nbytes_b + 3) // 4 * 4 + 64)
    t1 = time.time()
    gpu._check(gpu.L.gatb_gpu_bloom_dev(gpu.ctx, 2, size, nh, K, res.kmers_lo, None, solid, db))
    gpu.synchronize()
    tb = time.time() - t1
    gpu.free(db)
    out = {"workload": "Zipf(1.1) over %d species (%d nt of genomes), %d reads x %d bp, k=%d" % (ns, int(off[-1]), n, L, K),
           "ms_per_step": dt * 1e3, "bloom_ms": tb * 1e3, "bloom_bytes": int(nbytes_b), "distinct": int(res.stats[2]), "solid": solid,
           "records": int(res.stats[4]), "unique_records": int(res.stats[13]), "bins": int(res.stats[7]),
           "overflow_bins_tier1": int(res.stats[8]), "bins_to_global_table": int(res.stats[12]), "kmers_in_global_table": int(res.stats[11]),
           "kernel_ms": dict(zip(["k1", "k2a", "k2b", "k3", "overflow_tiers"], [round(float(x) * 1e3, 2) for x in res.kernel_seconds][:5])),
           "distinct_per_s": int(res.stats[2]) / dt}
    gpu.result_free(res)
print(json.dumps(out))
