"""Times gatb_gpu_count_dev on a synthetic slice of any BASELINE configuration (k, read length, reads):
python tools/bench_config.py K L N [steps].  Prints one JSON line with the step time and the kernel split."""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gatb_core_b200

K, L, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
flags = int(sys.argv[5]) if len(sys.argv) > 5 else 0
tlog = int(sys.argv[6]) if len(sys.argv) > 6 else 0
load = int(sys.argv[7]) if len(sys.argv) > 7 else 0
gpu = gatb_core_b200.GatbGpu(0)
nbytes = (N * L + 3) // 4
d_reads = gpu.malloc(nbytes + 64)
gpu.synth_reads_dev(44, N * L // 30, 0, N, L, d_reads)
gpu.synchronize()
params = gpu.make_params(K, 10, abundance_min=2, read_len=L, path_flags=flags, table_log2=tlog, bin_load_pct=load)
out = None
for i in range(steps + 1):
    t0 = time.time()
    res = gpu.count_dev(d_reads, None, N, params)
    dt = time.time() - t0
    out = {"flags": flags, "table_log2": tlog, "load": load, "k": K, "read_len": L, "reads": N, "ms_per_step": dt * 1e3, "distinct": int(res.stats[2]), "solid": int(res.stats[3]),
           "records": int(res.stats[4]), "overflow_bins": int(res.stats[8]),
           "kernel_ms": dict(zip(["k1", "k2a", "k2b", "k3", "overflow_tiers"], [round(float(x) * 1e3, 2) for x in res.kernel_seconds][:5])),
           "distinct_per_s": int(res.stats[2]) / dt, "bases_per_s": N * L / dt}
    gpu.result_free(res)
print(json.dumps(out))
