"""Times the partition kernel alone (gatb_gpu_partition_into) on synthetic reads; GATB_GPU_K1_DEBUG selects what is skipped."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gatb_core_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
L, K, M = 150, 31, 10
gpu = gatb_core_b200.GatbGpu(0)
nbytes = (n * L + 3) // 4
d_reads = gpu.malloc(nbytes + 64)
gpu.synth_reads_dev(42, n * L // 30, 0, n, L, d_reads)
params = gpu.make_params(K, M, abundance_min=2, read_len=L)
geom = gpu.plan(params, n * (L - K + 1), n, 1)
dev = torch.device("cuda", 0)
bins = torch.empty(geom.nb1 * geom.cap * geom.record_bytes, dtype=torch.uint8, device=dev)
cursors = torch.zeros(geom.nb1, dtype=torch.int32, device=dev)
for it in range(4):
    torch.cuda.synchronize(); t0 = time.time()
    st = gpu.partition_into(params, geom, d_reads, None, n, bins.data_ptr(), cursors.data_ptr())
    torch.cuda.synchronize(); dt = time.time() - t0
print("debug=%s reads=%d nb1=%d cap=%d: %.2f ms  stats=%s" % (os.environ.get("GATB_GPU_K1_DEBUG", "0"), n, geom.nb1, geom.cap, dt * 1e3, st))
