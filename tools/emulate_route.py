"""One GPU: count a batch, route the emitted k-mers as for N ranks (gatb_gpu_count_bins_routed) and sort the region of ONE destination rank
(gatb_gpu_sort_routed): time and diagnostics of the routed sort next to the plain one.  usage: python tools/emulate_route.py [reads] [ranks] [partitions]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gatb_core_b200
from gatb_core_b200.multigpu import _as_tensor

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
nparts = int(sys.argv[3]) if len(sys.argv) > 3 else 2816
L, K, M = 150, 31, 10
gpu = gatb_core_b200.GatbGpu(0)
dev = torch.device("cuda", 0)
repart = (np.arange(4 ** M, dtype=np.uint64) * 2654435761 % nparts).astype(np.uint16)
if len(sys.argv) > 4 and sys.argv[4] == "ref":       # the reference's own Repartitor table for an 8x job (its largest minimizers get partitions of their own)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    import bench
    class A: pass
    a = A(); a.nb_partitions = nparts; a.repart_sample_reads = 2_000_000
    _, _, repart, src = bench.reference_configuration(gpu, n, a)
    print("repartition:", src)
params = gpu.make_params(K, M, nb_partitions=nparts, abundance_min=2, read_len=L)
reads = torch.zeros((n * L + 3) // 4 + 64, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
gpu.synth_reads_dev(42, n * L // 30, 0, n, L, reads.data_ptr())
gpu.synchronize()
geom = gpu.plan(params, n * (L - K + 1), n, 1)
bins = torch.empty(geom.nb1 * geom.cap * geom.record_bytes, dtype=torch.uint8, device=dev)
cursors = torch.zeros(geom.nb1, dtype=torch.int32, device=dev)
torch.cuda.synchronize()
st = gpu.partition_into(params, geom, reads.data_ptr(), None, n, bins.data_ptr(), cursors.data_ptr())
assert st[3] == 0
for rep in range(2):
    res = gpu.count_bins(params, geom, [bins.data_ptr()], [cursors.data_ptr()], geom.nb1, n * (L - K + 1), repart=repart)
    print("plain: items %d k3 %.1f ms diag %s" % (int(res.n_items), float(res.kernel_seconds[3]) * 1e3, [float(res.kernel_seconds[i]) for i in (5, 6, 7)]))
    gpu.result_free(res)
for rep in range(2):
    res1, d_keys, send, cap = gpu.count_bins_routed(params, geom, [bins.data_ptr()], [cursors.data_ptr()], geom.nb1, n * (L - K + 1), world, repart=repart)
    print("route kernel %.1f ms, items per destination %s (region of %d)" % (float(res1.kernel_seconds[3]) * 1e3, send, cap))
    if rep == 0:
        ky0 = _as_tensor(d_keys, cap * world * 2, torch.uint8, dev)[:send[0] * 2].clone().view(torch.int16).to(torch.int64)
        per_key = torch.bincount(ky0, minlength=nparts // world + 1)
        top = torch.sort(per_key, descending=True).values[:8].tolist()
        print("destination 0: largest keys hold %s items, mean %.0f" % (top, float(per_key.float().mean())))
    # destination 0 receives its region from every "source": the same region 'world' times stands in for the other sources' (distinct
    # values are not needed to time the sort; duplicates of a k-mer stay next to each other)
    lo = _as_tensor(res1.kmers_lo, cap * world, torch.int64, dev)[:send[0]].clone()
    cn = _as_tensor(res1.counts, cap * world, torch.int32, dev)[:send[0]].clone()
    ky = _as_tensor(d_keys, cap * world * 2, torch.uint8, dev)[:send[0] * 2].clone()
    for copies in (1, world):
        lo_c, cn_c, ky_c = lo.repeat(copies), cn.repeat(copies), ky.repeat(copies)
        torch.cuda.synchronize()
        res = gpu.sort_routed(params, lo_c.data_ptr(), None, cn_c.data_ptr(), ky_c.data_ptr(), send[0] * copies, world, 0)
        print("routed sort of %d items (%d copies of region 0): %.1f ms diag %s" % (send[0] * copies, copies, float(res.kernel_seconds[3]) * 1e3, [float(res.kernel_seconds[i]) for i in (5, 6, 7)]))
