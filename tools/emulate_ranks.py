"""One GPU plays all ranks of a distributed count (staged C ABI): every rank's reads are partitioned into all regions,
then the bins of ONE owner rank are counted from the pieces of all sources.  Prints the overflow statistics next to the
one-shot count of the same reads.  usage: python tools/emulate_ranks.py [reads_total] [world]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gatb_core_b200

n_global = int(sys.argv[1]) if len(sys.argv) > 1 else 16_000_000
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
L, K, M = 150, 31, 10
gpu = gatb_core_b200.GatbGpu(0)
dev = torch.device("cuda", 0)
n = n_global // world
genome = n_global * L // 30
params = gpu.make_params(K, M, abundance_min=2, read_len=L)
total_kmers = n_global * (L - K + 1)
geom = gpu.plan(params, total_kmers, n_global, world)
print("geometry: nb1=%d cap=%d bins_per_rank=%d table_log2=%d m=%d w=%d" % (geom.nb1, geom.cap, geom.bins_per_rank, geom.table_log2, geom.m_device, geom.w))
if len(sys.argv) > 4:
    geom.cap = int(sys.argv[4])
bpr, cap, rb = geom.bins_per_rank, geom.cap, geom.record_bytes
pieces, curs = [], []
owner = 0
for s in range(world):
    reads = torch.zeros((n * L + 3) // 4 + 64, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    gpu.synth_reads_dev(42, genome, s * n, n, L, reads.data_ptr())
    bins = torch.empty(geom.nb1 * cap * rb, dtype=torch.uint8, device=dev)
    cursors = torch.zeros(geom.nb1, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()          # torch's fills run on torch's stream, the library on its own
    st = gpu.partition_into(params, geom, reads.data_ptr(), None, n, bins.data_ptr(), cursors.data_ptr())
    assert st[3] == 0, "bin overflow in partition: %s (raise the cap: 4th argument)" % st
    pieces.append(bins.view(world, -1)[owner].clone())
    curs.append(cursors.view(world, -1)[owner].clone())
    del bins, cursors, reads
gathered = int(sum(int(c.clamp(max=cap).sum().item()) for c in curs))
res = gpu.count_bins(params, geom, [p.data_ptr() for p in pieces], [c.data_ptr() for c in curs], bpr, gathered * geom.maxlen)
cur_all = torch.stack(curs).to(torch.int64)
print("cursors per source piece: max %d mean %.1f (cap %d)" % (int(cur_all.max()), float(cur_all.float().mean()), cap))
print("owner %d of %d ranks: records %d distinct %d solid %d bins %d overflow_bins %d overflow_kmers %d" % (
    owner, world, int(res.stats[4]), int(res.stats[2]), int(res.stats[3]), int(res.stats[7]), int(res.stats[8]), int(res.stats[11])))
gpu.result_free(res)
if len(sys.argv) > 3 and sys.argv[3] == 'noshot':
    sys.exit(0)
allr = torch.zeros((n_global * L + 3) // 4 + 64, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
gpu.synth_reads_dev(42, genome, 0, n_global, L, allr.data_ptr())
one = gpu.count_dev(allr.data_ptr(), None, n_global, params)
print("one-shot: records %d distinct %d solid %d bins %d overflow_bins %d" % (int(one.stats[4]), int(one.stats[2]), int(one.stats[3]), int(one.stats[7]), int(one.stats[8])))
