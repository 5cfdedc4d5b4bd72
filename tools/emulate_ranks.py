"""One GPU plays all ranks of a distributed count (staged C ABI): every rank's reads are partitioned into all regions
(in the pieces the distributed path would use), then the bins of ONE owner rank are counted from the pieces of all
sources.  Prints the time of every partition call and the kernel times of the owner's count next to the overflow
statistics -- what one rank of an N-GPU run spends, minus the exchange.
usage: python tools/emulate_ranks.py [reads_total] [world] [noshot|shot] [cap] [pieces]"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gatb_core_b200
from gatb_core_b200 import multigpu

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
npc = int(sys.argv[5]) if len(sys.argv) > 5 else multigpu.pieces_per_rank(world, n)
blk = geom.coarse_blk
if npc > 1:
    geom.cap = (int(geom.cap / npc * 1.15) + 64 + blk - 1) // blk * blk
if len(sys.argv) > 4 and int(sys.argv[4]) > 0:
    geom.cap = int(sys.argv[4])
print("geometry: nb1=%d cap=%d bins_per_rank=%d fine_bits=%d table_log2=%d m=%d w=%d pieces=%d" % (
    geom.nb1, geom.cap, geom.bins_per_rank, geom.fine_bits, geom.table_log2, geom.m_device, geom.w, npc))
bpr, cap, rb = geom.bins_per_rank, geom.cap, geom.record_bytes
firsts = multigpu.piece_bounds(n, npc)
pieces, curs = [], []
owner = 0
t_part = []
used_bytes = 0
for s in range(world):
    reads = torch.zeros((n * L + 3) // 4 + 64, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    gpu.synth_reads_dev(42, genome, s * n, n, L, reads.data_ptr())
    gpu.synchronize()
    for i in range(npc):
        bins = torch.empty(geom.nb1 * cap * rb, dtype=torch.uint8, device=dev)
        cursors = torch.zeros(geom.nb1, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()          # torch's fills run on torch's stream, the library on its own
        t0 = time.time()
        st = gpu.partition_into(params, geom, reads.data_ptr(), None, firsts[i + 1] - firsts[i], bins.data_ptr(), cursors.data_ptr(), first_read=firsts[i])
        t_part.append(time.time() - t0)
        assert st[3] == 0, "bin overflow in partition: %s (raise the cap: 4th argument)" % st
        rounds = (cursors.view(world, bpr).max(dim=1).values.clamp(max=cap).to(torch.int64) + blk - 1) // blk
        used_bytes += int((rounds * (bpr * blk * rb)).sum().item())
        pieces.append(bins.view(world, -1)[owner].clone())
        curs.append(cursors.view(world, -1)[owner].clone())
        del bins, cursors
    del reads
print("partition calls (ms): %s ; per rank %.1f ms" % (" ".join("%.1f" % (x * 1e3) for x in t_part[:2 * npc]), sum(t_part) / world * 1e3))
print("regions in use: %.2f GB per rank (records themselves: %.2f GB)" % (used_bytes / world / 1e9, sum(int(c.sum().item()) for c in curs) * rb / 1e9))
gathered = int(sum(int(c.clamp(max=cap).sum().item()) for c in curs))
for rep in range(2):
    res = gpu.count_bins(params, geom, [p.data_ptr() for p in pieces], [c.data_ptr() for c in curs], bpr, gathered * geom.maxlen)
    ks = [float(x) * 1e3 for x in res.kernel_seconds][:5]
    print("owner count, kernel ms: k2a %.1f k2b %.1f k3 %.1f overflow tiers %.1f ; stages %s" % (ks[1], ks[2], ks[3], ks[4], " ".join("%.1f" % (float(x) * 1e3) for x in res.seconds[:7])))
    if rep == 0:
        gpu.result_free(res)
cur_all = torch.stack(curs).to(torch.int64)
print("cursors per source piece: max %d mean %.1f (cap %d)" % (int(cur_all.max()), float(cur_all.float().mean()), cap))
print("owner %d of %d ranks: records %d unique %d distinct %d solid %d bins %d overflow_bins %d (leaving the warp tier %d, the 4096-slot tier %d) to_global %d overflow_kmers %d" % (
    owner, world, int(res.stats[4]), int(res.stats[13]), int(res.stats[2]), int(res.stats[3]), int(res.stats[7]), int(res.stats[8]), int(res.stats[14]), int(res.stats[15]), int(res.stats[12]), int(res.stats[11])))
gpu.result_free(res)
if len(sys.argv) > 3 and sys.argv[3] == 'noshot':
    sys.exit(0)
allr = torch.zeros((n_global * L + 3) // 4 + 64, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
gpu.synth_reads_dev(42, genome, 0, n_global, L, allr.data_ptr())
one = gpu.count_dev(allr.data_ptr(), None, n_global, params)
print("one-shot: records %d distinct %d solid %d bins %d overflow_bins %d" % (int(one.stats[4]), int(one.stats[2]), int(one.stats[3]), int(one.stats[7]), int(one.stats[8])))
