"""Distribution of the coarse-bin loads after the partition kernel (oriented vs canonical records), and the records of the
fullest bin: python tools/k1_bin_stats.py [reads] [path_flags]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gatb_core_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
L, K, M = 150, 31, 10
gpu = gatb_core_b200.GatbGpu(0)
nbytes = (n * L + 3) // 4
d_reads = gpu.malloc(nbytes + 64)
gpu.synth_reads_dev(42, n * L // 30, 0, n, L, d_reads)
params = gpu.make_params(K, M, abundance_min=2, read_len=L, path_flags=flags)
geom = gpu.plan(params, n * (L - K + 1), n, 1)
dev = torch.device("cuda", 0)
bins = torch.empty(geom.nb1 * geom.cap * geom.record_bytes, dtype=torch.uint8, device=dev)
cursors = torch.zeros(geom.nb1, dtype=torch.int32, device=dev)
st = gpu.partition_into(params, geom, d_reads, None, n, bins.data_ptr(), cursors.data_ptr())
c = cursors.cpu().numpy().astype(np.int64)
print("flags=%d reads=%d nb1=%d cap=%d fine_bits=%d stats=%s" % (flags, n, geom.nb1, geom.cap, geom.fine_bits, st))
print("cursor mean %.1f std %.1f max %d  p99.9 %d  over cap: %d" % (c.mean(), c.std(), c.max(), np.percentile(c, 99.9), (c > geom.cap).sum()))
top = np.argsort(c)[-5:][::-1]
print("top bins", top, c[top])
# records of the fullest bin (round-interleaved layout: block r of bin b at ((r*nb + b)*64)
b = int(top[0]); nb = geom.nb1; cnt = min(int(c[b]), geom.cap)
rec = bins.view(torch.int64).view(-1, 2)
idx = [((s // 64) * nb + b) * 64 + s % 64 for s in range(cnt)]
r = rec[torch.tensor(idx, device=dev)].cpu().numpy().view(np.uint64)
hi = r[:, 1]
ln = (hi >> np.uint64(44)) & np.uint64(31)
fine = hi >> np.uint64(49)
print("lengths histogram", np.bincount(ln.astype(np.int64), minlength=25))
fc = np.bincount(fine.astype(np.int64))
print("fine bins: max %d mean %.1f; top fine ids %s" % (fc.max(), fc.mean(), np.argsort(fc)[-3:]))
f0 = int(np.argmax(fc))
sel = r[fine == np.uint64(f0)][:12]
for lo_, hi_ in sel:
    nn = K + int((hi_ >> np.uint64(44)) & np.uint64(31)) - 1
    v = int(lo_) | ((int(hi_) & ((1 << 44) - 1)) << 64)
    print("".join("ACTG"[(v >> (2 * i)) & 3] for i in range(nn)))
