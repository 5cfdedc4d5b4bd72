#!/usr/bin/env python
"""Hot regions of a kernel from `ncu --page source --csv`: instructions executed and stall samples per contiguous
address window.  usage: python tools/ncu_hot.py src.csv [window]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
win = int(sys.argv[2]) if len(sys.argv) > 2 else 64
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index('Address'), hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
body = [r for r in rows[2:] if len(r) > iex and r[iex].isdigit()]
tot_ex = sum(int(r[iex]) for r in body)
tot_s = sum(int(r[isamp]) for r in body)
print('instructions', len(body), 'executed', tot_ex, 'samples', tot_s)
for i in range(0, len(body), win):
    blk = body[i:i + win]
    ex = sum(int(r[iex]) for r in blk)
    sm = sum(int(r[isamp]) for r in blk)
    top = max(blk, key=lambda r: int(r[isamp]))
    print('%5d..%5d  exec %5.1f%%  samples %5.1f%%   top: %s (%s)' % (i, i + len(blk) - 1, 100.0 * ex / tot_ex, 100.0 * sm / max(tot_s, 1),
                                                                      top[isrc].strip()[:60], top[isamp]))
