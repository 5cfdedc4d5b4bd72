#!/usr/bin/env python
"""Prints the handful of ncu raw-page metrics we steer by, one block per profiled launch.
usage: ncu -i X.ncu-rep --page raw --csv > X.raw.csv ; python tools/ncu_summary.py X.raw.csv"""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_issued.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sectors.sum', 'lts__t_sectors_op_atom.sum',
        'lts__t_sectors_op_red.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__cycles_active.avg',
        'l1tex__m_xbar2l1tex_read_sectors.sum', 'l1tex__m_l1tex2xbar_write_sectors.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active',
        'lts__t_sectors.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed_op_shared_atom.sum']

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if 'issue_stalled' in h and h.endswith('per_warp_active.pct')]
for r in rows[2:]:
    print('-----', r[idx['Kernel Name']][:60])
    for w in WANT:
        if w in idx:
            print('   %-70s %s %s' % (w, r[idx[w]], rows[1][idx[w]]))
    vals = []
    for h in stalls:
        try:
            vals.append((float(r[idx[h]].replace(',', '')), h))
        except ValueError:
            pass
    for v, h in sorted(vals, reverse=True)[:7]:
        print('      stall %6.2f%%  %s' % (v, h.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', '')))
