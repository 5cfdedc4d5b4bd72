#!/bin/bash
# Profile pass of a round (run on the GPU box through gpurun): launch list of one bench run, one `ncu --set full` capture of the
# six kernels of a count at full size, their raw / source pages as CSV, and the bench line itself (never taken under the profiler).
# usage: bash tools/profile_round.sh r02
R=${1:-r02}
OUT=gpurun_out
python bench.py > $OUT/${R}_bench_n1_final.json 2> $OUT/${R}_bench_n1_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${R}_launches_final.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${R}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k1_superkmer_fast|k2a_dedup_split|k2b_warp_bins|k3s_pool_scatter|k3c_sort' -s 6 -c 6 \
    -f -o /tmp/${R}_prof_final python tools/bench_config.py 31 150 100000000 1 > $OUT/${R}_prof_run.log 2>&1
ncu -i /tmp/${R}_prof_final.ncu-rep --page raw --csv > $OUT/${R}_prof_final.raw.csv 2>/dev/null
ncu -i /tmp/${R}_prof_final.ncu-rep --page source --csv -k regex:k2b_warp_bins > $OUT/${R}_prof_k2b.source.csv 2>/dev/null
ncu -i /tmp/${R}_prof_final.ncu-rep --page source --csv -k regex:k1_superkmer_fast > $OUT/${R}_prof_k1.source.csv 2>/dev/null
ls -la /tmp/${R}_prof_final.ncu-rep $OUT/${R}_*
