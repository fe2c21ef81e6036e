#!/bin/bash
# Reduced evidence refresh: default bench line, batch 1 / 32, launch list and ncu --set full of the main kernels at batch 64
TAG=${1:-r01_e}
O=gpurun_out
python bench.py > $O/${TAG}_bench_10M_batch64.json 2> $O/${TAG}_bench_b64.err
for b in 1 32; do python bench.py --batch $b --no-cpu-baseline --no-modes > $O/${TAG}_bench_10M_batch$b.json 2>/dev/null; done
CMD="python bench.py --batch 64 --steps 2 --warmup 3 --no-cpu-baseline --no-modes"
$CMD > $O/plain64.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:ais:: -c 900 --csv --log-file $O/${TAG}_launches_batch64_10M.csv $CMD > $O/ncu_l64.log 2>&1
$CMD > $O/plain64.log 2>&1 && ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"scan_tc_kernel|bm25_score_kernel|bm25_combine_kernel|segmax_kernel" -s 8 -c 4 -o $O/${TAG}_main_kernels $CMD > $O/ncu_d.log 2>&1
tail -1 $O/ncu_d.log | cut -c1-120
