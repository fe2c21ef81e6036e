#!/bin/bash
# Final evidence run of a round: bench lines, launch lists, ncu --set full of the main kernels -> gpurun_out/<tag>_*
TAG=${1:-r01_d}
O=gpurun_out
python bench.py > $O/${TAG}_bench_10M_batch64.json 2> $O/${TAG}_bench_b64.err
for b in 1 32 128; do python bench.py --batch $b --no-cpu-baseline --no-modes > $O/${TAG}_bench_10M_batch$b.json 2>/dev/null; done
python bench.py --prf full --no-cpu-baseline --no-modes > $O/${TAG}_bench_10M_batch64_prf_full.json 2>/dev/null
python bench.py --impl reference --steps 6 --warmup 1 > $O/${TAG}_bench_reference.json 2>/dev/null
CMD="python bench.py --batch 64 --steps 2 --warmup 3 --no-cpu-baseline --no-modes"
$CMD > $O/plain64.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:ais:: -c 900 --csv --log-file $O/${TAG}_launches_batch64_10M.csv $CMD > $O/ncu_l64.log 2>&1
CMD1="python bench.py --batch 1 --steps 3 --warmup 3 --no-cpu-baseline --no-modes"
$CMD1 > $O/plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:ais:: -c 400 --csv --log-file $O/${TAG}_launches_batch1_10M.csv $CMD1 > $O/ncu_l1.log 2>&1
$CMD > $O/plain64.log 2>&1 && ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"scan_tc_kernel|bm25_score_kernel|bm25_combine_kernel|segmax_kernel|column_scan_kernel" -s 12 -c 6 -o $O/${TAG}_main_kernels $CMD > $O/ncu_d.log 2>&1
CMD32="python bench.py --batch 32 --prf full --steps 2 --warmup 3 --no-cpu-baseline --no-modes"
$CMD32 > $O/plain32.log 2>&1 && ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:scan_tc_kernel -s 4 -c 1 -o $O/${TAG}_scan_tc32 $CMD32 > $O/ncu_32.log 2>&1
tail -2 $O/ncu_d.log | cut -c1-120
