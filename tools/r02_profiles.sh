#!/bin/bash
# Round-2 evidence at batch 256 (10 M docs, 1 x B200): plain run, then the launch list, then ncu --set full of the main kernels
TAG=${1:-r02_f}
O=gpurun_out
CMD="python bench.py --batch 256 --steps 2 --warmup 3 --no-cpu-baseline --no-modes --verify 0"
timeout 300 $CMD > $O/plain256.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:ais:: -c 900 --csv --log-file $O/${TAG}_launches_batch256_10M.csv $CMD > $O/ncu_l256.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"scan_pair_kernel|bm25_score_kernel|bm25_combine_kernel|rerank_max_kernel|collect_fast_kernel|bm25_slices_kernel" -s 6 -c 7 -o $O/${TAG}_main_kernels $CMD > $O/ncu_d256.log 2>&1
tail -2 $O/ncu_d256.log | cut -c1-160
ls -la $O/${TAG}_*
