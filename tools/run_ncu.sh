#!/bin/bash
# Profiling recipe for the bench's hot path (run on the GPU box through gpurun; outputs land in gpurun_out/).
#   tools/run_ncu.sh <tag>
# 1. launch list of every kernel of a short scan_agg bench run (shares of the step), 2. one `--set full` capture of the kernels of
# one step (three batches: Snappy prefix + warp pass, scout, TMA-staged scan kernel, tile kernel, merge), 3. DRAM traffic of the
# fused kernels of one FULL-SIZE step, 4. launch lists of the join and sort/shuffle workloads.
# Numbers printed by bench.py under ncu are NOT bench values.
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
export AURON_BENCH_NO_CPU=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_scan_agg.csv \
    python bench.py --workload scan_agg --steps 2 --warmup 1 --skip-e2e > $OUT/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'fz_staged_kernel|fz_kernel|fz_scout_kernel|pq_decompress|fz_merge_kernel' \
    -c 18 -f -o $OUT/${TAG}_prof python bench.py --workload scan_agg --steps 1 --warmup 1 --skip-e2e > $OUT/${TAG}_ncu_full.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'fz_staged_kernel|fz_kernel' -s 6 -c 6 --csv \
    --log-file $OUT/${TAG}_fused_traffic_full.csv python bench.py --workload scan_agg --steps 1 --warmup 1 --skip-e2e > $OUT/${TAG}_ncu_traffic.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_join.csv \
    python bench.py --workload join --steps 2 --warmup 1 > $OUT/${TAG}_ncu_join.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_sort_shuffle.csv \
    python bench.py --workload sort_shuffle --steps 2 --warmup 1 > $OUT/${TAG}_ncu_sort_shuffle.log 2>&1
ls -la $OUT/${TAG}_*
