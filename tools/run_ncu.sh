#!/bin/bash
# Profiling recipe for the bench's hot path (run on the GPU box through gpurun; outputs land in gpurun_out/).
#   tools/run_ncu.sh <tag> [rows]
# 1. launch list of every kernel of a short bench run (shares of the step), 2. one `--set full` capture of the
# dominant kernels, 3. DRAM traffic of the decode kernel at the FULL bench size (three launches = one step).
# Numbers printed by bench.py under ncu are NOT bench values.
set -u
TAG=${1:-r01}
ROWS=${2:-64000000}
OUT=gpurun_out
mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --rows $ROWS --steps 2 --warmup 1 --skip-e2e > $OUT/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'pq_decode_tiles_fast_kernel|pq_scout_kernel|pq_decompress_kernel|agg_direct_kernel|agg_fast_kernel|key_minmax|interval_predicate' \
    -c 14 -f -o $OUT/${TAG}_prof python bench.py --rows $ROWS --steps 1 --warmup 1 --skip-e2e > $OUT/${TAG}_ncu_full.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'pq_decode_tiles_fast_kernel' -s 3 -c 3 --csv \
    --log-file $OUT/${TAG}_decode_traffic_full.csv python bench.py --steps 1 --warmup 1 --skip-e2e > $OUT/${TAG}_ncu_traffic.log 2>&1
ls -la $OUT/${TAG}_*
