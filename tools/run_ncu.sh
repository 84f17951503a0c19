#!/bin/bash
# Profiling recipe for the bench's hot path (run on the GPU box through gpurun; outputs land in gpurun_out/).
#   tools/run_ncu.sh <tag> [rows]
# 1. launch list of every kernel of a short bench run (shares of the step), 2. one `--set full` capture of the
# dominant kernels.  Numbers printed by bench.py under ncu are NOT bench values.
set -u
TAG=${1:-r01}
ROWS=${2:-64000000}
OUT=gpurun_out
mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --rows $ROWS --steps 2 --warmup 1 --skip-e2e > $OUT/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'pq_decode_tiles_fast_kernel|pq_scout_kernel|agg_fast_kernel|interval_predicate_kernel|simple_predicate_kernel|mask_' \
    -c 12 -f -o $OUT/${TAG}_prof python bench.py --rows $ROWS --steps 1 --warmup 1 --skip-e2e > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT/${TAG}_*
