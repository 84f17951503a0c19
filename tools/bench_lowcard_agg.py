"""Micro-benchmark: low-cardinality GROUP BY (400 groups, 64M rows, SUM + COUNT) on device-resident input.
Compares the default L2-atomic aggregate kernel with the opt-in shared-memory pre-aggregation variant
(AURON_ENABLE_SMEM_AGG=1).  Device time from CUDA events inside the library (AURON_PROFILE=1).

    gpurun -- python tools/bench_lowcard_agg.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pyarrow as pa

from auron_b200 import proto as P
from auron_b200 import runtime

rng = np.random.default_rng(1)
n = 64_000_000
t = pa.table({"k": pa.array(rng.integers(0, 400, n), type=pa.int32()), "v": pa.array(rng.integers(-1000, 1000, n), type=pa.int64())})
runtime.put_device_batch("lc", t.to_batches()[0])
plan = P.task_definition(P.agg(P.ffi_reader(t.schema, "lc"), [P.col("k")], ["k"],
                               [P.agg_expr("SUM", [P.col("v")], pa.int64()), P.agg_expr("COUNT", [P.col("v")], pa.int64())], ["s", "c"], ["PARTIAL"] * 2))
os.environ["AURON_PROFILE"] = "1"
for mode in ("plain", "smem"):
    if mode == "smem":
        os.environ["AURON_ENABLE_SMEM_AGG"] = "1"
    for it in range(3):
        with runtime.Task(plan) as task:
            out = pa.Table.from_batches(list(task), schema=task.schema)
            m = task.metrics()
    print(mode, "groups", out.num_rows, [(name, v) for _, op, name, v in m if name.startswith("agg_update")])
