"""Per-operator throughput of the other BASELINE configs on ONE GPU, inputs resident in HBM (device resources), device
time per launch site from the library's CUDA-event timers (AURON_PROFILE=1).  These are the operator-level numbers the
north star asks for next to the config-2 bench line; they are not bench.py lines.

    gpurun -- python tools/bench_ops.py [join] [sort] [shuffle] [agg_lowcard]

  join     cfg 3: store_sales (N rows: ss_sold_date_sk int32, ss_item_sk int32, ss_quantity int32) JOIN date_dim (73,049 rows:
           d_date_sk int32, d_year int32) on the date key, inner, build = date_dim
  sort     cfg 4 (one GPU's share): ORDER BY ss_item_sk over (ss_item_sk int32, ss_ticket_number int64, ss_ext_sales_price decimal(7,2))
  shuffle  cfg 4: hash repartition of the same rows on ss_item_sk into 200 partitions, compacted shuffle format to /dev/shm
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import decimal

import numpy as np
import pyarrow as pa

from auron_b200 import proto as P
from auron_b200 import runtime

os.environ["AURON_PROFILE"] = "1"
N = int(os.environ.get("OPS_ROWS", 64_000_000))
CHUNK = 16_000_000
which = sys.argv[1:] or ["join", "sort", "shuffle"]
rng = np.random.default_rng(42)


def put(resource, table):
    for b in table.to_batches(max_chunksize=CHUNK):
        runtime.put_device_batch(resource, b)


PEAK = 6575.1   # GB/s, MEASURED_PEAKS.json hbm_gbs of this pool's B200s


def run(plan, label, rows, steps=4, alg=None):
    """alg: {launch site: algorithmic bytes per pass} (inputs once + outputs once)"""
    td = P.task_definition(plan)
    best = None
    for it in range(steps):
        t0 = time.perf_counter()
        with runtime.Task(td) as task:
            out_rows = 0
            for b in task:
                out_rows += b.num_rows
            m = task.metrics()
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, out_rows, m)
    dt, out_rows, m = best
    print(f"== {label}: {rows / dt / 1e6:.0f} Mrows/s end to end ({1000 * dt:.1f} ms per pass, {out_rows} rows out, result copied to the host)")
    kern = {}
    for _, op, name, v in m:
        if op == "__kernels__" and name.endswith(".device_us"):
            kern[name[:-10]] = v
    for k, v in sorted(kern.items(), key=lambda kv: -kv[1]):
        extra = ""
        if alg and alg.get(k) and v > 0:
            gbs = alg[k] / (v * 1e-6) / 1e9
            extra = f"   {alg[k] / 1e9:7.2f} GB algorithmic -> {gbs:7.0f} GB/s = {100 * gbs / PEAK:4.1f} % of {PEAK:.0f}"
        print(f"     {k:28s} {v / 1000:9.3f} ms{extra}")
    for _, op, name, v in m:
        if op != "__kernels__" and name.endswith("_ns") and v > 2e5:
            print(f"     [{op}.{name} = {v / 1e6:.2f} ms]")


if "join" in which:
    date_lo = 2450816
    dd = pa.table({"d_date_sk": pa.array(np.arange(2415022, 2415022 + 73049, dtype=np.int32)),
                   "d_year": pa.array((1900 + np.arange(73049) // 365).astype(np.int32))})
    ss = pa.table({"ss_sold_date_sk": pa.array(rng.integers(date_lo, date_lo + 1826, N, dtype=np.int32), mask=rng.random(N) < 0.04),
                   "ss_item_sk": pa.array(rng.integers(1, 204001, N, dtype=np.int32)),
                   "ss_quantity": pa.array(rng.integers(1, 101, N, dtype=np.int32))})
    put("ss_join", ss)
    put("dd_join", dd)
    out_schema = pa.schema(list(dd.schema) + list(ss.schema))
    # the join output stays on the device: aggregate it to one row so that the measurement is the join, not a 1 GB D2H
    j = P.hash_join(out_schema, P.ffi_reader(dd.schema, "dd_join"), P.ffi_reader(ss.schema, "ss_join"),
                    [(P.col("d_date_sk"), P.col("ss_sold_date_sk"))], "INNER", "LEFT")
    plan = P.agg(j, [], [], [P.agg_expr("SUM", [P.col("ss_quantity")], pa.int64()), P.agg_expr("SUM", [P.col("d_year")], pa.int64()),
                             P.agg_expr("COUNT", [P.col("ss_item_sk")], pa.int64())], ["q", "y", "c"], ["PARTIAL"] * 3)
    matched = int(N * 0.96)
    run(plan, f"cfg3 HashJoin build=date_dim(73,049) probe={N} rows + global SUM/COUNT of the joined rows", N,
        alg={"join_probe": N * 4 + N // 8 + matched * 8,           # probe keys + validity in, (probe row, build row) pairs out
             "take": matched * 2 * 20,                              # 5 int32 output columns gathered: row bytes in + out
             "join_build": 73049 * 4 * 2})
    runtime.drop_device_resource("ss_join")
    runtime.drop_device_resource("dd_join")

if "sort" in which or "shuffle" in which:
    price = rng.integers(0, 2_000_000, N)
    # decimal(7,2) column built from unscaled integers without a Python loop
    unscaled = price.astype(np.int64)
    lo = unscaled.view(np.uint64)
    hi = np.where(unscaled < 0, np.uint64(0xFFFFFFFFFFFFFFFF), np.uint64(0))
    buf = np.empty(2 * N, dtype=np.uint64)
    buf[0::2] = lo
    buf[1::2] = hi
    dec = pa.Array.from_buffers(pa.decimal128(7, 2), N, [None, pa.py_buffer(buf.tobytes())])
    t4 = pa.table({"ss_item_sk": pa.array(rng.integers(1, 204001, N, dtype=np.int32)),
                   "ss_ticket_number": pa.array(rng.integers(1, 240_000_000, N, dtype=np.int64)),
                   "ss_ext_sales_price": dec})
    put("t4", t4)
    if "sort" in which:
        # ORDER BY + LIMIT keeps the sort complete (all rows are ordered) but returns only the head to the host
        plan = P.sort(P.ffi_reader(t4.schema, "t4"), [P.sort_expr(P.col("ss_item_sk"))], limit=1000)
        run(plan, f"cfg4 SortExec ORDER BY ss_item_sk LIMIT 1000 over {N} rows x 28 B (top-k path)", N)
        plan = P.agg(P.sort(P.ffi_reader(t4.schema, "t4"), [P.sort_expr(P.col("ss_item_sk"))]), [], [],
                     [P.agg_expr("COUNT", [P.col("ss_item_sk")], pa.int64())], ["c"], ["PARTIAL"])
        run(plan, f"cfg4 SortExec full ORDER BY ss_item_sk over {N} rows x 28 B (+ COUNT so that only one row leaves the GPU)", N,
            alg={"radix_sort": 3 * 2 * 12 * N,                      # 18-bit key: 3 executed 8-bit passes x (u64 word + i32 row) read + write
                 "take": 2 * 28 * N})
    if "shuffle" in which:
        d = "/dev/shm/auron_ops_shuffle"
        os.makedirs(d, exist_ok=True)
        plan = P.shuffle_writer(P.ffi_reader(t4.schema, "t4"), P.hash_repartition([P.col("ss_item_sk")], 200), f"{d}/s.data", f"{d}/s.index")
        run(plan, f"cfg4 ShuffleWriterExec hash(ss_item_sk) -> 200 partitions, {N} rows x 28 B, LZ4 blocks written to /dev/shm", N, steps=3,
            alg={"murmur3_partition_ids": 8 * N, "partition_rows": 8 * N, "take": 2 * 28 * N, "serde_write": 2 * 28 * N,
                 "lz4_compress": 28 * N + os.path.getsize(d + "/s.data") if os.path.exists(d + "/s.data") else 28 * N})
        print(f"     shuffle file: {os.path.getsize(d + '/s.data') / 1e6:.0f} MB")
        # read side: IpcReaderExec over all 200 segments of the file just written (+ COUNT so that one row leaves the GPU)
        import struct
        offs = struct.unpack("<201q", open(f"{d}/s.index", "rb").read())
        blocks = [(f"{d}/s.data", offs[p], offs[p + 1] - offs[p]) for p in range(200)]
        rplan = P.agg(P.ipc_reader(t4.schema, "shuffle_in"), [], [], [P.agg_expr("COUNT", [P.col("ss_item_sk")], pa.int64()),
                                                                      P.agg_expr("SUM", [P.col("ss_ticket_number")], pa.int64())], ["c", "s"], ["PARTIAL"] * 2)
        td = P.task_definition(rplan)
        best = None
        for it in range(3):
            t0 = time.perf_counter()
            with runtime.Task(td, shuffle_blocks={"shuffle_in": list(blocks)}) as task:
                out = pa.Table.from_batches(list(task), schema=task.schema)
                m = task.metrics()
            dt = time.perf_counter() - t0
            if best is None or dt < best[0]:
                best = (dt, out, m)
        dt, out, m = best
        print(f"== cfg4 IpcReaderExec (shuffle read) of the same file: {N / dt / 1e6:.0f} Mrows/s ({1000 * dt:.1f} ms, count = {out.column(0).to_pylist()})")
        for _, op, name, v in m:
            if op != "__kernels__" and name.endswith("_ns") and v > 2e5:
                print(f"     [{op}.{name} = {v / 1e6:.2f} ms]")
    runtime.drop_device_resource("t4")
