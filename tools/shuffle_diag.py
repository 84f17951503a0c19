"""Where a ShuffleWriterExec step spends its time: the operator's own metrics (compress_ns, write_ns, ...) next to the wall clock.
usage: python tools/shuffle_diag.py [rows]"""
import os
import sys
import time

import numpy as np
import pyarrow as pa

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from auron_b200 import proto as P
from auron_b200 import runtime

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64_000_000
rng = np.random.default_rng(100)
keep: list = []
t4 = pa.table({"ss_item_sk": bench.pinned_array(torch, rng.integers(1, bench.N_ITEMS + 1, n, dtype=np.int32), pa.int32(), None, keep),
               "ss_ticket_number": bench.pinned_array(torch, rng.integers(1, 240_000_000, n, dtype=np.int64), pa.int64(), None, keep),
               "ss_ext_sales_price": bench.pinned_array(torch, bench.decimal_words(rng.integers(0, 2_000_000, n, dtype=np.int64)), pa.decimal128(7, 2), None, keep)})
for b in t4.to_batches(max_chunksize=16_000_000):
    runtime.put_device_batch("diag", b, device=0)
d = "/dev/shm/auron_bench_shuffle"
os.makedirs(d, exist_ok=True)
for it in range(4):
    td = P.task_definition(P.shuffle_writer(P.ffi_reader(t4.schema, "diag"), P.hash_repartition([P.col("ss_item_sk")], 200), f"{d}/diag{it}.data", f"{d}/diag{it}.index"))
    t0 = time.perf_counter()
    with runtime.Task(td) as task:
        for _ in task:
            pass
        t1 = time.perf_counter()
        m = task.metrics()
    print(f"run {it}: {1000 * (t1 - t0):.1f} ms", {name: round(v / 1e6, 2) if name.endswith("_ns") else v for _, _, name, v in m if v})
for it in range(4):
    for ext in ("data", "index"):
        os.remove(f"{d}/diag{it}.{ext}")
runtime.drop_device_resource("diag")
