"""Per-step breakdown of the bench's e2e leg (pinned host file images -> H2D -> decode -> filter -> aggregate).

    gpurun -- python tools/probe_e2e.py [chunk_rows ...]

Prints, per step, the wall time and the scan operator's host/device counters so that a slow step can be attributed
to fetch (H2D issue), waiting on the prefetch thread, page-header parsing, or decode.
"""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from auron_b200 import proto as P
from auron_b200 import runtime

chunks = [int(a) for a in sys.argv[1:]] or [48_000_000]
tmp = tempfile.mkdtemp(prefix="probe_e2e_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
files = bench.gen_dataset(tmp, bench.SF100_ROWS)
paths = [f for f, _ in files]
sizes = [os.path.getsize(f) for f in paths]
pins = []
for p, sz in zip(paths, sizes):
    buf = torch.empty(sz, dtype=torch.uint8).pin_memory()
    with open(p, "rb") as fh:
        fh.readinto(memoryview(buf.numpy()))
    hp = "pinned://" + os.path.basename(p)
    runtime.put_host_file(hp, buf)
    pins.append(hp)
os.environ["AURON_PROFILE"] = "1"
hbm = []
modes = os.environ.get("PROBE_MODES", "pinned,files,hbm").split(",")
if "hbm" in modes:
    for p in paths:
        hp = "hbm://" + os.path.basename(p)
        with open(p, "rb") as fh:
            runtime.put_device_file(hp, fh.read())
        hbm.append(hp)
for mode, pl in (("pinned", pins), ("files", paths), ("hbm", hbm)):
    if mode not in modes:
        continue
    plan = bench.build_plan(P, pl, sizes)
    for ch in chunks:
        os.environ["AURON_GPU_CHUNK_ROWS"] = str(ch)
        print(f"== {mode} chunk_rows={ch}")
        for it in range(int(os.environ.get("PROBE_STEPS", "8"))):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            task = runtime.Task(plan)
            task.__enter__()
            t1 = time.perf_counter()
            out = list(task)
            t2 = time.perf_counter()
            m = task.metrics()
            task.__exit__(None, None, None)
            t3 = time.perf_counter()
            ms = 1000 * (t3 - t0)
            if os.environ.get("PROBE_DUMP") and it == int(os.environ.get("PROBE_STEPS", "8")) - 1:
                for depth, op, name, v in m:
                    print(f"    {'  ' * depth}{op}.{name} = {v}")
            scan = {name: v for _, op, name, v in m if op == "ParquetExec"}
            keys = ("elapsed_ns", "fetch_ns", "wait_fetch_ns", "parse_ns", "decode_ns")
            print(f"  step {it}: {ms:7.1f} ms  " + "  ".join(f"{k[:-3]}={scan.get(k, 0) / 1e6:.1f}ms" for k in keys)
                  + f"  create={1e3 * (t1 - t0):.1f} iter={1e3 * (t2 - t1):.1f} close={1e3 * (t3 - t2):.1f}  h2d_device={scan.get('h2d_device_us', 0) / 1e3:.1f}ms  h2d_GB={scan.get('h2d_bytes', 0) / 1e9:.2f}")
