#!/usr/bin/env python
"""bench.py -- BASELINE.json config[1]: ParquetScan -> Filter -> HashAggregate (GROUP BY int64, SUM/COUNT) over a
synthetic TPC-DS SF100 `store_sales` (287,997,024 rows), one step = one pass of the whole plan.

  python bench.py --gpus N --steps K --warmup W            (N>1 via torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                     (CPU arm: Arrow C++ scan/filter + the oracle's C aggregate)

`value`  : rows/s with the Parquet file images already resident in HBM (decode -> filter -> aggregate on device).
`e2e`    : rows/s through the C ABI with the Parquet file images in pinned HOST memory: H2D of the encoded column
           chunks -> decode -> filter -> aggregate -> D2H of the result, every step.  `e2e.page_cache_files` is the
           same plan over plain files (pread from the page cache -> pinned staging -> H2D).
`roofline`: dominant kernel of the timed steps, algorithmic bytes / device time from CUDA events recorded on the
           launching stream inside the library (AURON_PROFILE=1), against MEASURED_PEAKS.json hbm_gbs.
The oracle is used only by the cpu_baseline / --impl reference legs (as the timed CPU arm), never by the product path.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pyarrow.parquet as pq

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SF100_ROWS = 287_997_024
ROWS_PER_FILE = 16_000_000
DATE_LO, DATE_HI = 2450816, 2452642          # ss_sold_date_sk window (~5 years)
FILTER_LO, FILTER_HI = 2451000, 2452000      # WHERE ss_sold_date_sk >= lo AND < hi
N_ITEMS = 204_000                            # item cardinality at SF100
CODEC = os.environ.get("AURON_BENCH_CODEC", "SNAPPY")   # page compression of the synthetic files (SNAPPY = Spark default; NONE = uncompressed)
SCHEMA = pa.schema([("ss_item_sk", pa.int32()), ("ss_quantity", pa.int32()), ("ss_sold_date_sk", pa.int32())])


def gen_file(path: str, rows: int, seed: int):
    rng = np.random.default_rng(seed)
    t = pa.table({
        "ss_item_sk": pa.array(rng.integers(1, N_ITEMS + 1, rows, dtype=np.int32)),
        "ss_quantity": pa.array(rng.integers(1, 101, rows, dtype=np.int32), mask=rng.random(rows) < 0.03),
        "ss_sold_date_sk": pa.array(rng.integers(DATE_LO, DATE_HI, rows, dtype=np.int32), mask=rng.random(rows) < 0.04),
    }, schema=SCHEMA)
    # Spark's writer defaults: SNAPPY pages, dictionary encoding, ~128 MB row groups (8M rows x 3 projected columns)
    pq.write_table(t, path, compression=CODEC, use_dictionary=True, row_group_size=8_000_000, data_page_size=1 << 20)


def gen_dataset(directory: str, total_rows: int) -> list[tuple[str, int]]:
    os.makedirs(directory, exist_ok=True)
    specs, left, i = [], total_rows, 0
    while left > 0:
        r = min(ROWS_PER_FILE, left)
        specs.append((os.path.join(directory, f"store_sales_{CODEC.lower()}_{i:03d}.parquet"), r, 42 + i))
        left -= r
        i += 1
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        list(ex.map(lambda s: gen_file(*s) if not os.path.exists(s[0]) else None, specs))
    return [(p, r) for p, r, _ in specs]


def build_plan(P, files: list[str], sizes: list[int]) -> bytes:
    scan = P.parquet_scan(SCHEMA, list(zip(files, sizes)), [0, 1, 2])
    flt = P.filter_(scan, [P.binary("GtEq", P.col("ss_sold_date_sk"), P.lit(FILTER_LO, pa.int32())),
                           P.binary("Lt", P.col("ss_sold_date_sk"), P.lit(FILTER_HI, pa.int32()))])
    agg = P.agg(flt, [P.try_cast(P.col("ss_item_sk"), pa.int64())], ["ss_item_sk"],
                [P.agg_expr("SUM", [P.col("ss_quantity")], pa.int64()), P.agg_expr("COUNT", [P.col("ss_quantity")], pa.int64())],
                ["sum_qty", "cnt_qty"], ["PARTIAL", "PARTIAL"])
    return P.task_definition(agg)


def bind_to_gpu_numa_node(torch, gpu_index: int):
    """One process per GPU: run on the cores of the GPU's NUMA node, so that the pinned host buffers (first touch) and the
    scan's worker threads sit next to the PCIe root the GPU hangs off.  Without it the 8-GPU e2e leg is bound by
    cross-socket traffic.  Returns the node id or None (single-node boxes, missing sysfs entries)."""
    try:
        p = torch.cuda.get_device_properties(gpu_index)
        if hasattr(p, "pci_bus_id"):
            bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        else:   # older torch: NVML reports the same bus id string
            import pynvml
            pynvml.nvmlInit()
            bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(gpu_index)).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            bdf = bus[-12:].lower()
        base = f"/sys/bus/pci/devices/{bdf}"
        node = int(open(base + "/numa_node").read())
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if node >= 0 and cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (profiling recipe), through NVML in a background thread
    (the same counters `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints, without a subprocess per sample)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int, period_s: float = 0.02):
        self.sm, self.mx, self.reasons, self.idx, self.period = [], None, set(), gpu_index, period_s
        self._stop = threading.Event()
        self._thr = None

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.idx]) if vis and vis.split(",")[self.idx].isdigit() else self.idx
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def loop():
                while not self._stop.is_set():
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                        for bit, name in self.REASONS.items():
                            if mask & bit:
                                self.reasons.add(name)
                    except Exception:
                        pass
                    self._stop.wait(self.period)

            self._thr = threading.Thread(target=loop, daemon=True)
            self._thr.start()
        except Exception:
            self._thr = None
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=1)

    def summary(self):
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                "samples": len(self.sm)}


def cpu_pipeline(files: list[str]) -> tuple[int, float]:
    """The CPU arm: Arrow C++ Parquet reader + filter (stand-in for the parquet crate / arrow-rs kernels the reference
    delegates to) feeding the oracle's C hash aggregate (port of agg_hash_map.rs / sum.rs / count.rs), all host threads."""
    import oracle
    rows = 0
    t0 = time.perf_counter()
    for f in files:
        t = pq.read_table(f, columns=["ss_item_sk", "ss_quantity", "ss_sold_date_sk"], use_threads=True)
        d = t["ss_sold_date_sk"]
        mask = pc.and_kleene(pc.greater_equal(d, FILTER_LO), pc.less(d, FILTER_HI))
        ft = t.filter(mask)
        oracle.agg_sum_count_i64(ft["ss_item_sk"].combine_chunks().cast(pa.int64()), ft["ss_quantity"].combine_chunks().cast(pa.int64()))
        rows += t.num_rows
    return rows, time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="auron")
    ap.add_argument("--rows", type=int, default=SF100_ROWS)
    ap.add_argument("--data-dir", default=os.path.join(tempfile.gettempdir(), "auron_b200_bench"))
    ap.add_argument("--skip-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    config = {"workload": "BASELINE configs[1]: ParquetScan->Filter->HashAggregate(GROUP BY int64 ss_item_sk, SUM/COUNT ss_quantity), "
                          "synthetic TPC-DS SF100 store_sales", "rows": args.rows, "groups": N_ITEMS, "filter_selectivity": "~0.53",
              "parquet": f"3 INT32 columns, RLE_DICTIONARY + PLAIN fallback pages, 8M-row (~128 MB) row groups, {CODEC} pages "
                         "(decompressed on the GPU)",
              "l2_policy": "inputs (>=1.4 GB encoded, 3.4 GB decoded per step) are far larger than the 126 MB L2",
              "parallelism": f"dp{args.gpus}: table partitions sharded per GPU, no data-path collective; one process per GPU bound to the GPU's NUMA node"}

    if args.impl == "reference":
        if rank != 0:
            return
        files = gen_dataset(args.data_dir, args.rows)
        sample = [f for f, _ in files[:2]]
        for _ in range(args.warmup):
            cpu_pipeline(sample[:1])
        rows, secs = 0, 0.0
        for _ in range(args.steps):
            r, s = cpu_pipeline(sample)
            rows += r
            secs += s
        v = rows / secs
        print(json.dumps({"impl": "reference", "metric": "rows_per_sec", "value": v, "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 1000 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "int64", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": "rows/s", "cores": cores, "kind": "port",
                                           "sample": f"{len(sample)} of {len(files)} files ({sum(r for _, r in files[:2])} rows) per step"},
                          "e2e": {"value": v, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(torch, local_rank)    # before any pinned allocation / worker thread exists
    config["numa_node"] = numa
    from auron_b200 import proto as P
    from auron_b200 import runtime

    if rank == 0:
        files = gen_dataset(args.data_dir, args.rows)
    if world > 1:
        dist.barrier()
    files = gen_dataset(args.data_dir, args.rows)       # no-op when the files exist
    paths = [f for f, _ in files]
    sizes = [os.path.getsize(f) for f in paths]
    total_rows = sum(r for _, r in files)
    h2d_bytes = unc_bytes = 0      # column-chunk bytes as stored (what crosses PCIe) / after page decompression
    for f in paths:
        md = pq.ParquetFile(f).metadata
        for g in range(md.num_row_groups):
            for c in range(md.num_columns):
                h2d_bytes += md.row_group(g).column(c).total_compressed_size
                unc_bytes += md.row_group(g).column(c).total_uncompressed_size

    os.environ["AURON_PROFILE"] = "1"
    # device chunk = the whole SF100 partition set of this GPU (3.4 GB decoded; HBM is 180 GB)
    os.environ.setdefault("AURON_GPU_CHUNK_ROWS", str(320_000_000))

    def barrier_sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_ms: list[float] = []

    def spread():
        v = sorted(step_ms)
        return {"min": v[0], "median": v[len(v) // 2], "max": v[-1]} if v else None

    def run_steps(plan: bytes, steps: int, collect: bool):
        kern, launches, out_bytes = {}, 0, 0
        step_ms.clear()
        barrier_sync()
        t0 = time.perf_counter()
        for _ in range(steps):
            ts = time.perf_counter()
            with runtime.Task(plan, device=local_rank) as task:
                out = pa.Table.from_batches(list(task), schema=task.schema)
                out_bytes = out.nbytes
                if collect or os.environ.get("AURON_BENCH_VERBOSE"):
                    for depth, op, name, v in task.metrics():
                        if op == "__kernels__":
                            if collect:
                                kern[name] = kern.get(name, 0) + v
                        elif os.environ.get("AURON_BENCH_VERBOSE"):
                            print(f"[metric{'' if collect else ' e2e'}] {op}.{name} = {v}", file=sys.stderr)
            step_ms.append(1000 * (time.perf_counter() - ts))
        barrier_sync()
        dt = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        return dt, kern, out_bytes, out

    # ---- value: file images resident in HBM
    hbm_paths = [f"hbm://{os.path.basename(p)}@{local_rank}" for p in paths]
    for p, hp in zip(paths, hbm_paths):
        with open(p, "rb") as fh:
            runtime.put_device_file(hp, fh.read(), device=local_rank)
    plan_hbm = build_plan(P, hbm_paths, sizes)
    run_steps(plan_hbm, args.warmup, False)
    with ClockSampler(local_rank) as cs:
        dt, kern, out_bytes, out = run_steps(plan_hbm, args.steps, True)
    clocks = cs.summary()
    value = world * total_rows * args.steps / dt
    value_spread = spread()
    for hp in hbm_paths:
        runtime.drop_device_file(hp)

    # ---- e2e: the same call with HOST inputs; every step uploads the projected column chunks inside the timed region
    #   e2e            : file images in pinned host memory (the contract's "from pinned host memory"): H2D per chunk
    #   e2e_page_cache : plain files (OS page cache): pread -> pinned staging -> H2D, overlapped with decode
    e2e = None
    if not args.skip_e2e:
        # smaller device batches so that uploading batch k+1 overlaps decoding batch k
        os.environ["AURON_GPU_CHUNK_ROWS"] = os.environ.get("AURON_E2E_CHUNK_ROWS", str(48_000_000))
        pin_paths = [f"pinned://{os.path.basename(p)}@{local_rank}" for p in paths]
        for p, hp, sz in zip(paths, pin_paths, sizes):
            buf = torch.empty(sz, dtype=torch.uint8).pin_memory()
            with open(p, "rb") as fh:
                fh.readinto(memoryview(buf.numpy()))
            runtime.put_host_file(hp, buf)
        plan_pin = build_plan(P, pin_paths, sizes)
        run_steps(plan_pin, max(1, args.warmup), False)
        dte, _, out_bytes_e, _ = run_steps(plan_pin, args.steps, False)
        e2e = {"value": world * total_rows * args.steps / dte, "unit": "rows/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": out_bytes_e,
               "ms_per_step": 1000 * dte / args.steps, "step_ms": spread(), "input": "parquet file images in pinned host memory"}
        for hp in pin_paths:
            runtime.drop_host_file(hp)
        plan_host = build_plan(P, paths, sizes)
        run_steps(plan_host, 1, False)
        dtf, _, _, _ = run_steps(plan_host, args.steps, False)
        e2e["page_cache_files"] = {"value": world * total_rows * args.steps / dtf, "ms_per_step": 1000 * dtf / args.steps, "step_ms": spread(),
                                   "input": "parquet files in the OS page cache (pread into pinned staging, then H2D)"}

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    # ---- roofline of the dominant kernel (device time from CUDA events on the launching stream)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    names = sorted({k.rsplit(".", 1)[0] for k in kern if k.endswith(".device_us")}, key=lambda n: -kern[n + ".device_us"])
    # partial-mode output = [group, sum acc, count acc] (accumulator fields are unnamed, agg_ctx.rs:127-150)
    sel_rows = int(out.column(2).to_numpy().sum() / 0.97) if out.num_rows else 0        # filtered rows reaching the aggregate (approx)
    decoded_bytes = total_rows * 12 + 2 * total_rows // 8                               # 3 x int32 out + 2 validity bitmaps
    alg = {   # algorithmic bytes per step for each launch site (inputs once + outputs once)
        "pq_decompress": (h2d_bytes + unc_bytes) if CODEC != "NONE" else None,
        "pq_decode_pages": unc_bytes + decoded_bytes,
        "simple_predicate": total_rows * 4 + 2 * (total_rows // 8),                 # date column + its validity in, mask out
        "agg_key_range": total_rows * 4 + total_rows // 8,                            # key column + validity
        "agg_update": total_rows // 8 + sel_rows * (4 + 4) + sel_rows // 8,          # mask + selected (key, value) + value validity
    }
    roofs = []
    for n in names:
        us = kern[n + ".device_us"] / args.steps
        r = {"kernel": n, "device_ms_per_step": us / 1000.0, "launches_per_step": kern[n + ".launches"] / args.steps,
             "share_of_step": (us / 1e6) / (dt / args.steps)}
        if alg.get(n) and us > 0:
            r["algorithmic_bytes_per_step"] = alg[n]
            r["achieved_gbs"] = alg[n] / (us * 1e-6) / 1e9
            r["frac_of_peak"] = r["achieved_gbs"] / peak
        roofs.append(r)
    dom = names[0] if names else None
    dom_us = kern[dom + ".device_us"] / args.steps if dom else None
    dom_bytes = alg.get(dom) if dom else None
    # DRAM traffic of the dominant kernel per step, from the committed ncu pass at this workload's full size (profiles/)
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if dom in tj and args.rows == SF100_ROWS and CODEC == "SNAPPY":
            traffic, traffic_src = tj[dom]["dram_bytes_per_step"], tj[dom]["source"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": (dom_bytes / (dom_us * 1e-6) / 1e9) if dom_bytes and dom_us else None, "peak": peak,
                "unit": "GB/s", "frac": (dom_bytes / (dom_us * 1e-6) / 1e9 / peak) if dom_bytes and dom_us else None, "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": peak_src, "algorithmic_bytes_per_step": dom_bytes, "kernels": roofs}

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only)
    cpu = None
    if world == 1:
        sample = paths[:2]
        cpu_pipeline(sample[:1])
        r, s = cpu_pipeline(sample)
        cpu = {"value": r / s, "unit": "rows/s", "cores": cores, "kind": "port",
               "sample": f"{len(sample)} of {len(paths)} files ({r} rows): Arrow C++ scan+filter, oracle C hash aggregate"}

    line = {"metric": "rows_per_sec", "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000 * dt / args.steps, "step_ms": value_spread, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
            "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(kern.get("total_launches", 0)), "roofline": roofline, "cpu_baseline": cpu,
            "result_groups": out.num_rows, "selected_rows_est": sel_rows}
    print(json.dumps(line))


def _json_only_stdout():
    """Everything libraries print to stdout while the bench runs (NCCL's "NCCL version ..." banner comes from C code) is sent
    to stderr; only the JSON line reaches the real stdout."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)


if __name__ == "__main__":
    _json_only_stdout()
    main()
