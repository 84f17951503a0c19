"""Two-stage aggregate across GPUs, one process per GPU (run under torch.distributed.run):

    FinalAgg( ShuffleWriter[nccl://, hash(k), P partitions]( PartialAgg( FFIReader(shard of this rank) ) ) )

Every rank aggregates its shard, the partial rows are hash-repartitioned on device and exchanged with an NCCL
all-to-all-v over NVLink (exchange.cu), and each rank finishes the groups of the partitions it owns.  Rank 0 gathers
the per-rank results and checks (a) the union equals the oracle's aggregate over all shards, (b) every group landed on
the rank that owns its Spark partition id (pmod(murmur3(k, 42), P) * world / P).

A second task checks the broadcast exchange ("nccl-bcast://", an all-gather-v): every rank holds a slice of a dimension table,
collects the whole of it over NVLink and joins its own shard against it.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_exchange.py
"""
import os
import sys
import time

import numpy as np
import pyarrow as pa
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from auron_b200 import proto as P  # noqa: E402
from auron_b200 import runtime  # noqa: E402


def shard(rank: int, n: int) -> pa.Table:
    rng = np.random.default_rng(1000 + rank)
    words = ["", "a", "bb", "ccc", "天地"]
    return pa.table({
        "k": pa.array(rng.integers(0, 50_000, n), type=pa.int64(), mask=rng.random(n) < 0.01),
        "s": pa.array([words[int(i)] for i in rng.integers(0, len(words), n)], mask=rng.random(n) < 0.05),
        "v": pa.array(rng.integers(-1000, 1000, n), type=pa.int64(), mask=rng.random(n) < 0.03),
    })


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ids = [runtime.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    runtime.nccl_init(ids[0], rank, world, local)

    n, nparts = int(os.environ.get("EXCHANGE_ROWS", 400_000)), 24
    t = shard(rank, n)
    src = P.ffi_reader(t.schema, "shard")
    aggs = lambda mode, args: [P.agg_expr("SUM", [args("v")], pa.int64()), P.agg_expr("COUNT", [args("v")], pa.int64()),
                               P.agg_expr("MAX", [args("v")], pa.int64())]
    partial = P.agg(src, [P.col("k"), P.col("s")], ["k", "s"], aggs("PARTIAL", P.col), ["sum", "cnt", "mx"], ["PARTIAL"] * 3)
    exch = P.shuffle_writer(partial, P.hash_repartition([P.col("k"), P.col("s")], nparts), "nccl://stage1", "")
    final = P.agg(exch, [P.col("k"), P.col("s")], ["k", "s"], aggs("FINAL", lambda c: P.lit(None, pa.null())), ["sum", "cnt", "mx"], ["FINAL"] * 3)
    td = P.task_definition(final, stage_id=2, partition_id=rank)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    with runtime.Task(td, {"shard": t.to_batches()}, device=local) as task:
        out = pa.Table.from_batches(list(task), schema=task.schema)
        metrics = task.metrics()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0

    gathered = [None] * world if rank == 0 else None
    dist.gather_object((rank, out.to_pydict(), {f"{op}.{name}": v for _, op, name, v in metrics if "exchange" in name or name == "data_size"}, dt), gathered, dst=0)
    shards = [None] * world if rank == 0 else None
    dist.gather_object(t.to_pydict(), shards, dst=0)
    ok = True
    if rank == 0:
        import oracle
        exp = {}
        for sh in shards:
            for k, s, v in zip(sh["k"], sh["s"], sh["v"]):
                a = exp.setdefault((k, s), [None, 0, None])
                if v is not None:
                    a[0] = (a[0] or 0) + v
                    a[1] += 1
                    a[2] = v if a[2] is None else max(a[2], v)
        got = {}
        for r, d, m, secs in gathered:
            keys = list(zip(d["k"], d["s"]))
            pid = oracle.partition_ids([pa.array(d["k"], type=pa.int64()), pa.array(d["s"], type=pa.string())], nparts) if keys else []
            for (k, s), sm, c, mx, p in zip(keys, d["sum"], d["cnt"], d["mx"], pid):
                assert (k, s) not in got, f"group {(k, s)} finished on two ranks"
                assert runtime.owner_of_partition(int(p), world, nparts) == r, f"group {(k, s)} (partition {p}) landed on rank {r}"
                got[(k, s)] = [sm, c, mx]
            print(f"[rank {r}] groups={len(keys)} step={secs * 1e3:.1f} ms {m}")
        ok = got == exp
        print("EXCHANGE_OK" if ok else f"EXCHANGE_MISMATCH got={len(got)} exp={len(exp)}")
    # ---- broadcast exchange: the build side of a join lives in slices on the ranks ("nccl-bcast://": all-gather-v over NVLink),
    # every rank joins its own shard against the whole of it
    dim_ids = np.arange(rank, 50_000, world, dtype=np.int64)            # rank r holds the ids congruent to r
    dim = pa.table({"id": pa.array(dim_ids), "label": pa.array([f"item-{int(i) % 97}" for i in dim_ids]),
                    "w": pa.array((dim_ids * 7 % 1000).astype(np.int64), mask=dim_ids % 11 == 0)})
    bcast = P.shuffle_writer(P.ffi_reader(dim.schema, "dim"), P.single_repartition(1), "nccl-bcast://dim", "")
    js = pa.schema(list(t.schema) + list(dim.schema))
    join = P.hash_join(js, P.ffi_reader(t.schema, "shard"), bcast, [(P.col("k"), P.col("id"))], "INNER", "RIGHT")
    agg = P.agg(join, [P.col("label")], ["label"], [P.agg_expr("COUNT", [P.col("k")], pa.int64()), P.agg_expr("SUM", [P.col("w")], pa.int64())], ["c", "sw"],
                ["PARTIAL"] * 2)
    with runtime.Task(P.task_definition(agg, stage_id=3, partition_id=rank), {"shard": t.to_batches(), "dim": dim.to_batches()}, device=local) as task:
        jout = pa.Table.from_batches(list(task), schema=task.schema)
    got_j = {l: (c, sw) for l, c, sw in zip(jout.column(0).to_pylist(), jout.column(1).to_pylist(), jout.column(2).to_pylist())}   # (partial aggregate: columns by position)
    exp_j = {}
    for k in t["k"].to_pylist():
        if k is None:
            continue
        a = exp_j.setdefault(f"item-{k % 97}", [0, None])
        a[0] += 1
        if k % 11 != 0:
            a[1] = (a[1] or 0) + k * 7 % 1000
    bok = got_j == {l: (c, sw) for l, (c, sw) in exp_j.items()}
    flags = [None] * world if rank == 0 else None
    dist.gather_object(bok, flags, dst=0)
    if rank == 0:
        print("BROADCAST_OK" if all(flags) else f"BROADCAST_MISMATCH {flags}")
        ok = ok and all(flags)
    runtime.nccl_finalize()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
