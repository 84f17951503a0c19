"""Parity of the fused ParquetScan -> Filter -> HashAggregate pass (k_fused.cu) -- checked three ways on the same files:
against Arrow C++ (pyarrow reader + group_by: the stand-in for the `parquet` crate + arrow kernels the reference delegates
to), against numpy reductions of the generated arrays, and against the engine's own operator-by-operator
path (AURON_DISABLE_FUSED_SCAN_AGG=1).  Shapes: pages of different columns that do not line up, PLAIN fallback pages next to
dictionary pages, all-NULL pages, NULL keys, v2 pages, several row groups / files / device batches, uncompressed / Snappy /
ZSTD, empty selections, a key range that widens from batch to batch."""
import os
import sys

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pyarrow.parquet as pq
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from auron_b200 import proto as P
from auron_b200 import runtime

pytestmark = pytest.mark.gpu


def _plan(paths, schema, key, preds, aggs, key_cast=pa.int64()):
    scan = P.parquet_scan(schema, [(p, os.path.getsize(p)) for p in paths], list(range(len(schema))))
    src = scan
    if preds:
        src = P.filter_(scan, [P.binary(op, P.col(c), P.lit(v, schema.field(c).type)) for c, op, v in preds])
    kexpr = P.try_cast(P.col(key), key_cast) if key_cast is not None else P.col(key)
    ae, names = [], []
    for fn, c in aggs:
        if fn == "COUNT*":
            ae.append(P.agg_expr("COUNT", [], pa.int64()))
        elif fn in ("MIN", "MAX"):
            ae.append(P.agg_expr(fn, [P.col(c)], schema.field(c).type))
        else:
            ae.append(P.agg_expr(fn, [P.col(c)], pa.int64()))
        names.append(f"{fn}_{c}")
    return P.task_definition(P.agg(src, [kexpr], [key], ae, names, ["PARTIAL"] * len(ae)))


def _run(td, fused=True, chunk_rows=None):
    env = {"AURON_DISABLE_FUSED_SCAN_AGG": None if fused else "1", "AURON_GPU_CHUNK_ROWS": None if chunk_rows is None else str(chunk_rows)}
    old = {k: os.environ.get(k) for k in env}
    try:
        for k, v in env.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        with runtime.Task(td) as task:
            out = pa.Table.from_batches(list(task), schema=task.schema)
            met = {(op, name): v for _, op, name, v in task.metrics()}
        return out, met
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _rows(t):
    cols = [c.to_pylist() for c in t.columns]
    return sorted(zip(*cols), key=lambda r: tuple((0, 0) if v is None else (1, v) for v in r))


def _expected(table, key, preds, aggs):
    """numpy / python restatement: NULL predicate => row dropped; SUM of no valid input => NULL; COUNT => 0; NULL key is a group"""
    n = table.num_rows
    keep = np.ones(n, dtype=bool)
    for c, op, v in preds:
        col = table[c].combine_chunks()
        valid = ~np.asarray(col.is_null().to_numpy(zero_copy_only=False))
        x = np.asarray(col.fill_null(0).to_numpy(zero_copy_only=False)).astype(np.int64)
        cmp = {"Eq": x == v, "Lt": x < v, "LtEq": x <= v, "Gt": x > v, "GtEq": x >= v}[op]
        keep &= valid & cmp
    kcol = table[key].combine_chunks()
    kvalid = ~np.asarray(kcol.is_null().to_numpy(zero_copy_only=False))
    kx = np.asarray(kcol.fill_null(0).to_numpy(zero_copy_only=False)).astype(np.int64)
    groups = {}
    idx = np.nonzero(keep)[0]
    argv = {}
    for fn, c in aggs:
        if c is not None and c not in argv:
            col = table[c].combine_chunks()
            argv[c] = (~np.asarray(col.is_null().to_numpy(zero_copy_only=False)), np.asarray(col.fill_null(0).to_numpy(zero_copy_only=False)).astype(np.int64))
    for i in idx:
        k = int(kx[i]) if kvalid[i] else None
        g = groups.setdefault(k, [None if fn in ("SUM", "MIN", "MAX") else 0 for fn, _ in aggs])
        for a, (fn, c) in enumerate(aggs):
            if fn == "COUNT*":
                g[a] += 1
                continue
            ok, x = argv[c]
            if not ok[i]:
                continue
            v = int(x[i])
            if fn == "COUNT":
                g[a] += 1
            elif fn == "SUM":
                g[a] = v if g[a] is None else g[a] + v
            elif fn == "MIN":
                g[a] = v if g[a] is None else min(g[a], v)
            elif fn == "MAX":
                g[a] = v if g[a] is None else max(g[a], v)
    return sorted([(k, *vals) for k, vals in groups.items()], key=lambda r: tuple((0, 0) if v is None else (1, v) for v in r))


def _check(paths, table, key, preds, aggs, chunk_rows=None, expect_fused=True, key_cast=pa.int64()):
    td = _plan(paths, table.schema, key, preds, aggs, key_cast)
    got, met = _run(td, fused=True, chunk_rows=chunk_rows)
    ref, met0 = _run(td, fused=False, chunk_rows=chunk_rows)
    fused_batches = met.get(("ParquetExec", "fused_batches"), 0)
    assert (fused_batches > 0) == expect_fused, met
    assert met0.get(("ParquetExec", "fused_batches"), 0) == 0
    assert got.schema == ref.schema
    exp = _expected(table, key, preds, [(fn, c) for fn, c in aggs])
    assert _rows(ref) == exp              # operator-by-operator path vs the numpy restatement
    assert _rows(got) == exp              # fused pass vs the numpy restatement
    if preds and expect_fused:            # FilterExec's output_rows is reported by the fused pass as well
        assert met.get(("FilterExec", "output_rows")) == met0.get(("FilterExec", "output_rows"))
    return got, met


def _write(path, table, **kw):
    kw.setdefault("compression", "SNAPPY")
    kw.setdefault("use_dictionary", True)
    pq.write_table(table, path, **kw)
    return path


def _store_sales(rng, n, n_items=5000, null_key=0.0, null_q=0.03, null_d=0.04):
    return pa.table({
        "item": pa.array(rng.integers(1, n_items + 1, n, dtype=np.int32), mask=(rng.random(n) < null_key) if null_key else None),
        "qty": pa.array(rng.integers(1, 101, n, dtype=np.int32), mask=rng.random(n) < null_q),
        "date": pa.array(rng.integers(2450816, 2452642, n, dtype=np.int32), mask=rng.random(n) < null_d),
    })


DATE_PREDS = [("date", "GtEq", 2451000), ("date", "Lt", 2452000)]
SUM_COUNT = [("SUM", "qty"), ("COUNT", "qty")]


def test_fused_config2_shape(tmp_path):
    rng = np.random.default_rng(1)
    t = _store_sales(rng, 300_000)
    p = _write(str(tmp_path / "a.parquet"), t, row_group_size=120_000)
    _check([p], t, "item", DATE_PREDS, SUM_COUNT)


def test_fused_ragged_pages_and_plain_fallback(tmp_path):
    # tiny data pages (different row counts per column), a dictionary that overflows into PLAIN pages for the key
    rng = np.random.default_rng(2)
    n = 200_000
    t = pa.table({
        "item": pa.array(rng.integers(1, 150_000, n, dtype=np.int32), mask=rng.random(n) < 0.01),
        "qty": pa.array(rng.integers(-50, 101, n, dtype=np.int32), mask=rng.random(n) < 0.2),
        "date": pa.array(rng.integers(2450816, 2452642, n, dtype=np.int32), mask=rng.random(n) < 0.04),
    })
    p = _write(str(tmp_path / "a.parquet"), t, row_group_size=90_000, data_page_size=3000, dictionary_pagesize_limit=40_000, write_batch_size=257)
    _check([p], t, "item", DATE_PREDS, SUM_COUNT + [("MIN", "qty"), ("MAX", "qty"), ("COUNT*", None)])


@pytest.mark.parametrize("codec,version", [("NONE", "1.0"), ("SNAPPY", "2.0"), ("ZSTD", "1.0")])
def test_fused_codecs_and_page_versions(tmp_path, codec, version):
    rng = np.random.default_rng(3)
    t = _store_sales(rng, 150_000, n_items=300, null_key=0.02)
    p = _write(str(tmp_path / "a.parquet"), t, compression=codec, data_page_version=version, row_group_size=64_000, data_page_size=20_000)
    _check([p], t, "item", DATE_PREDS, SUM_COUNT)


def test_fused_no_filter_and_not_null_columns(tmp_path):
    rng = np.random.default_rng(4)
    n = 100_000
    schema = pa.schema([pa.field("item", pa.int32(), nullable=False), pa.field("qty", pa.int32(), nullable=False), pa.field("date", pa.int32())])
    t = pa.table({"item": pa.array(rng.integers(1, 1000, n, dtype=np.int32)), "qty": pa.array(rng.integers(1, 101, n, dtype=np.int32)),
                  "date": pa.array(rng.integers(2450816, 2452642, n, dtype=np.int32), mask=rng.random(n) < 0.5)}, schema=schema)
    p = _write(str(tmp_path / "a.parquet"), t, row_group_size=33_333)
    _check([p], t, "item", [], SUM_COUNT + [("COUNT", "date")])
    _check([p], t, "item", [("qty", "Gt", 50)], [("SUM", "qty"), ("COUNT*", None)])     # predicate and argument on the same column
    _check([p], t, "item", [("item", "LtEq", 10)], [("MAX", "date")])                    # predicate on the key column


def test_fused_empty_selection_and_all_null_pages(tmp_path):
    rng = np.random.default_rng(5)
    n = 60_000
    qty = pa.array(rng.integers(1, 101, n, dtype=np.int32), mask=np.arange(n) < 45_000)   # long all-NULL stretch: whole pages without values
    t = pa.table({"item": pa.array(rng.integers(1, 50, n, dtype=np.int32)), "qty": qty,
                  "date": pa.array(rng.integers(2450816, 2452642, n, dtype=np.int32))})
    p = _write(str(tmp_path / "a.parquet"), t, data_page_size=8_000, data_page_version="2.0")
    _check([p], t, "item", DATE_PREDS, SUM_COUNT)
    got, _ = _check([p], t, "item", [("date", "Lt", 0)], SUM_COUNT)
    assert got.num_rows == 0


def test_fused_several_files_batches_and_growing_key_range(tmp_path):
    rng = np.random.default_rng(6)
    paths, parts = [], []
    for i, (lo, hi) in enumerate([(1000, 2000), (500, 1500), (1800, 9000), (1, 100)]):
        n = 70_000 + 1111 * i
        t = pa.table({"item": pa.array(rng.integers(lo, hi, n, dtype=np.int32), mask=rng.random(n) < 0.01),
                      "qty": pa.array(rng.integers(1, 101, n, dtype=np.int32), mask=rng.random(n) < 0.03),
                      "date": pa.array(rng.integers(2450816, 2452642, n, dtype=np.int32), mask=rng.random(n) < 0.04)})
        paths.append(_write(str(tmp_path / f"f{i}.parquet"), t, row_group_size=30_000))
        parts.append(t)
    full = pa.concat_tables(parts)
    got, met = _check(paths, full, "item", DATE_PREDS, SUM_COUNT, chunk_rows=65_000)
    assert met[("ParquetExec", "fused_batches")] >= 4          # one device batch per ~2 row groups
    _check(paths, full, "item", DATE_PREDS, SUM_COUNT)         # everything in one batch


def test_fused_falls_back_when_it_must(tmp_path):
    rng = np.random.default_rng(7)
    n = 50_000
    # int64 physical key column: the fused kernels take INT32 pages only -> regular path, same answer
    t = pa.table({"item": pa.array(rng.integers(1, 1000, n, dtype=np.int64)), "qty": pa.array(rng.integers(1, 101, n, dtype=np.int32)),
                  "date": pa.array(rng.integers(2450816, 2452642, n, dtype=np.int32))})
    p = _write(str(tmp_path / "a.parquet"), t)
    _check([p], t, "item", DATE_PREDS, SUM_COUNT, expect_fused=False, key_cast=None)
    # key range wider than the direct table: every batch falls back, the hash table takes over
    t2 = pa.table({"item": pa.array(rng.integers(-2**30, 2**30, n, dtype=np.int32)), "qty": pa.array(rng.integers(1, 101, n, dtype=np.int32)),
                   "date": pa.array(rng.integers(2450816, 2452642, n, dtype=np.int32))})
    p2 = _write(str(tmp_path / "b.parquet"), t2)
    _check([p2], t2, "item", DATE_PREDS, SUM_COUNT, expect_fused=False)


def test_fused_survives_wrong_key_statistics(tmp_path):
    # chunk statistics that do not cover the data (buggy writer): the fused pass notices the out-of-range key, its result is
    # discarded and the input is read again operator by operator -- same answer
    rng = np.random.default_rng(8)
    n = 80_000
    k = rng.integers(1000, 2000, n).astype(np.int32)
    k[:10] = [1000, 1999] * 5
    t = pa.table({"item": pa.array(k), "qty": pa.array(rng.integers(1, 101, n, dtype=np.int32)), "date": pa.array(rng.integers(2450816, 2452642, n, dtype=np.int32))})
    path = str(tmp_path / "stats.parquet")
    pq.write_table(t, path, compression="NONE", use_dictionary=False)
    raw = bytearray(open(path, "rb").read())
    footer_len = int.from_bytes(raw[-8:-4], "little")
    foot = len(raw) - 8 - footer_len
    patched = raw[:foot] + raw[foot:].replace((1999).to_bytes(4, "little"), (1500).to_bytes(4, "little"))
    assert patched != raw and len(patched) == len(raw)
    open(path, "wb").write(bytes(patched))
    assert pq.ParquetFile(path).metadata.row_group(0).column(0).statistics.max == 1500
    td = _plan([path], t.schema, "item", DATE_PREDS, SUM_COUNT)
    got, met = _run(td, fused=True)
    assert met.get(("ParquetExec", "restarted_unfused")) == 1
    assert _rows(got) == _expected(t, "item", DATE_PREDS, SUM_COUNT)


def test_scan_partition_columns_and_row_group_pruning(tmp_path):
    # Hive partition columns (FileScanConfig: projection indices past the file schema) + pruning_predicates
    rng = np.random.default_rng(9)
    paths, parts, pvals = [], [], []
    for i, (day, lo) in enumerate([(2451000, 0), (2451001, 100_000), (2451002, 200_000)]):
        n = 50_000
        t = pa.table({"item": pa.array(np.sort(rng.integers(lo, lo + 100_000, n, dtype=np.int32))), "qty": pa.array(rng.integers(1, 101, n, dtype=np.int32))})
        p = str(tmp_path / f"p{i}.parquet")
        pq.write_table(t, p, row_group_size=10_000)
        paths.append(p)
        pvals.append([day, f"store{i}"])
        parts.append(t.append_column("day", pa.array([day] * n, type=pa.int32())).append_column("store", pa.array([f"store{i}"] * n)))
    full = pa.concat_tables(parts)
    file_schema = pa.schema([("item", pa.int32()), ("qty", pa.int32())])
    part_schema = pa.schema([("day", pa.int32()), ("store", pa.string())])
    files = [(p, os.path.getsize(p)) for p in paths]
    # every column, partition columns included (indices 2, 3), in a shuffled order
    scan = P.parquet_scan(file_schema, files, [3, 0, 2, 1], partition_schema=part_schema, partition_values=pvals)
    with runtime.Task(P.task_definition(scan)) as task:
        out = pa.Table.from_batches(list(task), schema=task.schema)
    assert out.column_names == ["store", "item", "day", "qty"]
    assert out.select(["item", "qty", "day", "store"]).equals(full.cast(out.select(["item", "qty", "day", "store"]).schema))
    # filter on a partition column + aggregate; row groups that cannot match `item` are pruned by their statistics
    prune = [P.binary("GtEq", P.col("item"), P.lit(150_000, pa.int32())), P.binary("Lt", P.col("item"), P.lit(160_000, pa.int32()))]
    scan = P.parquet_scan(file_schema, files, [0, 1, 2], pruning_predicates=prune, partition_schema=part_schema, partition_values=pvals)
    flt = P.filter_(scan, prune + [P.binary("Eq", P.col("day"), P.lit(2451001, pa.int32()))])
    plan = P.agg(flt, [P.col("day")], ["day"], [P.agg_expr("SUM", [P.col("qty")], pa.int64()), P.agg_expr("COUNT", [P.col("item")], pa.int64())], ["s", "c"], ["PARTIAL"] * 2)
    with runtime.Task(P.task_definition(plan)) as task:
        got = pa.Table.from_batches(list(task), schema=task.schema)
        met = {(op, name): v for _, op, name, v in task.metrics()}
    item, qty = full["item"].to_numpy(), full["qty"].to_numpy()
    keep = (item >= 150_000) & (item < 160_000) & (full["day"].to_numpy() == 2451001)
    assert got.to_pydict() == {"day": [2451001], "": [int(qty[keep].sum())], "_": [int(keep.sum())]} or \
        (got.column(0).to_pylist(), got.column(1).to_pylist(), got.column(2).to_pylist()) == ([2451001], [int(qty[keep].sum())], [int(keep.sum())])
    assert met.get(("ParquetExec", "row_groups_pruned"), 0) >= 12, met      # 15 row groups, at most 2-3 can hold items in [150000, 160000)
    assert met[("ParquetExec", "output_rows")] <= 30_000


def test_fused_rle_runs_and_mixed_streams(tmp_path):
    # sorted / clustered columns: the dictionary-index streams are RLE runs and short literal runs, not the maximal bit-packed runs
    # of random data -- the scout's arithmetic checkpoints do not apply and it walks the run headers
    rng = np.random.default_rng(10)
    n = 150_000
    item = np.sort(rng.integers(1, 400, n)).astype(np.int32)                      # long RLE runs
    qty = np.repeat(rng.integers(1, 101, n // 10), 10).astype(np.int32)           # runs of exactly 10: RLE(8) + literal groups mixed
    date = rng.integers(2450816, 2452642, n, dtype=np.int32)
    date[: n // 2] = 2451500                                                      # half constant, half random
    t = pa.table({"item": pa.array(item, mask=rng.random(n) < 0.01), "qty": pa.array(qty, mask=rng.random(n) < 0.03), "date": pa.array(date, mask=rng.random(n) < 0.04)})
    p = _write(str(tmp_path / "a.parquet"), t, row_group_size=70_000, data_page_size=16_000)
    _check([p], t, "item", DATE_PREDS, SUM_COUNT + [("COUNT*", None)])
    _check([p], t, "qty", [("item", "Gt", 100)], [("SUM", "date"), ("MIN", "item"), ("MAX", "item")])


def test_fused_staged_tiles_match_the_tile_kernel(tmp_path, monkeypatch):
    # The TMA-staged kernel takes the tiles whose columns are single segments of regular index streams (large pages of
    # high-cardinality columns); page boundaries, NULLs at any density and PLAIN fallback pages leave a mix of staged and
    # unstaged tiles.  Both kernels must produce the same groups as the numpy restatement, with and without staging.
    rng = np.random.default_rng(41)
    n = 700_000
    t = pa.table({
        "item": pa.array(rng.integers(1, 60_000, n, dtype=np.int32), mask=rng.random(n) < 0.01),
        "qty": pa.array(rng.integers(-100, 101, n, dtype=np.int32), mask=rng.random(n) < 0.3),
        "date": pa.array(rng.integers(2450816, 2452642, n, dtype=np.int32), mask=rng.random(n) < 0.04),
        "dense": pa.array(rng.integers(0, 1 << 20, n, dtype=np.int32)),
    })
    p = _write(str(tmp_path / "a.parquet"), t, row_group_size=300_000, data_page_size=256 * 1024)
    aggs = SUM_COUNT + [("MIN", "qty"), ("COUNT*", None), ("MAX", "dense")]
    got, met = _check([p], t, "item", DATE_PREDS, aggs)
    staged = sum(v for (op, name), v in met.items() if name == "fused_staged_tiles")
    assert staged > 0.8 * (n // 1024), met
    monkeypatch.setenv("AURON_FUSED_NO_TMA", "1")
    got2, met2 = _check([p], t, "item", DATE_PREDS, aggs)
    assert sum(v for (op, name), v in met2.items() if name == "fused_staged_tiles") == 0
    assert _rows(got) == _rows(got2)
    # no predicate, a nullable key, PLAIN-only argument column (dictionary disabled for it)
    monkeypatch.delenv("AURON_FUSED_NO_TMA")
    p2 = _write(str(tmp_path / "b.parquet"), t, row_group_size=300_000, data_page_size=128 * 1024, use_dictionary=["item", "date", "qty"])
    got3, met3 = _check([p2], t, "item", [("date", "Gt", 2451500)], [("SUM", "dense"), ("COUNT", "qty")])
    assert sum(v for (op, name), v in met3.items() if name == "fused_staged_tiles") > 0, met3


def test_predicates_on_several_columns_run_operator_by_operator(tmp_path):
    # a conjunction over two columns is not fused (scan_parquet.cc can_fuse): the regular Filter -> HashAggregate path answers
    rng = np.random.default_rng(13)
    t = _store_sales(rng, 120_000)
    p = _write(str(tmp_path / "a.parquet"), t, row_group_size=50_000)
    _check([p], t, "item", DATE_PREDS + [("qty", "Gt", 40)], SUM_COUNT + [("COUNT*", None)], expect_fused=False)
