"""ParquetScanExec parity (row P1): the device decode of RLE / bit-packed hybrid levels and dictionary indices,
PLAIN values and byte arrays must reproduce what Arrow C++'s reader (the oracle for the third-party `parquet`
crate) returns for the same files."""
import decimal
import os

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pyarrow.parquet as pq
import pytest

import oracle
from auron_b200 import proto as P
from auron_b200 import runtime
from helpers import assert_same_rows, run

pytestmark = pytest.mark.gpu


def _table(n, seed, nulls=0.04):
    rng = np.random.default_rng(seed)

    def m():
        return rng.random(n) < nulls if nulls else None

    words = ["", "a", "store", "sales", "天地玄黄", "x" * 40, "b200"]
    return pa.table({
        "sk": pa.array(rng.integers(1, 2000, n), type=pa.int32(), mask=m()),           # dictionary friendly
        "big": pa.array(rng.integers(-2**62, 2**62, n), type=pa.int64(), mask=m()),     # high cardinality -> PLAIN fallback
        "q": pa.array(rng.integers(1, 100, n), type=pa.int32()),                        # no nulls
        "price": pa.array([None if x else decimal.Decimal(int(v)) / 100 for x, v in zip(rng.random(n) < nulls, rng.integers(0, 99999, n))],
                          type=pa.decimal128(7, 2)),
        "f": pa.array(rng.standard_normal(n).astype(np.float32), mask=m()),
        "d": pa.array(rng.standard_normal(n), mask=m()),
        "flag": pa.array(rng.random(n) < 0.3, mask=m()),
        "s": pa.array([words[int(i)] for i in rng.integers(0, len(words), n)], mask=m()),
        "u": pa.array([f"unique-{i}-{int(x)}" for i, x in enumerate(rng.integers(0, 10**9, n))], mask=m()),  # PLAIN strings
        "dt": pa.array(rng.integers(10000, 20000, n).astype(np.int32), type=pa.date32(), mask=m()),
    })


def _scan(path, schema, proj=None):
    plan = P.parquet_scan(schema, [(path, os.path.getsize(path))], proj if proj is not None else list(range(len(schema))))
    return run(plan, {})


@pytest.mark.parametrize("compression", ["NONE", "SNAPPY", "ZSTD"])
@pytest.mark.parametrize("version,dict_", [("1.0", True), ("2.0", True), ("1.0", False)])
def test_scan_matches_arrow_reader(tmp_path, compression, version, dict_):
    t = _table(50_000, seed=7)
    path = str(tmp_path / "t.parquet")
    pq.write_table(t, path, compression=compression, use_dictionary=dict_, data_page_version=version, row_group_size=17_001, data_page_size=8 * 1024,
                   store_decimal_as_integer=True)
    got = _scan(path, t.schema)
    exp = pq.read_table(path)
    assert got.num_rows == exp.num_rows
    for name in t.column_names:
        assert got[name].to_pylist() == exp[name].to_pylist(), name       # scans preserve file order


def test_scan_flba_decimal_required_columns_and_projection(tmp_path):
    t = _table(20_000, seed=9, nulls=0)
    path = str(tmp_path / "r.parquet")
    # default pyarrow: decimal as FIXED_LEN_BYTE_ARRAY; non-nullable schema -> REQUIRED columns (max_def = 0)
    schema = pa.schema([pa.field(f.name, f.type, nullable=False) for f in t.schema])
    pq.write_table(t.cast(schema), path, compression="NONE", row_group_size=6_000)
    got = _scan(path, t.schema, proj=[3, 0, 7])
    exp = pq.read_table(path, columns=["price", "sk", "s"])
    assert got.schema.names == ["price", "sk", "s"]
    for name in got.schema.names:
        assert got[name].to_pylist() == exp[name].to_pylist(), name


def test_scan_schema_adaptation(tmp_path):
    # scan/mod.rs:56-160: case-insensitive match, missing column -> NULL, INT32 -> int64 / decimal128 widening
    t = pa.table({"A": pa.array([1, None, 3], type=pa.int32()), "p": pa.array([150, 275, None], type=pa.int32())})
    path = str(tmp_path / "a.parquet")
    pq.write_table(t, path)
    want = pa.schema([("a", pa.int64()), ("p", pa.decimal128(7, 2)), ("missing", pa.string())])
    got = _scan(path, want)
    assert got["a"].to_pylist() == [1, None, 3]
    assert got["p"].to_pylist() == [decimal.Decimal("1.50"), decimal.Decimal("2.75"), None]
    assert got["missing"].to_pylist() == [None, None, None]


def test_scan_file_ranges_partition_row_groups(tmp_path):
    t = _table(30_000, seed=11).select(["sk", "q"])
    path = str(tmp_path / "g.parquet")
    pq.write_table(t, path, row_group_size=5_000)
    size = os.path.getsize(path)
    md = pq.ParquetFile(path).metadata
    starts = [md.row_group(i).column(0).dictionary_page_offset or md.row_group(i).column(0).data_page_offset for i in range(md.num_row_groups)]
    mid = starts[3]
    a = run(P.parquet_scan(t.schema, [(path, size)], [0, 1], ranges=[(0, mid)]), {})
    b = run(P.parquet_scan(t.schema, [(path, size)], [0, 1], ranges=[(mid, size)]), {})
    assert a.num_rows == 15_000 and b.num_rows == 15_000
    assert pa.concat_tables([a, b])["sk"].to_pylist() == t["sk"].to_pylist()


def test_scan_filter_aggregate_config2_small(tmp_path):
    # BASELINE config 2 at test size: ParquetScan -> Filter -> HashAggregate(GROUP BY int64, SUM/COUNT), host file and HBM-resident file
    rng = np.random.default_rng(13)
    n = 400_000
    t = pa.table({"ss_item_sk": pa.array(rng.integers(1, 20_000, n), type=pa.int32()),
                  "ss_quantity": pa.array(rng.integers(1, 101, n), type=pa.int32(), mask=rng.random(n) < 0.03),
                  "ss_sold_date_sk": pa.array(rng.integers(2450816, 2452642, n), type=pa.int32(), mask=rng.random(n) < 0.04)})
    path = str(tmp_path / "store_sales.parquet")
    pq.write_table(t, path, compression="NONE", row_group_size=150_000, data_page_size=64 * 1024)

    def plan(p):
        scan = P.parquet_scan(t.schema, [(p, os.path.getsize(path))], [0, 1, 2])
        flt = P.filter_(scan, [P.binary("GtEq", P.col("ss_sold_date_sk"), P.lit(2451000, pa.int32())),
                               P.binary("Lt", P.col("ss_sold_date_sk"), P.lit(2452000, pa.int32()))])
        return P.agg(flt, [P.try_cast(P.col("ss_item_sk"), pa.int64())], ["k"],
                     [P.agg_expr("SUM", [P.col("ss_quantity")], pa.int64()), P.agg_expr("COUNT", [P.col("ss_quantity")], pa.int64())],
                     ["s", "c"], ["PARTIAL", "PARTIAL"])

    d = t["ss_sold_date_sk"].combine_chunks()
    pred = np.asarray(pc.and_kleene(pc.greater_equal(d, 2451000), pc.less(d, 2452000)).fill_null(False))
    exp = oracle.agg_sum_count_i64(t["ss_item_sk"].combine_chunks().cast(pa.int64()), t["ss_quantity"].combine_chunks().cast(pa.int64()), pred)
    got = run(plan(path), {})
    assert_same_rows(got, exp)
    # same query with the file image resident in HBM
    with open(path, "rb") as f:
        data = f.read()
    runtime.put_device_file("hbm://store_sales", data)
    try:
        got2 = run(plan("hbm://store_sales"), {})
    finally:
        runtime.drop_device_file("hbm://store_sales")
    assert_same_rows(got2, exp)
    # and with the file image handed over as a (pinned) host buffer, small device chunks so the prefetch pipeline runs
    import torch

    buf = torch.frombuffer(bytearray(data), dtype=torch.uint8).pin_memory()
    runtime.put_host_file("pinned://store_sales", buf)
    old = os.environ.get("AURON_GPU_CHUNK_ROWS")
    os.environ["AURON_GPU_CHUNK_ROWS"] = "100000"
    try:
        got3 = run(plan("pinned://store_sales"), {})
    finally:
        runtime.drop_host_file("pinned://store_sales")
        if old is None:
            del os.environ["AURON_GPU_CHUNK_ROWS"]
        else:
            os.environ["AURON_GPU_CHUNK_ROWS"] = old
    assert_same_rows(got3, exp)


@pytest.mark.parametrize("version", ["1.0", "2.0"])
@pytest.mark.parametrize("dict_", [True, False])
def test_scan_snappy_pages_decompressed_on_device(tmp_path, version, dict_, monkeypatch):
    # compressible fixed-width columns: long runs, short periodic patterns (overlapping back references), sorted keys,
    # next to an incompressible column (a single literal per 64 KB block); v1 pages carry their level length inside the body
    rng = np.random.default_rng(5)
    n = 200_003
    t = pa.table({
        "runs": pa.array(np.repeat(rng.integers(0, 50, n // 1000 + 1), 1000)[:n].astype(np.int32), mask=rng.random(n) < 0.02),
        "period": pa.array((np.arange(n) % 7).astype(np.int64) * 1_000_003),
        "sorted": pa.array(np.sort(rng.integers(0, 10**6, n)).astype(np.int64), mask=rng.random(n) < 0.3),
        "noise": pa.array(rng.integers(-2**31, 2**31 - 1, n).astype(np.int32), mask=rng.random(n) < 0.05),
        "f": pa.array(np.round(rng.standard_normal(n), 1)),
        "flag": pa.array((np.arange(n) // 100) % 2 == 0, mask=rng.random(n) < 0.01),
        "allnull": pa.array([None] * n, type=pa.int32()),
    })
    path = str(tmp_path / "c.parquet")
    pq.write_table(t, path, compression="SNAPPY", use_dictionary=dict_, data_page_version=version, row_group_size=70_000, data_page_size=32 * 1024)
    exp = pq.read_table(path)
    got = _scan(path, t.schema)
    monkeypatch.setenv("AURON_HOST_SNAPPY", "1")
    got_host = _scan(path, t.schema)
    for name in t.column_names:
        assert got[name].to_pylist() == exp[name].to_pylist(), name
        assert got_host[name].to_pylist() == exp[name].to_pylist(), name


@pytest.mark.parametrize("version", ["1.0", "2.0"])
def test_scan_snappy_large_pages_split_into_head_and_stored_pieces(tmp_path, version):
    # 1 MB pages whose bodies are a few back references (the compressed level bytes) followed by a chain of 64 KB literals: the
    # host hands the literals out as stored-copy jobs and the elements before them as a Snappy job of their own; a body whose
    # back references come last ("tail_runs") and one with a dictionary ("keys") take the other routes
    rng = np.random.default_rng(17)
    n = 900_001
    half = n // 2
    t = pa.table({
        "noise": pa.array(rng.integers(-2**31, 2**31 - 1, n).astype(np.int32), mask=rng.random(n) < 0.05),
        "noise_nn": pa.array(rng.integers(-2**62, 2**62, n).astype(np.int64)),
        "tail_runs": pa.array(np.concatenate([rng.integers(-2**31, 2**31 - 1, half), np.repeat(rng.integers(0, 9, (n - half) // 500 + 1), 500)[:n - half]]).astype(np.int32),
                              mask=rng.random(n) < 0.02),
        "keys": pa.array(rng.integers(1, 150_000, n).astype(np.int32), mask=rng.random(n) < 0.03),
    })
    path = str(tmp_path / "big.parquet")
    pq.write_table(t, path, compression="SNAPPY", use_dictionary=["keys"], data_page_version=version, row_group_size=600_000)
    exp = pq.read_table(path)
    got = _scan(path, t.schema)
    for name in t.column_names:
        assert got[name].to_numpy(zero_copy_only=False).tobytes() == exp[name].to_numpy(zero_copy_only=False).tobytes(), name
        assert got[name].null_count == exp[name].null_count, name


@pytest.mark.parametrize("version", ["1.0", "2.0"])
@pytest.mark.parametrize("dict_", [True, False])
def test_scan_snappy_string_pages_decompressed_on_device(tmp_path, version, dict_, monkeypatch):
    # string columns: dictionary pages (PLAIN byte arrays) + RLE index pages, PLAIN data pages (v2; v1 when the column is
    # required), all decompressed by the GPU; nullable v1 PLAIN pages are the one case left to the host (their non-null count
    # sits inside the compressed body) and share a column chunk with device-decompressed dictionary pages when the writer
    # falls back from dictionary to PLAIN mid-chunk ("wide": the dictionary page size limit is hit)
    rng = np.random.default_rng(12)
    n = 120_001
    brands = [f"brand #{i:03d} " + "x" * (i % 17) for i in range(300)]
    wide = [f"{int(v):09d}-" + "payload" * 6 for v in rng.integers(0, 40_000, n)]
    schema = pa.schema([pa.field("brand", pa.string()), pa.field("req", pa.string(), nullable=False), pa.field("wide", pa.string()),
                        pa.field("bin", pa.binary()), pa.field("k", pa.int32())])
    t = pa.table({"brand": pa.array([brands[int(i)] for i in rng.integers(0, 300, n)], mask=rng.random(n) < 0.04),
                  "req": pa.array([brands[int(i) % 40] for i in rng.integers(0, 10**6, n)]),
                  "wide": pa.array(wide, mask=rng.random(n) < 0.1),
                  "bin": pa.array([bytes([int(i) % 251]) * (int(i) % 9) for i in rng.integers(0, 10**6, n)], type=pa.binary(), mask=rng.random(n) < 0.02),
                  "k": pa.array(rng.integers(0, 1000, n), type=pa.int32())}, schema=schema)
    path = str(tmp_path / "s.parquet")
    pq.write_table(t, path, compression="SNAPPY", use_dictionary=dict_, data_page_version=version, row_group_size=50_000, data_page_size=16 * 1024,
                   dictionary_pagesize_limit=64 * 1024)
    exp = pq.read_table(path)
    got = _scan(path, t.schema)
    monkeypatch.setenv("AURON_HOST_SNAPPY", "1")
    got_host = _scan(path, t.schema)
    for name in t.column_names:
        assert got[name].to_pylist() == exp[name].to_pylist(), name
        assert got_host[name].to_pylist() == exp[name].to_pylist(), name


def test_scan_reports_corrupt_snappy_page(tmp_path):
    n = 20_000
    t = pa.table({"a": pa.array(np.arange(n, dtype=np.int64) % 13)})
    path = str(tmp_path / "bad.parquet")
    pq.write_table(t, path, compression="SNAPPY", use_dictionary=False)
    md = pq.ParquetFile(path).metadata.row_group(0).column(0)
    raw = bytearray(open(path, "rb").read())
    off = md.data_page_offset + 40          # inside the first page body: turn literals into impossible back references
    for i in range(off, off + 64):
        raw[i] = 0xFF
    open(path, "wb").write(bytes(raw))
    with pytest.raises(runtime.AuronError):
        _scan(path, t.schema)


@pytest.mark.parametrize("lie", ["key_max", "value_max"])
def test_aggregate_survives_wrong_column_statistics(tmp_path, lie):
    # the scan hands the chunk statistics (min / max) to the aggregate's direct-address path; a file whose statistics do not
    # cover its data (buggy writer) must still aggregate correctly (the kernel flags the out-of-range key, the hash table takes over)
    rng = np.random.default_rng(3)
    n = 100_000
    k = rng.integers(1000, 2000, n).astype(np.int32)
    k[:10] = [1000, 1999] * 5                                  # make sure both bounds occur
    v = rng.integers(0, 100, n)
    v[:4] = [0, 99, 99, 0]
    t = pa.table({"k": pa.array(k), "v": pa.array(v, type=pa.int64())})
    path = str(tmp_path / "stats.parquet")
    pq.write_table(t, path, compression="NONE", use_dictionary=False)
    raw = bytearray(open(path, "rb").read())
    md = pq.ParquetFile(path).metadata
    footer_len = int.from_bytes(raw[-8:-4], "little")
    foot = len(raw) - 8 - footer_len
    if lie == "key_max":
        patched = raw[:foot] + raw[foot:].replace((1999).to_bytes(4, "little"), (1500).to_bytes(4, "little"))
    else:
        patched = raw[:foot] + raw[foot:].replace((99).to_bytes(8, "little"), (50).to_bytes(8, "little"))
    assert patched != raw and len(patched) == len(raw)
    open(path, "wb").write(bytes(patched))
    st = pq.ParquetFile(path).metadata.row_group(0)
    assert (st.column(0).statistics.max, st.column(1).statistics.max) == ((1500, 99) if lie == "key_max" else (1999, 50))
    scan = P.parquet_scan(t.schema, [(path, os.path.getsize(path))], [0, 1])
    plan = P.agg(scan, [P.col("k")], ["k"], [P.agg_expr("SUM", [P.col("v")], pa.int64()), P.agg_expr("COUNT", [P.col("v")], pa.int64())], ["s", "c"], ["PARTIAL"] * 2)
    got = run(plan, {})
    exp = oracle.agg_sum_count_i64(t["k"].combine_chunks().cast(pa.int64()), t["v"].combine_chunks())
    got = pa.table({"k": got.column(0).cast(pa.int64()), "s": got.column(1), "c": got.column(2)})
    assert_same_rows(got, exp)


@pytest.mark.parametrize("compression", ["NONE", "SNAPPY", "ZSTD"])
@pytest.mark.parametrize("version", ["1.0", "2.0"])
def test_scan_delta_encodings(tmp_path, compression, version):
    # DELTA_BINARY_PACKED INT32 / INT64 pages are transcribed to PLAIN on the device (one warp per page walks the blocks);
    # DELTA_LENGTH_BYTE_ARRAY / DELTA_BYTE_ARRAY string pages are rewritten as PLAIN by the host walker.  Sorted keys (tiny deltas,
    # zero-width miniblocks), noise (wide miniblocks, wrap-around deltas), NULLs, page counts that are not multiples of a block.
    rng = np.random.default_rng(23)
    n = 150_011
    words = ["", "a", "prefix", "prefix-shared", "prefix-shared-longer", "天地玄黄", "z" * 70]
    t = pa.table({
        "sorted32": pa.array(np.sort(rng.integers(-2**31, 2**31 - 1, n)).astype(np.int32), mask=rng.random(n) < 0.03),
        "noise32": pa.array(rng.integers(-2**31, 2**31 - 1, n).astype(np.int32)),
        "const64": pa.array(np.full(n, 7_000_000_000, dtype=np.int64), mask=rng.random(n) < 0.5),
        "noise64": pa.array(rng.integers(-2**63, 2**63 - 1, n, dtype=np.int64), mask=rng.random(n) < 0.02),
        "ts": pa.array(np.cumsum(rng.integers(0, 1000, n)).astype(np.int64), type=pa.timestamp("us")),
        "dl": pa.array([words[int(i)] + str(int(i) * 37 % 11) for i in rng.integers(0, len(words), n)], mask=rng.random(n) < 0.05),
        "db": pa.array(sorted(f"key-{int(x):09d}" for x in rng.integers(0, 10**9, n))),
        "allnull": pa.array([None] * n, type=pa.int64()),
    })
    enc = {"sorted32": "DELTA_BINARY_PACKED", "noise32": "DELTA_BINARY_PACKED", "const64": "DELTA_BINARY_PACKED", "noise64": "DELTA_BINARY_PACKED",
           "ts": "DELTA_BINARY_PACKED", "dl": "DELTA_LENGTH_BYTE_ARRAY", "db": "DELTA_BYTE_ARRAY", "allnull": "DELTA_BINARY_PACKED"}
    path = str(tmp_path / "delta.parquet")
    pq.write_table(t, path, compression=compression, use_dictionary=False, column_encoding=enc, data_page_version=version, row_group_size=70_001,
                   data_page_size=64 * 1024)
    md = pq.ParquetFile(path).metadata.row_group(0)
    assert "DELTA_BINARY_PACKED" in md.column(0).encodings and "DELTA_BYTE_ARRAY" in md.column(6).encodings
    exp = pq.read_table(path)
    got = _scan(path, t.schema)
    assert got.num_rows == exp.num_rows
    for name in t.column_names:
        assert got[name].to_pylist() == exp[name].to_pylist(), name


@pytest.mark.parametrize("dict_", [True, False])
def test_scan_int96_timestamps(tmp_path, dict_):
    # Spark's default timestamp encoding in Parquet is INT96 (nanoseconds of the day + Julian day); the scan converts it to the
    # unit of the table schema's timestamp column (parquet_exec.rs:192 coerce_int96 + AuronSchemaAdapter)
    rng = np.random.default_rng(12)
    n = 30_000
    us = rng.integers(-2_000_000_000_000_000, 4_000_000_000_000_000, n)          # 1906 .. 2096, microseconds
    if dict_:
        us = us[rng.integers(0, 500, n)]                                        # few distinct values -> dictionary pages
    t = pa.table({"ts": pa.array(us, type=pa.timestamp("us"), mask=rng.random(n) < 0.05), "k": pa.array(np.arange(n), type=pa.int32())})
    path = str(tmp_path / "int96.parquet")
    pq.write_table(t, path, use_deprecated_int96_timestamps=True, use_dictionary=dict_, compression="SNAPPY", data_page_size=20_000)
    assert pq.ParquetFile(path).schema.column(0).physical_type == "INT96"
    for unit in ("us", "ms"):
        schema = pa.schema([("ts", pa.timestamp(unit)), ("k", pa.int32())])
        got = _scan(path, schema)
        exp = us // (1 if unit == "us" else 1000)
        g = got["ts"].cast(pa.int64()).to_pylist()
        e = t["ts"].to_pylist()
        assert [None if v is None else int(x) for v, x in zip(e, exp)] == g
        assert got["k"].to_pylist() == list(range(n))
