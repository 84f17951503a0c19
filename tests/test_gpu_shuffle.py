"""ShuffleWriterExec parity (rows S2-S6): the .data/.index files written by the device path must be readable by
the reference's reader contract (IpcCompressionReader + read_batch, datafusion-ext-commons/src/io/
ipc_compression.rs:115-176, batch_serde.rs:81-101) -- restated by the oracle -- and every row must land in the
partition Spark's murmur3 partitioner assigns it (shuffle/mod.rs:163-188).  Row order inside a partition is
unspecified in the reference (rdx_sort.rs:55-73), so partitions are compared as multisets."""
import decimal
import os
import struct

import numpy as np
import pyarrow as pa
import pytest

import oracle
from auron_b200 import proto as P
from auron_b200 import runtime
from helpers import assert_same_rows, batches, run

pytestmark = pytest.mark.gpu


def read_shuffle_files(data_file, index_file, schema, codec="lz4"):
    """-> list of pa.Table, one per partition (reader side of the format)"""
    idx = open(index_file, "rb").read()
    offsets = list(struct.unpack(f"<{len(idx) // 8}q", idx))
    data = open(data_file, "rb").read()
    assert offsets[0] == 0 and offsets[-1] == len(data)
    parts = []
    for p in range(len(offsets) - 1):
        seg = data[offsets[p]:offsets[p + 1]]
        pos, rows = 0, []
        while pos < len(seg):                                             # block := u32_le len | codec stream
            (blen,) = struct.unpack_from("<I", seg, pos)
            pos += 4
            raw = pa.CompressedInputStream(pa.BufferReader(seg[pos:pos + blen]), codec).read()
            pos += blen
            bpos = 0
            while bpos < len(raw):                                        # payload := batch*
                b, bpos = oracle.serde_read_batch(raw, schema, bpos)
                rows.append(b)
        parts.append(pa.Table.from_batches(rows, schema=schema) if rows else schema.empty_table())
    return parts, offsets


def _table(n, seed):
    rng = np.random.default_rng(seed)
    words = ["", "a", "shuffle", "天地", "zz" * 20]
    return pa.table({
        "k": pa.array(rng.integers(0, 10_000, n), type=pa.int32(), mask=rng.random(n) < 0.02),
        "t": pa.array(rng.integers(-2**40, 2**40, n), type=pa.int64()),
        "p": pa.array([None if x else decimal.Decimal(int(v)) / 100 for x, v in zip(rng.random(n) < 0.03, rng.integers(0, 99999, n))], type=pa.decimal128(7, 2)),
        "s": pa.array([words[int(i)] for i in rng.integers(0, len(words), n)], mask=rng.random(n) < 0.05),
        "b": pa.array(rng.random(n) < 0.5, mask=rng.random(n) < 0.05),
        "f": pa.array(rng.standard_normal(n)),
    })


@pytest.mark.parametrize("codec", ["lz4", "zstd", "lz4-host"])
@pytest.mark.parametrize("n,nparts,chunk_rows", [(1000, 4, None), (100_000, 200, None), (100_000, 13, 30_000), (200_000, 3, None)])
def test_hash_partition_shuffle_write(tmp_path, codec, n, nparts, chunk_rows, monkeypatch):
    # "lz4" frames are produced on the GPU (64 KB independent blocks; the 3-partition case has ~40 blocks per stream),
    # "lz4-host" by liblz4 on the host cores, "zstd" by libzstd on the host cores
    if codec == "lz4-host":
        monkeypatch.setenv("AURON_HOST_LZ4", "1")
        codec = "lz4"
    t = _table(n, seed=n + nparts)
    data, index = str(tmp_path / "shuffle.data"), str(tmp_path / "shuffle.index")
    plan = P.shuffle_writer(P.ffi_reader(t.schema, "t"), P.hash_repartition([P.col("k"), P.col("s")], nparts), data, index)
    os.environ["AURON_IO_COMPRESSION_CODEC"] = codec
    if chunk_rows:
        os.environ["AURON_GPU_CHUNK_ROWS"] = str(chunk_rows)
    try:
        out = run(plan, {"t": t}, chunk=chunk_rows)
    finally:
        os.environ.pop("AURON_IO_COMPRESSION_CODEC", None)
        os.environ.pop("AURON_GPU_CHUNK_ROWS", None)
    assert out.num_rows == 0                                              # the writer's output stream is empty
    parts, offsets = read_shuffle_files(data, index, t.schema, codec)
    assert len(parts) == nparts and len(offsets) == nparts + 1           # N+1 offsets, first = 0 (buffered_data.rs:146,153)
    pid = oracle.partition_ids([t["k"].combine_chunks(), t["s"].combine_chunks()], nparts)
    total = 0
    for p in range(nparts):
        exp = t.filter(pa.array(pid == p))
        assert_same_rows(parts[p], exp)
        total += parts[p].num_rows
        if exp.num_rows == 0:
            assert offsets[p] == offsets[p + 1]                           # empty partitions repeat the previous offset
    assert total == n


def test_single_partition_and_empty_input(tmp_path):
    t = _table(5000, seed=3)
    data, index = str(tmp_path / "s.data"), str(tmp_path / "s.index")
    run(P.shuffle_writer(P.ffi_reader(t.schema, "t"), P.single_repartition(1), data, index), {"t": t})
    parts, offsets = read_shuffle_files(data, index, t.schema)
    assert len(offsets) == 2 and offsets[1] == os.path.getsize(data)      # single_repartitioner.rs:73-96
    assert_same_rows(parts[0], t)
    run(P.shuffle_writer(P.ffi_reader(t.schema, "t"), P.hash_repartition([P.col("k")], 8), data, index), {"t": t.slice(0, 0)})
    assert os.path.getsize(data) == 0
    assert struct.unpack("<9q", open(index, "rb").read()) == (0,) * 9     # zero rows => all-zero index (buffered_data.rs:124-126)


@pytest.mark.parametrize("partition_id,chunk_rows", [(0, None), (7, 20_000)])
def test_round_robin_shuffle_write(tmp_path, partition_id, chunk_rows):
    # evaluate_robin_partition_ids (shuffle/mod.rs:190-202) with the start offset of buffered_data.rs:291-312:
    # row i of the task goes to (partition_id * 1000193 + i) % N
    n, nparts = 50_001, 7
    t = _table(n, seed=99)
    data, index = str(tmp_path / "rr.data"), str(tmp_path / "rr.index")
    plan = P.shuffle_writer(P.ffi_reader(t.schema, "t"), P.round_robin_repartition(nparts), data, index)
    td = P.task_definition(plan, stage_id=1, partition_id=partition_id, task_id=3)
    if chunk_rows:
        os.environ["AURON_GPU_CHUNK_ROWS"] = str(chunk_rows)
    try:
        runtime.run_task(td, {"t": batches(t, chunk_rows)})
    finally:
        os.environ.pop("AURON_GPU_CHUNK_ROWS", None)
    parts, offsets = read_shuffle_files(data, index, t.schema)
    pid = (partition_id * 1000193 + np.arange(n)) % nparts
    for p in range(nparts):
        assert_same_rows(parts[p], t.filter(pa.array(pid == p)))


@pytest.mark.parametrize("case", ["int_asc", "int_desc_nulls_last", "string_then_int"])
def test_range_partition_shuffle_write(tmp_path, case):
    # evaluate_range_partition_ids / get_partition (shuffle/mod.rs:204-262): partition = number of bounds that sort strictly
    # before the row's key (a key equal to a bound stays in the lower partition); bounds arrive as List ScalarValues
    n = 40_000
    rng = np.random.default_rng(17)
    words = ["", "a", "ab", "b", "zz", "天"]
    t = pa.table({"k": pa.array(rng.integers(-500, 500, n), type=pa.int32(), mask=rng.random(n) < 0.05),
                  "s": pa.array([words[int(i)] for i in rng.integers(0, len(words), n)], mask=rng.random(n) < 0.05),
                  "row": pa.array(np.arange(n), type=pa.int64())})
    if case == "int_asc":
        keys, bounds = [("k", True, True)], [([-300, -1, 0, 0, 250], pa.int32())]
    elif case == "int_desc_nulls_last":
        keys, bounds = [("k", False, False)], [([400, 100, -100, None], pa.int32())]
    else:
        keys, bounds = [("s", True, True), ("k", True, True)], [(["a", "b", "b"], pa.string()), ([0, -10, 10], pa.int32())]
    nparts = len(bounds[0][0]) + 1
    data, index = str(tmp_path / "r.data"), str(tmp_path / "r.index")
    plan = P.shuffle_writer(P.ffi_reader(t.schema, "t"),
                            P.range_repartition([P.sort_expr(P.col(c), asc, nf) for c, asc, nf in keys], nparts, bounds), data, index)
    run(plan, {"t": t}, chunk=15_000)
    parts, offsets = read_shuffle_files(data, index, t.schema)

    def sort_key(vals):          # tuple comparable in the requested order: (null rank, value) per column
        out = []
        for v, (_, asc, nf) in zip(vals, keys):
            if v is None:
                out.append((0 if nf else 2, 0))
            else:
                x = v.encode() if isinstance(v, str) else v
                out.append((1, x if asc else _Neg(x)))
        return tuple(out)

    class _Neg:                  # reverses the order of the wrapped value
        def __init__(self, x):
            self.x = x

        def __lt__(self, o):
            return o.x < self.x

        def __eq__(self, o):
            return o.x == self.x

        def __gt__(self, o):
            return o.x > self.x

    bkeys = [sort_key([b[0][i] for b in bounds]) for i in range(nparts - 1)]
    cols = [t[c].to_pylist() for c, _, _ in keys]
    exp_pid = []
    for vals in zip(*cols):
        kk = sort_key(vals)
        exp_pid.append(sum(1 for b in bkeys if b < kk))
    exp_pid = np.array(exp_pid)
    for p in range(nparts):
        assert sorted(parts[p]["row"].to_pylist()) == np.nonzero(exp_pid == p)[0].tolist(), p


def _reference_style_segment(t, codec, batch_rows):
    """A shuffle segment as the reference writes it: batches of `batch_rows` rows in the compacted format (oracle port of
    write_batch), grouped into `u32 len | codec stream` blocks of a few batches each (IpcCompressionWriter)."""
    out = b""
    bs = t.to_batches(max_chunksize=batch_rows)
    for i in range(0, len(bs), 3):
        payload = b"".join(oracle.serde_write_batch(b) for b in bs[i:i + 3])
        sink = pa.BufferOutputStream()
        with pa.CompressedOutputStream(sink, codec) as z:
            z.write(payload)
        comp = sink.getvalue().to_pybytes()
        out += struct.pack("<I", len(comp)) + comp
    return out


@pytest.mark.parametrize("codec", ["lz4", "zstd"])
def test_ipc_reader_reads_reference_style_blocks(tmp_path, codec):
    # IpcReaderExec (ipc_reader_exec.rs:166-275): file segments and in-memory buffers, many small batches per block,
    # written by the oracle's port of the reference writer and compressed by Arrow C++'s LZ4-frame / ZSTD codecs
    t = _table(30_000, seed=5)
    seg_a = _reference_style_segment(t.slice(0, 12_000), codec, 1000)
    seg_b = _reference_style_segment(t.slice(12_000, 10_000), codec, 777)
    seg_c = _reference_style_segment(t.slice(22_000), codec, 5000)
    path = str(tmp_path / "seg.data")
    with open(path, "wb") as f:
        f.write(b"\x00" * 13 + seg_a + seg_b)                      # segments live at arbitrary file offsets
    blocks = [(path, 13, len(seg_a)), (path, 13 + len(seg_a), len(seg_b)), seg_c, (path, 13, 0)]
    td = P.task_definition(P.ipc_reader(t.schema, "shuffle_in"))
    got = runtime.run_task(td, shuffle_blocks={"shuffle_in": blocks})
    assert got.schema.names == t.schema.names
    for name in t.column_names:
        assert got[name].to_pylist() == t[name].to_pylist(), name       # block order = row order


@pytest.mark.parametrize("codec", ["lz4", "zstd"])
def test_ipc_reader_batches_straddling_blocks(codec):
    # the reference's reader chains blocks into one byte stream (ipc_compression.rs:131-170), so a batch may in principle end in
    # the next block: cut the payload at arbitrary byte positions
    t = _table(20_000, seed=8)
    payload = b"".join(oracle.serde_write_batch(b) for b in t.to_batches(max_chunksize=3000))
    seg = b""
    for o in range(0, len(payload), 10_007):
        sink = pa.BufferOutputStream()
        with pa.CompressedOutputStream(sink, codec) as z:
            z.write(payload[o:o + 10_007])
        comp = sink.getvalue().to_pybytes()
        seg += struct.pack("<I", len(comp)) + comp
    got = runtime.run_task(P.task_definition(P.ipc_reader(t.schema, "in")), shuffle_blocks={"in": [seg]})
    for name in t.column_names:
        assert got[name].to_pylist() == t[name].to_pylist(), name


@pytest.mark.parametrize("codec", ["lz4", "zstd"])
def test_shuffle_write_then_read_two_stage_aggregate(tmp_path, codec, monkeypatch):
    # the file-based exchange end to end: stage 1 partial aggregate -> ShuffleWriterExec (hash on the group key), stage 2 per
    # partition IpcReaderExec -> final aggregate; the union of the partitions equals a one-stage aggregate
    monkeypatch.setenv("AURON_IO_COMPRESSION_CODEC", codec)
    rng = np.random.default_rng(23)
    n, nparts = 300_000, 5
    t = pa.table({"k": pa.array(rng.integers(0, 5000, n), type=pa.int64(), mask=rng.random(n) < 0.01),
                  "v": pa.array(rng.integers(-1000, 1000, n), type=pa.int64(), mask=rng.random(n) < 0.05)})
    data, index = str(tmp_path / "x.data"), str(tmp_path / "x.index")
    part = P.agg(P.ffi_reader(t.schema, "t"), [P.col("k")], ["k"],
                 [P.agg_expr("SUM", [P.col("v")], pa.int64()), P.agg_expr("COUNT", [P.col("v")], pa.int64())], ["s", "c"], ["PARTIAL"] * 2)
    with runtime.Task(P.task_definition(part), {"t": batches(t, 100_000)}) as task:
        pschema = task.schema
    run(P.shuffle_writer(part, P.hash_repartition([P.col("k")], nparts), data, index), {"t": t}, chunk=100_000)
    offsets = struct.unpack(f"<{nparts + 1}q", open(index, "rb").read())
    finals = []
    for p in range(nparts):
        final = P.agg(P.ipc_reader(pschema, "in"), [P.col("k")], ["k"],
                      [P.agg_expr("SUM", [P.lit(None, pa.null())], pa.int64()), P.agg_expr("COUNT", [P.lit(None, pa.null())], pa.int64())], ["s", "c"], ["FINAL"] * 2)
        finals.append(runtime.run_task(P.task_definition(final), shuffle_blocks={"in": [(data, offsets[p], offsets[p + 1] - offsets[p])]}))
    got = pa.concat_tables([f for f in finals if f.num_rows])
    exp = oracle.agg_sum_count_i64(t["k"].combine_chunks(), t["v"].combine_chunks())
    assert_same_rows(got, exp)
    keys = [f.column(0).to_pylist() for f in finals]
    assert sum(len(k) for k in keys) == len(set(x for k in keys for x in k))       # a key lives in exactly one partition


@pytest.mark.parametrize("decode", ["device", "host"])
def test_ipc_reader_round_trip_all_types(tmp_path, decode, monkeypatch):
    # ShuffleWriterExec's LZ4 frames (independent 64 KB blocks) are decoded on the GPU, layout walk included; AURON_HOST_LZ4_DECODE=1
    # sends the same bytes through liblz4 and the host-side walk.  Every column type of the format, several chunks per partition.
    if decode == "host":
        monkeypatch.setenv("AURON_HOST_LZ4_DECODE", "1")
    n, nparts = 150_000, 4
    t = _table(n, seed=31)
    data, index = str(tmp_path / "rt.data"), str(tmp_path / "rt.index")
    run(P.shuffle_writer(P.ffi_reader(t.schema, "t"), P.hash_repartition([P.col("k")], nparts), data, index), {"t": t}, chunk=40_000)
    offsets = struct.unpack(f"<{nparts + 1}q", open(index, "rb").read())
    pid = oracle.partition_ids([t["k"].combine_chunks()], nparts)
    td = P.task_definition(P.ipc_reader(t.schema, "in"))
    for p in range(nparts):
        got = runtime.run_task(td, shuffle_blocks={"in": [(data, offsets[p], offsets[p + 1] - offsets[p])]})
        assert_same_rows(got, t.filter(pa.array(pid == p)))
    # all partitions through one reader, mixed block shapes (file segments and an in-memory copy)
    raw = open(data, "rb").read()
    blocks = [(data, offsets[0], offsets[1] - offsets[0]), raw[offsets[1]:offsets[2]], (data, offsets[2], offsets[4] - offsets[2])]
    got = runtime.run_task(td, shuffle_blocks={"in": blocks})
    assert_same_rows(got, t)


def test_partitioning_goldens_from_the_reference(tmp_path):
    # shuffle/buffered_data.rs:396-542: round-robin(4) starting at 3, range partitioning with one-key bounds [11, 14, 17] and
    # two-key bounds [(11, 1), (14, 3), (17, 5)] over the same 10-row table; rows inside a partition are unordered (radix sort)
    t = pa.table({"a": pa.array([19, 18, 17, 16, 15, 14, 13, 12, 11, 10], type=pa.int32()),
                  "b": pa.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 9], type=pa.int32()),
                  "c": pa.array([5, 6, 7, 8, 9, 0, 1, 2, 3, 4], type=pa.int32())})
    data, index = str(tmp_path / "g.data"), str(tmp_path / "g.index")

    def parts_of(repartition, partition_id=0):
        td = P.task_definition(P.shuffle_writer(P.ffi_reader(t.schema, "t"), repartition, data, index), stage_id=0, partition_id=partition_id, task_id=1)
        runtime.run_task(td, {"t": t.to_batches()})
        parts, _ = read_shuffle_files(data, index, t.schema)
        return [sorted(p["a"].to_pylist()) for p in parts]

    # start = (partition_id * 1000193 + rows so far) % 4 = 3 for partition_id 3 (the reference test passes current_num_rows = 3)
    assert parts_of(P.round_robin_repartition(4), partition_id=3) == [[10, 14, 18], [13, 17], [12, 16], [11, 15, 19]]
    asc = lambda c: P.sort_expr(P.col(c), True, True)
    assert parts_of(P.range_repartition([asc("a")], 4, [([11, 14, 17], pa.int32())])) == [[10, 11], [12, 13, 14], [15, 16, 17], [18, 19]]
    assert parts_of(P.range_repartition([asc("a"), asc("b")], 4, [([11, 14, 17], pa.int32()), ([1, 3, 5], pa.int32())])) == [
        [10], [11, 12, 13], [14, 15, 16, 17], [18, 19]]


@pytest.mark.parametrize("codec", ["lz4", "zstd"])
def test_ipc_writer_delivers_blocks_to_the_consumer(codec, monkeypatch):
    # IpcWriterExec (ipc_writer_exec.rs:106-190; the broadcast-exchange path NativeBroadcastExchangeBase.scala:317-328): the input in
    # the Auron compacted format, `u32 length | codec stream` blocks, handed to the consumer registered under the resource id --
    # decoded here by the oracle's reader + Arrow's codecs, and fed back through IpcReaderExec as in-memory blocks
    monkeypatch.setenv("AURON_IO_COMPRESSION_CODEC", codec)
    monkeypatch.setenv("AURON_GPU_CHUNK_ROWS", "25000")
    t = _table(70_000, 5)
    sink = []
    td = P.task_definition(P.ipc_writer(P.ffi_reader(t.schema, "in"), "consumer-1"))
    with runtime.Task(td, {"in": batches(t, 10_000)}, ipc_consumers={"consumer-1": sink}) as task:
        assert list(task) == []                       # the writer's own output stream is empty
        met = {(op, name): v for _, op, name, v in task.metrics()}
    assert len(sink) >= 3 and met[("IpcWriterExec", "output_rows")] == t.num_rows
    rows = []
    for seg in sink:
        pos = 0
        while pos < len(seg):
            (blen,) = struct.unpack_from("<I", seg, pos)
            pos += 4
            raw = pa.CompressedInputStream(pa.BufferReader(seg[pos:pos + blen]), codec).read()
            pos += blen
            bpos = 0
            while bpos < len(raw):
                b, bpos = oracle.serde_read_batch(raw, t.schema, bpos)
                rows.append(b)
    assert pa.Table.from_batches(rows, schema=t.schema).equals(t)
    # and the engine's own reader takes the same bytes (hasByteBuffer blocks)
    rplan = P.task_definition(P.ipc_reader(t.schema, "bcast"))
    with runtime.Task(rplan, shuffle_blocks={"bcast": list(sink)}) as task:
        back = pa.Table.from_batches(list(task), schema=task.schema)
    assert back.equals(t)


def test_corrupt_shuffle_block_is_reported_as_fetch_failure():
    # ipc_reader_exec.rs:211-219: undecodable shuffle data -> AuronBlockObject.throwFetchFailed (Spark then recomputes the map output);
    # over the C ABI: the fetch_failed upcall with the resource id, then the error from next_batch
    t = _table(5_000, 9)
    good = _reference_style_segment(t, "lz4", 1000)
    bad = bytearray(good)
    for i in range(40, 200):
        bad[i] ^= 0x5A
    td = P.task_definition(P.ipc_reader(t.schema, "blocks"))
    task = runtime.Task(td, shuffle_blocks={"blocks": [bytes(bad)]})
    with pytest.raises(runtime.AuronError):
        list(task)
    assert task.fetch_failures and task.fetch_failures[0][0] == "blocks" and task.fetch_failures[0][1].startswith("shuffle read:")
    task.close()


@pytest.mark.parametrize("codec", ["lz4", "zstd"])
@pytest.mark.parametrize("pwrite", [False, True])
def test_shuffle_writer_spills_buffered_chunks_and_writes_the_same_file(tmp_path, codec, pwrite, monkeypatch):
    # sort_repartitioner.rs:98-112: the repartitioner spills under memory pressure.  Here finished chunks wait in pinned host memory;
    # with a budget of one byte every chunk goes to a spill file next to the data file, and the final .data / .index must be
    # byte-identical to the unspilled run (same blocks, same chunk order inside every partition), the spill files gone.
    t = _table(120_000, seed=77)
    monkeypatch.setenv("AURON_IO_COMPRESSION_CODEC", codec)
    monkeypatch.setenv("AURON_GPU_CHUNK_ROWS", "25000")
    if pwrite:
        monkeypatch.setenv("AURON_SHUFFLE_PWRITE", "1")
    files = {}
    for mode in ("mem", "spill"):
        if mode == "spill":
            monkeypatch.setenv("AURON_SHUFFLE_SPILL_BYTES", "1")
        data, index = str(tmp_path / f"{mode}.data"), str(tmp_path / f"{mode}.index")
        plan = P.shuffle_writer(P.ffi_reader(t.schema, "t"), P.hash_repartition([P.col("k"), P.col("s")], 17), data, index)
        with runtime.Task(P.task_definition(plan), {"t": t.to_batches(max_chunksize=25_000)}) as task:
            assert list(task) == []
            met = {name: v for _, _, name, v in task.metrics()}
        files[mode] = (open(data, "rb").read(), open(index, "rb").read(), met)
    assert files["mem"][2].get("mem_spill_count", 0) == 0 and files["spill"][2]["mem_spill_count"] >= 4
    assert sorted(os.listdir(tmp_path)) == ["mem.data", "mem.index", "spill.data", "spill.index"]
    # same rows in the same order inside every partition (the GPU LZ4 match finder is free to pick different matches from run to run,
    # so the compressed bytes are compared only when the host codec produced them)
    parts_s, off_s = read_shuffle_files(str(tmp_path / "spill.data"), str(tmp_path / "spill.index"), t.schema, codec)
    parts_m, off_m = read_shuffle_files(str(tmp_path / "mem.data"), str(tmp_path / "mem.index"), t.schema, codec)
    assert len(parts_s) == len(parts_m) == 17 and sum(p.num_rows for p in parts_s) == t.num_rows
    for a, b in zip(parts_s, parts_m):
        assert a.equals(b)
    if codec == "zstd":
        assert files["spill"][0] == files["mem"][0] and files["spill"][1] == files["mem"][1]
