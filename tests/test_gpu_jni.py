"""The JNI face (auron_b200/csrc/jni_face.cc) driven end to end by a mock JVM (tests/jni_mock/mock_jvm.cc): the natives of
JniBridge.java:49-55 are called the way AuronCallNativeWrapper.java does, every upcall the reference makes on this path
(jni_bridge.rs class tables) is answered by a mock class with the real name and signatures, and the batches that arrive through
importBatch are compared with the C-ABI path on the same plan."""
import struct

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from auron_b200 import proto as P
from auron_b200 import runtime
from helpers import assert_same_rows, batches, canon
from jni_helpers import MockJvm
import oracle

pytestmark = pytest.mark.gpu


def _input(n=50_000, seed=3):
    rng = np.random.default_rng(seed)
    return pa.table({"k": pa.array(rng.integers(0, 97, n), type=pa.int32(), mask=rng.random(n) < 0.03),
                     "v": pa.array(rng.integers(-1000, 1000, n), type=pa.int64()),
                     "s": pa.array([f"s{int(i)}" for i in rng.integers(0, 20, n)])})


def _agg_plan(src, t):
    f = P.filter_(src, [P.binary("Gt", P.col("v"), P.lit(-500, pa.int64()))])
    return P.agg(f, [P.col("k")], ["k"], [P.agg_expr("SUM", [P.col("v")], pa.int64()), P.agg_expr("COUNT", [P.col("s")], pa.int64())],
                 ["sv", "c"], ["PARTIAL"] * 2)


def test_ffi_reader_plan_through_jni_matches_the_c_abi():
    t = _input()
    td = P.task_definition(_agg_plan(P.ffi_reader(t.schema, "in"), t), stage_id=2, partition_id=0, task_id=11)
    exp = runtime.run_task(td, {"in": batches(t, 8_000)})
    jvm = MockJvm(td)
    jvm.put_exporter("in", batches(t, 8_000))
    got = jvm.run()
    assert got.schema == exp.schema                                   # importSchema (rt.rs:167-170)
    assert_same_rows(got, exp)                                        # importBatch (rt.rs:258-262)
    assert jvm.M.mock_exporter_closed(jvm.vm, b"in") == 1             # ffi_reader_exec.rs:195
    # update_metric_node (metrics.rs:22-58): root = AggExec, child 0 = FilterExec, its child 0 = FFIReaderExec
    m = jvm.metrics()
    assert m[":output_rows"] == exp.num_rows
    assert m["/0/0:output_rows"] == t.num_rows
    assert not any(k.startswith("/1") for k in m), m                  # single-child chain: nothing lands on a sibling
    jvm.assert_clean()


def test_parquet_scan_reads_through_the_hadoop_fs_wrapper(tmp_path):
    rng = np.random.default_rng(5)
    n = 3_000_000                                                      # ~40 MB of incompressible values: many 4 MB read slices
    t = pa.table({"k": pa.array(rng.integers(0, 97, n), type=pa.int32()), "v": pa.array(rng.integers(-2**40, 2**40, n), type=pa.int64()),
                  "s": pa.array(rng.integers(0, 2**60, n), type=pa.int64())})
    path = str(tmp_path / "t.parquet")
    pq.write_table(t, path, row_group_size=500_000, compression="snappy", use_dictionary=False)
    import os
    scan = P.parquet_scan(t.schema, [(path, os.path.getsize(path))], [0, 1, 2], fs_resource_id="fs-1")
    td = P.task_definition(_agg_plan(scan, t))
    exp = runtime.run_task(td, {})
    jvm = MockJvm(td)
    jvm.put_fs_provider("fs-1")
    got = jvm.run()
    assert_same_rows(got, exp)
    # every byte came through FSDataInputWrapper.readFully (hadoop_fs.rs:85-96), one wrapper per file, closed at the end
    assert jvm.counter("input_wrappers") == 1 and jvm.counter("input_wrappers_closed") == 1
    assert jvm.counter("input_reads") >= 10                           # footer + 4 MB slices of the column chunks
    # the scan's own threads (producer + read workers) attached themselves to the JVM; the producer ended with the scan and
    # detached again
    assert jvm.counter("attached_threads") >= 2 and jvm.counter("detached_threads") >= 1
    # ... after taking over the task thread's class loader and thread context (rt.rs:117-134)
    assert jvm.counter("reads_without_context") == 0
    jvm.assert_clean()


def test_ipc_reader_takes_blocks_of_every_kind_from_the_scala_iterator(tmp_path):
    # write a real shuffle file with the engine, then read its segments back as the four shapes a BlockObject can take
    # (ipc_reader_exec.rs:199-207: file segment, direct ByteBuffer, heap ByteBuffer, ReadableByteChannel)
    t = _input(40_000, seed=9)
    data, index = str(tmp_path / "s.data"), str(tmp_path / "s.index")
    w = P.shuffle_writer(P.ffi_reader(t.schema, "in"), P.hash_repartition([P.col("k")], 4), data, index)
    runtime.run_task(P.task_definition(w), {"in": batches(t)})
    offs = struct.unpack("<5q", open(index, "rb").read())
    raw = open(data, "rb").read()
    td = P.task_definition(P.ipc_reader(t.schema, "blocks"))
    jvm = MockJvm(td)
    kinds = ["file", "direct", "heap", "channel"]
    for p, kind in enumerate(kinds):
        seg = raw[offs[p]:offs[p + 1]]
        assert len(seg) > 0
        if kind == "file":
            jvm.add_block("blocks", "file", path=data, offset=offs[p], length=len(seg))
        else:
            jvm.add_block("blocks", kind, data=seg)
    got = jvm.run()
    assert_same_rows(got, t)
    assert jvm.M.mock_blocks_closed(jvm.vm, b"blocks") == 4           # each block closed exactly once (:383-402)
    jvm.assert_clean()


def test_corrupt_block_raises_fetch_failed_on_the_block_object():
    # ipc_reader_exec.rs:211-219: a block that does not decode -> BlockObject.throwFetchFailed(message); Spark's
    # FetchFailedException is the cause the task fails with, so the scheduler recomputes the map output
    t = _input(8_000, seed=3)
    payload = b"".join(oracle.serde_write_batch(b) for b in batches(t, 2_000))
    sink = pa.BufferOutputStream()
    with pa.CompressedOutputStream(sink, "zstd") as z:
        z.write(payload)
    comp = bytearray(sink.getvalue().to_pybytes())
    for i in range(30, 90):
        comp[i] ^= 0xA5
    jvm = MockJvm(P.task_definition(P.ipc_reader(t.schema, "blocks")))
    jvm.add_block("blocks", "heap", data=struct.pack("<I", len(comp)) + bytes(comp))
    assert jvm.call_native()
    while jvm.load_next_batch() is not None:
        pass
    err = jvm.error()
    assert "shuffle read:" in err and "<- org/apache/spark/shuffle/FetchFailedException: shuffle read:" in err, err
    jvm.close()
    jvm.assert_clean()


def test_exporter_failure_reaches_set_error_with_its_cause():
    t = _input(20_000)
    td = P.task_definition(_agg_plan(P.ffi_reader(t.schema, "in"), t))
    jvm = MockJvm(td)
    jvm.put_exporter("in", batches(t, 5_000), fail_after=2)
    assert jvm.call_native()
    assert jvm.load_next_batch() is None
    # set_error (rt.rs:309-318): RuntimeException(message, cause) handed to wrapper.setError; nothing left pending
    err = jvm.error()
    assert err.startswith("java/lang/RuntimeException: ") and err.endswith("<- java/lang/IllegalStateException: exporter failed"), err
    assert jvm.pending_exception() == ""
    jvm.close()
    jvm.assert_clean()


def test_consumer_failure_in_import_batch_stays_pending_for_the_caller():
    t = _input(20_000)
    td = P.task_definition(P.ffi_reader(t.schema, "in"))
    jvm = MockJvm(td)
    jvm.put_exporter("in", batches(t, 5_000))
    jvm.M.mock_wrapper_fail_import_after(jvm.wrapper, 0)
    assert jvm.call_native()
    assert jvm.load_next_batch() is None
    assert jvm.pending_exception() == "java/lang/IllegalStateException: consumer failed"
    jvm.close()
    jvm.assert_clean()


def test_bad_plan_raises_from_call_native_and_killed_task_stops():
    jvm = MockJvm(b"\x0a\x03abc")
    assert not jvm.call_native()
    assert jvm.pending_exception().startswith("java/lang/RuntimeException: ")
    jvm.assert_clean()
    # JniBridge.isTaskRunning() == false (auron-jni-bridge/src/lib.rs:35-50): the stream ends with an error, not with data
    t = _input(20_000)
    td = P.task_definition(_agg_plan(P.ffi_reader(t.schema, "in"), t))
    jvm = MockJvm(td)
    jvm.put_exporter("in", batches(t, 5_000))
    jvm.M.mock_set_task_running(jvm.vm, 0)
    assert jvm.call_native()
    assert jvm.load_next_batch() is None
    assert "RuntimeException" in jvm.error()
    jvm.close()
    jvm.assert_clean()


def test_union_lands_metrics_on_sibling_nodes():
    a, b = _input(10_000, seed=1), _input(7_000, seed=2)
    plan = P.union([P.ffi_reader(a.schema, "a"), P.ffi_reader(b.schema, "b")], a.schema)
    jvm = MockJvm(P.task_definition(plan))
    jvm.put_exporter("a", batches(a))
    jvm.put_exporter("b", batches(b))
    got = jvm.run()
    assert got.num_rows == 17_000
    m = jvm.metrics()
    assert m["/0:output_rows"] == 10_000 and m["/1:output_rows"] == 7_000   # MetricNode.getChild(i) per plan child
    jvm.assert_clean()


def test_shuffle_writer_takes_its_codec_from_the_jvm_conf(tmp_path):
    # conf.rs:46-47 + ipc_compression.rs:180-200: JniBridge.stringConf("SPARK_IO_COMPRESSION_CODEC") / intConf(...ZSTD_LEVEL)
    t = _input(30_000, seed=4)
    magic = {"lz4": b"\x04\x22\x4d\x18", "zstd": b"\x28\xb5\x2f\xfd"}
    for codec in ("zstd", "lz4", None):
        data, index = str(tmp_path / f"{codec}.data"), str(tmp_path / f"{codec}.index")
        w = P.shuffle_writer(P.ffi_reader(t.schema, "in"), P.hash_repartition([P.col("k")], 3), data, index)
        jvm = MockJvm(P.task_definition(w))
        jvm.put_exporter("in", batches(t))
        if codec:
            jvm.set_conf("SPARK_IO_COMPRESSION_CODEC", codec)
            jvm.set_conf("SPARK_IO_COMPRESSION_ZSTD_LEVEL", "3")
        jvm.run()
        raw = open(data, "rb").read()
        assert raw[4:8] == magic[codec or "lz4"]                      # u32 block length, then the codec's frame magic
        jvm.assert_clean()
        # and the file reads back through the JNI block iterator
        offs = struct.unpack("<4q", open(index, "rb").read())
        rd = MockJvm(P.task_definition(P.ipc_reader(t.schema, "blocks")))
        for p in range(3):
            rd.add_block("blocks", "file", path=data, offset=offs[p], length=offs[p + 1] - offs[p])
        assert_same_rows(rd.run(), t)
        rd.assert_clean()


def test_whole_query_through_jni(tmp_path):
    # the q3-shaped plan of test_gpu_queries.py (three Parquet scans through the Hadoop FS wrapper, two joins, two-phase aggregate,
    # sort + limit) driven through the JNI natives: same rows, in the same order, as through the C ABI
    from test_gpu_queries import FS_RESOURCE, q3_plan
    plan, _ = q3_plan(tmp_path)
    td = P.task_definition(plan)
    exp = runtime.run_task(td, {})
    jvm = MockJvm(td)
    jvm.put_fs_provider(FS_RESOURCE)
    got = jvm.run()
    assert got.schema == exp.schema and got.num_rows == 100
    assert got.to_pylist() == exp.to_pylist()
    assert jvm.counter("input_wrappers") == 3 == jvm.counter("input_wrappers_closed")
    jvm.assert_clean()
    # a scan whose FS resource was never registered fails the task with an error, not a crash
    jvm = MockJvm(td)
    assert jvm.call_native()
    assert jvm.load_next_batch() is None
    assert "RuntimeException" in jvm.error()
    jvm.close()
    jvm.assert_clean()
