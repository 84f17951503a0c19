"""The plan bytes that cross the boundary.  Every GPU parity test hands the engine a TaskDefinition built by auron_b200/proto.py; this
module pins that encoder -- and through it the planner that decodes those bytes on the other side -- to the reference's own wire
contract: tests/golden/auron_proto_schema.json is extracted from native-engine/auron-planner/proto/auron.proto (tools/
extract_proto_schema.py, run where the reference is mounted), and a schema-driven protobuf decoder walks the encoder's output:
every field number must exist in the message the reference declares at that position, with the wire type its declared type implies,
enums must carry declared values, and nested messages are walked recursively."""
import decimal
import json
import os

import pyarrow as pa
import pytest

from auron_b200 import proto as P

HERE = os.path.dirname(os.path.abspath(__file__))
SCHEMA = json.load(open(os.path.join(HERE, "golden", "auron_proto_schema.json")))
MSG, ENUM = SCHEMA["messages"], SCHEMA["enums"]
VARINT_TYPES = {"int32", "int64", "uint32", "uint64", "sint32", "sint64", "bool"}
LEN_TYPES = {"string", "bytes"}
FIXED64, FIXED32 = {"double", "fixed64", "sfixed64"}, {"float", "fixed32", "sfixed32"}


def _varint(buf, i):
    v = s = 0
    while True:
        b = buf[i]
        i += 1
        v |= (b & 0x7F) << s
        if not b & 0x80:
            return v, i
        s += 7


def walk(msg: str, buf: bytes, seen: set, path: str = ""):
    """decode `buf` as reference message `msg`; returns {field name: [values]} and records (message, field) in `seen`"""
    by_num = {v[0]: (k, v[1], v[2]) for k, v in MSG[msg].items()}
    out, i = {}, 0
    while i < len(buf):
        key, i = _varint(buf, i)
        num, wire = key >> 3, key & 7
        assert num in by_num, f"{path}{msg}: field number {num} is not declared by the reference"
        name, typ, repeated = by_num[num]
        seen.add((msg, name))
        where = f"{path}{msg}.{name}"
        if wire == 0:
            v, i = _varint(buf, i)
            assert typ in VARINT_TYPES or typ in ENUM, f"{where}: varint on the wire, declared {typ}"
            if typ in ENUM:
                assert v in ENUM[typ].values(), f"{where}: {v} is not a value of enum {typ}"
        elif wire == 2:
            n, i = _varint(buf, i)
            v = buf[i:i + n]
            i += n
            assert len(v) == n, f"{where}: truncated"
            if typ in MSG:
                v = walk(typ, v, seen, path + msg + ".")
            elif repeated and (typ in VARINT_TYPES or typ in ENUM):   # packed repeated scalars
                vals, j = [], 0
                while j < len(v):
                    x, j = _varint(v, j)
                    vals.append(x)
                v = vals
            else:
                assert typ in LEN_TYPES, f"{where}: length-delimited on the wire, declared {typ}"
                if typ == "string":
                    v = v.decode("utf-8")
        elif wire == 1:
            assert typ in FIXED64, f"{where}: 64-bit on the wire, declared {typ}"
            v, i = buf[i:i + 8], i + 8
        elif wire == 5:
            assert typ in FIXED32, f"{where}: 32-bit on the wire, declared {typ}"
            v, i = buf[i:i + 4], i + 4
        else:
            raise AssertionError(f"{where}: wire type {wire}")
        if not repeated and name in out and typ not in MSG:
            raise AssertionError(f"{where}: singular field encoded twice")
        out.setdefault(name, []).append(v)
    return out


I, L, S = pa.int32(), pa.int64(), pa.string()
T = pa.schema([("a", I), ("b", L), ("s", S), ("d", pa.decimal128(7, 2)), ("t", pa.timestamp("us")), ("x", pa.float64()), ("dt", pa.date32()),
               ("f", pa.bool_())])


def _plans():
    """one plan per node / expression kind the engine accepts, built exactly as the GPU tests build them"""
    src = P.ffi_reader(T, "in")
    dec = decimal.Decimal
    exprs = [P.binary("Plus", P.col("a"), P.lit(1, I)), P.binary("And", P.is_null(P.col("a")), P.is_not_null(P.col("b"))), P.not_(P.col("f")),
             P.negative(P.col("a")), P.case([(P.binary("Gt", P.col("a"), P.lit(0, I)), P.lit("p", S))], P.lit(None, S)),
             P.cast(P.col("a"), L), P.try_cast(P.col("s"), pa.decimal128(12, 3)), P.in_list(P.col("a"), [P.lit(1, I), P.lit(2, I)], negated=True),
             P.scalar_fn("Spark_Hour", [P.col("t"), P.lit("Asia/Shanghai", S)], I), P.scalar_fn("Substr", [P.col("s"), P.lit(1, L), P.lit(2, L)], S),
             P.like(P.col("s"), P.lit("a%", S)), P.sc_and(P.col("f"), P.col("f")), P.sc_or(P.col("f"), P.col("f")),
             P.starts_with(P.col("s"), "a"), P.ends_with(P.col("s"), "z"), P.contains(P.col("s"), "m"), P.bound_ref(0, I),
             P.lit(dec("12.34"), pa.decimal128(7, 2)), P.lit(1.5, pa.float64()), P.lit(True, pa.bool_())]
    names = [f"e{i}" for i in range(len(exprs))]
    types = [I, pa.bool_(), pa.bool_(), I, S, L, pa.decimal128(12, 3), pa.bool_(), I, S] + [pa.bool_()] * 6 + [I, pa.decimal128(7, 2), pa.float64(), pa.bool_()]
    proj = P.projection(P.filter_(src, [P.binary("GtEq", P.col("a"), P.lit(0, I))]), exprs, names, types)
    aggs = [P.agg_expr(f, [P.col("b")], L) for f in ("SUM", "COUNT", "MIN", "MAX", "AVG", "FIRST")]
    agg = P.agg(src, [P.col("a"), P.col("s")], ["a", "s"], aggs, [f"g{i}" for i in range(6)], ["PARTIAL"] * 6)
    agg_f = P.agg(agg, [P.col("a")], ["a"], [P.agg_expr("SUM", [P.lit(None, pa.null())], L)], ["g"], ["FINAL"])
    on = [(P.col("a"), P.col("a"))]
    js = pa.schema(list(T) + list(T))
    joins = [P.hash_join(js, src, src, on, "LEFT", "RIGHT"), P.sort_merge_join(js, src, src, on, "FULL"),
             P.broadcast_join(js, src, src, on, "SEMI", "LEFT", cached_id="bc-1", null_aware_anti=True)]
    sort = P.sort(src, [P.sort_expr(P.col("a"), False, False), P.sort_expr(P.col("s"))], limit=10, offset=2)
    scan = P.parquet_scan(T, [("/tmp/x.parquet", 1234), ("/tmp/y.parquet", 99)], [0, 2, 3], fs_resource_id="fs", ranges=[(0, 100), (5, 50)],
                          pruning_predicates=[P.binary("Lt", P.col("a"), P.lit(5, I))])
    parts = [P.hash_repartition([P.col("a"), P.col("s")], 7), P.single_repartition(), P.round_robin_repartition(5),
             P.range_repartition([P.sort_expr(P.col("a"))], 3, [([10, 20], I)])]
    writers = [P.shuffle_writer(src, p, "/tmp/d", "/tmp/i") for p in parts]
    misc = [P.limit(src, 5, 1), P.rename_columns(src, [f"c{i}" for i in range(len(T))]), P.union([src, src], T), P.ipc_reader(T, "blocks")]
    # nodes added in round 2 (appended: the positions above are used by index)
    expand = P.expand(src, pa.schema([("a", L), ("g", L)]), [[P.col("a"), P.lit(0, L)], [P.lit(None, L), P.lit(1, L)]])
    window = P.window(P.sort(src, [P.sort_expr(P.col("a")), P.sort_expr(P.col("b"))]),
                      [P.window_expr("rk", I, "RANK"), P.window_expr("sb", L, "SUM", [P.col("b")]), P.window_expr("ld", L, "LEAD", [P.col("b"), P.lit(1, I), P.lit(None, L)]),
                       P.window_expr("nv", S, "NTH_VALUE_IGNORE_NULLS", [P.col("s"), P.lit(2, I)]), P.window_expr("cd", pa.float64(), "CUME_DIST")],
                      [P.col("a")], [P.sort_expr(P.col("b"))])
    limited = P.window(src, [P.window_expr("rk", I, "DENSE_RANK")], [P.col("a")], [P.sort_expr(P.col("b"), False, False)], group_limit=2, output_window_cols=False)
    return [proj, agg, agg_f, sort, scan] + joins + writers + misc + [expand, window, limited, P.ipc_writer(src, "consumer")]


def test_every_encoded_plan_is_a_well_formed_reference_message():
    seen = set()
    for plan in _plans():
        td = walk("TaskDefinition", P.task_definition(plan, stage_id=3, partition_id=4, task_id=5), seen)
        pid = td["task_id"][0]
        assert (pid["stage_id"], pid["partition_id"], pid["task_id"]) == ([3], [4], [5])
    # the walk really went through the messages of the path (a vacuous pass would not)
    nodes = {f for m, f in seen if m == "PhysicalPlanNode"}
    assert nodes == {"ffi_reader", "ipc_reader", "filter", "projection", "agg", "sort", "parquet_scan", "hash_join", "sort_merge_join", "broadcast_join",
                     "shuffle_writer", "limit", "rename_columns", "union", "expand", "window", "ipc_writer"}
    exprs = {f for m, f in seen if m == "PhysicalExprNode"}
    assert {"column", "literal", "bound_reference", "binary_expr", "agg_expr", "is_null_expr", "is_not_null_expr", "not_expr", "case_", "cast",
            "try_cast", "sort", "negative", "in_list", "scalar_function", "like_expr", "sc_and_expr", "sc_or_expr", "string_starts_with_expr",
            "string_ends_with_expr", "string_contains_expr"} <= exprs
    reps = {f for m, f in seen if m == "PhysicalRepartition"}
    assert reps == {"single_repartition", "hash_repartition", "round_robin_repartition", "range_repartition"}
    assert {("AggExecNode", "mode"), ("FetchLimit", "offset"), ("FileScanExecConf", "projection"), ("ParquetScanExecNode", "fsResourceId"),
            ("BroadcastJoinExecNode", "cached_build_hash_map_id"), ("BroadcastJoinExecNode", "is_null_aware_anti_join"), ("PartitionedFile", "range"),
            ("FileRange", "end"), ("PhysicalRangeRepartition", "sort_expr") if "sort_expr" in MSG["PhysicalRangeRepartition"] else ("AggExecNode", "mode")} <= seen


def test_decoded_values_mean_what_the_builders_were_given():
    seen = set()
    plan = P.agg(P.ffi_reader(T, "rid-7"), [P.col("a")], ["ka"], [P.agg_expr("MAX", [P.col("b")], L), P.agg_expr("AVG", [P.col("d")], pa.decimal128(11, 6))],
                 ["m", "v"], ["PARTIAL", "PARTIAL"], supports_partial_skipping=True)
    node = walk("PhysicalPlanNode", plan, seen)["agg"][0]
    assert node["grouping_expr_name"] == ["ka"] and node["agg_expr_name"] == ["m", "v"]
    modes = [x for v in node["mode"] for x in (v if isinstance(v, list) else [v])]
    assert modes == [ENUM["AggMode"]["PARTIAL"]] * 2
    assert node["supports_partial_skipping"] == [1]
    fns = [a["agg_expr"][0]["agg_function"][0] for a in node["agg_expr"]]
    assert fns == [ENUM["AggFunction"]["MAX"], ENUM["AggFunction"]["AVG"]]
    reader = node["input"][0]["ffi_reader"][0]
    assert reader["export_iter_provider_resource_id"] == ["rid-7"]
    cols = reader["schema"][0]["columns"]
    assert [c["name"][0] for c in cols] == T.names
    assert "DECIMAL" in cols[3]["arrow_type"][0] and cols[3]["arrow_type"][0]["DECIMAL"][0]["whole"] == [7]
    join = walk("PhysicalPlanNode", P.hash_join(pa.schema(list(T) + list(T)), P.ffi_reader(T, "l"), P.ffi_reader(T, "r"), [(P.col("a"), P.col("b"))],
                                                "ANTI", "LEFT"), seen)["hash_join"][0]
    assert join["join_type"] == [ENUM["JoinType"]["ANTI"]]
    assert join["on"][0]["left"][0]["column"][0]["name"] == ["a"] and join["on"][0]["right"][0]["column"][0]["name"] == ["b"]
    srt = walk("PhysicalPlanNode", P.sort(P.ffi_reader(T, "x"), [P.sort_expr(P.col("a"), False, True)], limit=9, offset=4), seen)["sort"][0]
    key = srt["expr"][0]["sort"][0]
    assert key.get("asc", [0]) == [0] and key["nulls_first"] == [1]
    assert (srt["fetch_limit"][0]["limit"], srt["fetch_limit"][0]["offset"]) == ([9], [4])


def _explain(plan: bytes, **ids) -> dict:
    import ctypes as C
    from auron_b200 import runtime
    L = runtime.lib()
    L.auron_b200_explain.restype = C.c_int64
    L.auron_b200_explain.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_int64]
    L.auron_b200_last_error.restype = C.c_char_p
    td = P.task_definition(plan, **ids)
    buf = C.create_string_buffer(1 << 20)
    n = L.auron_b200_explain(td, len(td), buf, len(buf))
    assert n > 0, L.auron_b200_last_error().decode()
    return json.loads(buf.value.decode())


def test_planner_decodes_every_plan_into_the_operators_it_was_built_as():
    # the other side of the wire: the engine's own planner (planner.cc, the PhysicalPlanner::create_plan of this engine) decodes the
    # same bytes -- on no device -- and its operator tree is read back: reference operator names, output schemas, every attribute of
    # every node and every expression / literal (Arrow IPC ScalarValues) as they were encoded
    plans = _plans()
    top = _explain(plans[0], stage_id=3, partition_id=4, task_id=5)
    assert (top["stage_id"], top["partition_id"], top["task_id"]) == (3, 4, 5)
    proj = top["plan"]
    assert proj["op"] == "ProjectExec" and proj["children"][0]["op"] == "FilterExec" and proj["children"][0]["children"][0]["op"] == "FFIReaderExec"
    assert proj["exprs"] == ["Plus(col(a), lit(int32:1))", "And(IsNull(col(a)), IsNotNull(col(b)))", "Not(col(f))", "Negative(col(a))",
                             "Case[else](Gt(col(a), lit(int32:0)), lit(utf8:'p'), lit(utf8:NULL))", "Cast(col(a) AS int64)",
                             "TryCast(col(s) AS decimal128(12,3))", "NotIn(col(a), lit(int32:1), lit(int32:2))",
                             "Spark_Hour(col(t), lit(utf8:'Asia/Shanghai')) -> int32", "Substr(col(s), lit(int64:1), lit(int64:2)) -> utf8",
                             "Like(col(s), lit(utf8:'a%'))", "SCAnd(col(f), col(f))", "SCOr(col(f), col(f))", "StartsWith(col(s), 'a')", "EndsWith(col(s), 'z')",
                             "Contains(col(s), 'm')", "col(#0)", "lit(decimal128(7,2):1234)", "lit(float64:1.5)", "lit(bool:1)"]
    assert [f[0] for f in proj["schema"]] == [f"e{i}" for i in range(20)] and proj["schema"][6][1] == "decimal128(12,3)"
    assert proj["children"][0]["predicates"] == ["GtEq(col(a), lit(int32:0))"]
    src = proj["children"][0]["children"][0]
    assert src["resource_id"] == "in" and [f[0] for f in src["schema"]] == T.names
    assert [f[1] for f in src["schema"]] == ["int32", "int64", "utf8", "decimal128(7,2)", "timestamp", "float64", "date32", "bool"]

    agg = _explain(plans[1])["plan"]
    assert agg["op"] == "AggExec" and agg["grouping"] == ["col(a)", "col(s)"]
    assert [(a["fn"], a["mode"], a["args"]) for a in agg["aggs"]] == [(f, "PARTIAL", ["col(b)"]) for f in ("SUM", "COUNT", "MIN", "MAX", "AVG", "FIRST")]
    fin = _explain(plans[2])["plan"]
    assert fin["aggs"] == [{"fn": "SUM", "mode": "FINAL", "args": ["lit(null:NULL)"], "return_type": "int64"}] and fin["children"][0]["op"] == "AggExec"
    assert fin["schema"] == [["a", "int32"], ["g", "int64"]]

    srt = _explain(plans[3])["plan"]
    assert (srt["op"], srt["keys"], srt["limit"], srt["offset"]) == ("SortExec", ["col(a) DESC NULLS LAST", "col(s) ASC NULLS FIRST"], 10, 2)

    scan = _explain(plans[4])["plan"]
    assert scan["op"] == "ParquetExec" and scan["fs_resource_id"] == "fs" and scan["projection"] == [0, 2, 3]
    assert scan["files"] == [{"path": "/tmp/x.parquet", "size": 1234, "range": [0, 100]}, {"path": "/tmp/y.parquet", "size": 99, "range": [5, 50]}]
    assert scan["schema"] == [["a", "int32"], ["s", "utf8"], ["d", "decimal128(7,2)"]]

    attrs = lambda d: {k: v for k, v in d.items() if k not in ("schema", "children")}
    assert attrs(_explain(plans[5])["plan"]) == {"op": "BroadcastJoin", "join_type": "LEFT", "build_side": "RIGHT", "left_keys": ["col(a)"],
                                                 "right_keys": ["col(a)"], "cached_build_hash_map_id": "", "null_aware_anti": False}
    smj = _explain(plans[6])["plan"]
    assert (smj["op"], smj["join_type"]) == ("SortMergeJoinExec", "FULL") and len(smj["schema"]) == 16 and len(smj["children"]) == 2
    assert attrs(_explain(plans[7])["plan"]) == {"op": "BroadcastJoin", "join_type": "SEMI", "build_side": "LEFT", "left_keys": ["col(a)"],
                                                 "right_keys": ["col(a)"], "cached_build_hash_map_id": "bc-1", "null_aware_anti": True}

    writers = [attrs(_explain(p)["plan"]) for p in plans[8:12]]
    assert [(w["partitioning"], w["partition_count"], w["exprs"], w["range_bound_rows"]) for w in writers] == [
        ("hash", 7, ["col(a)", "col(s)"], []), ("single", 1, [], []), ("round_robin", 5, [], []), ("range", 3, ["col(a) ASC NULLS FIRST"], [2])]
    assert all(w["op"] == "ShuffleWriterExec" and (w["data_file"], w["index_file"], w["codec"]) == ("/tmp/d", "/tmp/i", "lz4") for w in writers)

    lim, ren, uni, ipc = (_explain(p)["plan"] for p in plans[12:16])
    assert (lim["op"], lim["limit"], lim["offset"]) == ("LimitExec", 5, 1)
    assert ren["op"] == "RenameColumnsExec" and [f[0] for f in ren["schema"]] == [f"c{i}" for i in range(8)]
    assert uni["op"] == "UnionExec" and [c["op"] for c in uni["children"]] == ["FFIReaderExec"] * 2
    assert (ipc["op"], ipc["resource_id"]) == ("IpcReaderExec", "blocks")


def test_planner_rejects_what_is_outside_the_path_with_a_message():
    import ctypes as C
    from auron_b200 import runtime
    L = runtime.lib()
    L.auron_b200_explain.restype = C.c_int64
    L.auron_b200_explain.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_int64]
    L.auron_b200_last_error.restype = C.c_char_p
    src = P.ffi_reader(T, "in")
    generate = P.f_bytes(23, P.f_bytes(1, src))                                   # PhysicalPlanNode{generate}: not on the path
    lead = P.window(src, [P.window_expr("l", pa.int32(), "FIRST", [P.col("a")])], [], [])   # a window aggregate that is not built
    for plan, needle in [(generate, "not native"), (lead, "not native"), (b"\x0a\x03abc", ""), (P.filter_(src, []), "predicate")]:
        td = P.task_definition(plan) if plan != b"\x0a\x03abc" else plan
        assert L.auron_b200_explain(td, len(td), None, 0) == -1
        assert needle in L.auron_b200_last_error().decode().lower()


def test_planner_skips_fields_of_the_reference_message_it_does_not_use():
    # the JVM side fills fields this engine has no use for (TaskDefinition.output_partitioning, AggExecNode.initial_input_buffer_offset,
    # FileScanExecConf.statistics / partition_schema, PartitionedFile.last_modified_ns ...): they must be skipped, not rejected
    src = P.ffi_reader(T, "in")
    agg = P.agg(src, [P.col("a")], ["a"], [P.agg_expr("COUNT", [P.col("b")], L)], ["c"], ["PARTIAL"])
    # re-encode the agg node with an extra field 8 (uint64 initial_input_buffer_offset) appended to its body
    seen = set()
    node = walk("PhysicalPlanNode", agg, seen)
    assert "agg" in node

    def first_field(buf):
        key, i = _varint(buf, 0)
        n, j = _varint(buf, i)
        return key, buf[j:j + n]

    key, inner = first_field(agg)
    inner2 = inner + P.f_varint(8, 12345)
    agg2 = P.f_bytes(key >> 3, inner2)
    walk("PhysicalPlanNode", agg2, seen)                        # still a valid reference message
    td = P.task_definition(agg2) + P.f_bytes(3, P.single_repartition())      # TaskDefinition.output_partitioning = 3
    walk("TaskDefinition", td, seen)
    import ctypes as C
    from auron_b200 import runtime
    Lb = runtime.lib()
    Lb.auron_b200_explain.restype = C.c_int64
    Lb.auron_b200_explain.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_int64]
    buf = C.create_string_buffer(1 << 16)
    assert Lb.auron_b200_explain(td, len(td), buf, len(buf)) > 0
    d = json.loads(buf.value.decode())["plan"]
    assert d["op"] == "AggExec" and d["aggs"][0]["fn"] == "COUNT" and d["children"][0]["resource_id"] == "in"


def test_wrapper_nodes_of_real_spark_plans_decode_under_their_reference_names():
    # NativeBroadcastJoinBase wraps the broadcast side in BroadcastJoinBuildHashMapExecNode; converters also emit CoalesceBatches, Debug and
    # EmptyPartitions nodes: pass-through / trivial operators here, named as the reference names them
    src = P.ffi_reader(T, "in")
    nodes = {"BroadcastJoinBuildHashMapExec": P.f_bytes(12, P.f_bytes(1, src) + P.f_bytes(2, P.col("a"))),
             "CoalesceBatchesExec": P.f_bytes(19, P.f_bytes(1, src) + P.f_varint(2, 8192)),
             "DebugExec": P.f_bytes(1, P.f_bytes(1, src) + P.f_str(2, "dbg")),
             "EmptyPartitionsExec": P.f_bytes(15, P.f_bytes(1, P.schema(T)) + P.f_varint(2, 4))}
    seen = set()
    for name, plan in nodes.items():
        walk("PhysicalPlanNode", plan, seen)
        d = _explain(plan)["plan"]
        assert d["op"] == name and [f[0] for f in d["schema"]] == T.names
        assert [c["op"] for c in d["children"]] == ([] if name == "EmptyPartitionsExec" else ["FFIReaderExec"])


def test_malformed_plans_are_errors_not_crashes():
    # "Panics/aborts must not kill the JVM" (lib.rs:57-72 catches unwinds in the reference): byte-level damage to valid plans -- in the
    # protobuf framing, in the Arrow IPC flatbuffers of literals, in operand counts -- and absurd nesting must come back as an error
    # (or still decode); the planner is run on thousands of mutants in this process
    import ctypes as C
    import random
    from auron_b200 import runtime
    Lb = runtime.lib()
    Lb.auron_b200_explain.restype = C.c_int64
    Lb.auron_b200_explain.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_int64]
    random.seed(20250921)
    tds = [P.task_definition(p) for p in _plans()]
    ok = bad = 0
    for _ in range(6000):
        b = bytearray(random.choice(tds))
        for _ in range(random.randint(1, 3)):
            r = random.random()
            if r < 0.6:
                b[random.randrange(len(b))] = random.randrange(256)
            elif r < 0.8:
                del b[random.randrange(len(b))]
            else:
                b = b[:random.randrange(1, len(b) + 1)]
        n = Lb.auron_b200_explain(bytes(b), len(b), None, 0)
        ok += n > 0
        bad += n < 0
    assert ok > 100 and bad > 100                                  # both outcomes occur; reaching this line is the point
    deep = P.col("f")
    for _ in range(5000):
        deep = P.not_(deep)
    td = P.task_definition(P.filter_(P.ffi_reader(T, "in"), [deep]))
    assert Lb.auron_b200_explain(td, len(td), None, 0) == -1
    assert _explain(P.filter_(P.ffi_reader(T, "in"), [P.not_(P.col("f"))]))["plan"]["predicates"] == ["Not(col(f))"]   # and the planner still works


def test_count_without_arguments_is_count_star():
    # AggCount over no children counts rows (agg/count.rs:89-157): the argument check of the planner must not reject it
    plan = P.agg(P.ffi_reader(T, "in"), [P.col("a")], ["a"], [P.agg_expr("COUNT", [], L)], ["c"], ["PARTIAL"])
    assert _explain(plan)["plan"]["aggs"] == [{"fn": "COUNT", "mode": "PARTIAL", "args": [], "return_type": "int64"}]
    import ctypes as C
    from auron_b200 import runtime
    Lb = runtime.lib()
    Lb.auron_b200_explain.restype = C.c_int64
    Lb.auron_b200_explain.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_int64]
    bad = P.task_definition(P.agg(P.ffi_reader(T, "in"), [P.col("a")], ["a"], [P.agg_expr("MAX", [], L)], ["m"], ["PARTIAL"]))
    assert Lb.auron_b200_explain(bad, len(bad), None, 0) == -1                # MAX() of nothing is an error, not a crash


def test_expand_and_partitioned_scan_decode_on_the_cpu():
    # ExpandExecNode (auron.proto:745-754) and the Hive-partition fields of FileScanExecConf / PartitionedFile decode into the operators
    # the reference builds (planner.rs:587-602, 1415-1501); walk() checks the bytes against the reference schema first
    src = P.ffi_reader(T, "in")
    out = pa.schema([("a", pa.int64()), ("g", pa.int64())])
    ex = P.expand(src, out, [[P.col("a"), P.lit(0, pa.int64())], [P.lit(None, pa.int64()), P.lit(1, pa.int64())]])
    seen = set()
    walk("PhysicalPlanNode", ex, seen)
    assert ("ExpandExecNode", "projections") in seen and ("ExpandProjection", "expr") in seen
    d = _explain(ex)["plan"]
    assert d["op"] == "ExpandExec" and [f[0] for f in d["schema"]] == ["a", "g"] and len(d["projections"]) == 2
    assert [c["op"] for c in d["children"]] == ["FFIReaderExec"]
    file_schema = pa.schema([("x", pa.int32()), ("y", pa.string())])
    part_schema = pa.schema([("day", pa.int32()), ("region", pa.string())])
    scan = P.parquet_scan(file_schema, [("/data/day=7/region=eu/f.parquet", 123)], [1, 2, 3, 0], partition_schema=part_schema, partition_values=[[7, "eu"]],
                          pruning_predicates=[P.binary("Gt", P.col("x"), P.lit(5, pa.int32()))])
    walk("PhysicalPlanNode", scan, seen)
    assert {("FileScanExecConf", "partition_schema"), ("PartitionedFile", "partition_values"), ("ParquetScanExecNode", "pruning_predicates")} <= seen
    d = _explain(scan)["plan"]
    assert d["op"] == "ParquetExec" and [f[0] for f in d["schema"]] == ["y", "day", "region", "x"]      # projection order over [file columns..., partition columns...]


def test_window_node_decodes_on_the_cpu():
    # WindowExecNode / WindowExprNode / WindowGroupLimit (auron.proto:566-591): bytes checked against the reference schema, then the
    # planner's operator (planner.rs:604-760)
    src = P.sort(P.ffi_reader(T, "in"), [P.sort_expr(P.col("a")), P.sort_expr(P.col("b"))])
    w = P.window(src, [P.window_expr("rk", pa.int32(), "RANK"), P.window_expr("sb", pa.int64(), "SUM", [P.col("b")])], [P.col("a")], [P.sort_expr(P.col("b"))])
    seen = set()
    walk("PhysicalPlanNode", w, seen)
    assert {("WindowExecNode", "window_expr"), ("WindowExecNode", "partition_spec"), ("WindowExecNode", "order_spec"), ("WindowExprNode", "return_type"),
            ("WindowExprNode", "agg_func"), ("WindowExprNode", "children")} <= seen
    d = _explain(w)["plan"]
    assert d["op"] == "WindowExec" and d["functions"] == ["RANK AS rk", "SUM AS sb"] and d["partition_by"] == ["col(a)"] and d["order_by"] == ["col(b)"]
    assert [f[0] for f in d["schema"]][-2:] == ["rk", "sb"] and d["output_window_cols"] is True and d["group_limit"] == -1
    lim = P.window(src, [P.window_expr("rk", pa.int32(), "RANK")], [P.col("a")], [P.sort_expr(P.col("b"))], group_limit=3, output_window_cols=False)
    walk("PhysicalPlanNode", lim, seen)
    assert ("WindowGroupLimit", "k") in seen
    d = _explain(lim)["plan"]
    assert d["group_limit"] == 3 and d["output_window_cols"] is False and [f[0] for f in d["schema"]] == [f.name for f in T]
