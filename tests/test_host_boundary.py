"""CPU-only checks of the boundary: the C-ABI library loads and exports every symbol declared in
include/auron_b200.h, the plan encoder produces well-formed protobuf, and the product refuses to run
without a CUDA device (no CPU fallback)."""
import os
import re

import pyarrow as pa
import pytest

from auron_b200 import proto as P
from auron_b200 import runtime

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "auron_b200.h")).read()
    return sorted(set(re.findall(r"\b(auron_b200_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    L = runtime.lib()
    declared = _declared_symbols()
    assert len(declared) >= 13
    for sym in declared:
        assert hasattr(L, sym), f"{sym} declared in include/auron_b200.h but not exported"
    assert set(runtime.EXPORTED_SYMBOLS) <= set(declared)


def test_jni_symbols_exported():
    # same names as native-engine/auron/src/exec.rs:42,122,133,144
    import ctypes
    L = ctypes.CDLL(runtime.LIB_PATH)
    for s in ("callNative", "nextBatch", "finalizeNative", "onExit"):
        assert hasattr(L, f"Java_org_apache_auron_jni_JniBridge_{s}")


def test_jni_natives_against_the_mock_jvm_without_a_gpu():
    # The JNI natives run against the mock JVM (tests/jni_mock/mock_jvm.cc): callNative fetches the task definition through
    # getRawTaskDefinition, and -- no GPU here, no CPU fallback -- fails by raising RuntimeException on the calling thread
    # (exec.rs:42-118 throws from callNative the same way), with every JNI reference released again.
    import torch
    if torch.cuda.is_available():
        pytest.skip("the GPU suite (test_gpu_jni.py) covers the same entry points with a device")
    from jni_helpers import MockJvm
    jvm = MockJvm(P.task_definition(P.ffi_reader(pa.schema([("a", pa.int32())]), "in")))
    assert not jvm.call_native()
    assert jvm.pending_exception() == "java/lang/RuntimeException: auron_b200 requires a CUDA device (no CPU fallback)"
    jvm.assert_clean()


@pytest.mark.parametrize("compression,version,dict_", [("NONE", "1.0", True), ("SNAPPY", "1.0", True), ("SNAPPY", "2.0", True), ("SNAPPY", "1.0", False),
                                                       ("ZSTD", "2.0", True)])
def test_parquet_metadata_reader_against_pyarrow(tmp_path, compression, version, dict_):
    # P1's host half (parquet_meta.cc: Thrift compact protocol footer + page headers, Snappy block decoder) pinned on the CPU against
    # Arrow C++'s reader of the same file: every footer field the scan uses, the page chain of every chunk tiling it exactly, page
    # value counts adding up, and every SNAPPY page body decoding to its declared size with the engine's own decoder
    import ctypes as C
    import decimal
    import json
    import numpy as np
    import pyarrow.parquet as pq
    rng = np.random.default_rng(3)
    n = 60_000
    t = pa.table({"i": pa.array(rng.integers(0, 500, n), type=pa.int32(), mask=rng.random(n) < 0.05),
                  "l": pa.array(rng.integers(-2**40, 2**40, n), type=pa.int64()),
                  "s": pa.array([f"v{int(x) % 300}" for x in rng.integers(0, 10**6, n)], mask=rng.random(n) < 0.1),
                  "d": pa.array([decimal.Decimal(int(x)) / 100 for x in rng.integers(0, 10**6, n)], type=pa.decimal128(7, 2)),
                  "f": pa.array(rng.standard_normal(n)), "b": pa.array(rng.random(n) < 0.5),
                  "req": pa.array(np.arange(n, dtype=np.int64))},
                 schema=pa.schema([("i", pa.int32()), ("l", pa.int64()), ("s", pa.string()), ("d", pa.decimal128(7, 2)), ("f", pa.float64()),
                                   ("b", pa.bool_()), pa.field("req", pa.int64(), nullable=False)]))
    path = str(tmp_path / "m.parquet")
    pq.write_table(t, path, compression=compression, data_page_version=version, use_dictionary=dict_, row_group_size=25_000, data_page_size=16 * 1024,
                   store_decimal_as_integer=True)
    L = runtime.lib()
    L.auron_b200_parquet_describe.restype = C.c_int64
    L.auron_b200_parquet_describe.argtypes = [C.c_char_p, C.c_char_p, C.c_int64]
    buf = C.create_string_buffer(4 << 20)
    size = L.auron_b200_parquet_describe(path.encode(), buf, len(buf))
    assert size > 0, buf.value
    d = json.loads(buf.value.decode())
    md = pq.ParquetFile(path).metadata
    assert d["num_rows"] == md.num_rows == n and len(d["row_groups"]) == md.num_row_groups == 3
    assert d["created_by"] == md.created_by
    leaves = [e for e in d["schema"] if e["num_children"] == 0]
    assert [e["name"] for e in leaves] == t.column_names
    phys = {"BOOLEAN": 0, "INT32": 1, "INT64": 2, "INT96": 3, "FLOAT": 4, "DOUBLE": 5, "BYTE_ARRAY": 6, "FIXED_LEN_BYTE_ARRAY": 7}
    codecs = {"UNCOMPRESSED": 0, "SNAPPY": 1, "GZIP": 2, "LZ4": 5, "ZSTD": 6, "LZ4_RAW": 7}
    assert [e["repetition"] for e in leaves] == [1] * 6 + [0]                       # OPTIONAL ... REQUIRED
    assert (leaves[3]["precision"], leaves[3]["scale"]) == (7, 2)                    # decimal(7,2) stored as INT32
    snappy_pages = 0
    for g, rg in enumerate(d["row_groups"]):
        ref_rg = md.row_group(g)
        assert rg["num_rows"] == ref_rg.num_rows
        for c, cm in enumerate(rg["columns"]):
            ref = ref_rg.column(c)
            assert cm["path"] == ref.path_in_schema
            assert cm["type"] == phys[ref.physical_type] and cm["codec"] == codecs[ref.compression]
            assert cm["num_values"] == ref.num_values == cm["page_values"]           # the data pages add up to the chunk
            assert cm["total_compressed"] == ref.total_compressed_size and cm["total_uncompressed"] == ref.total_uncompressed_size
            assert cm["pages_uncompressed"] == ref.total_uncompressed_size           # page headers + uncompressed bodies
            assert cm["data_page_offset"] == ref.data_page_offset
            assert (cm["dictionary_page_offset"] > 0) == ref.has_dictionary_page == (cm["dictionary_pages"] == 1)
            assert cm["data_pages"] >= 1
            st = ref.statistics
            if st is not None and st.has_null_count:
                assert cm["null_count"] == st.null_count
            if st is not None and st.has_min_max and ref.physical_type in ("INT32", "INT64"):
                width = 4 if ref.physical_type == "INT32" else 8
                lo, hi = (int.from_bytes(bytes.fromhex(cm[k]), "little", signed=True) for k in ("min", "max"))
                assert len(cm["min"]) == 2 * width
                if ref.path_in_schema == "d":
                    assert (decimal.Decimal(lo) / 100, decimal.Decimal(hi) / 100) == (st.min, st.max)
                else:
                    assert (lo, hi) == (st.min, st.max)
            if ref.path_in_schema == "s" and st is not None and st.has_min_max:
                assert bytes.fromhex(cm["min"]).decode() == st.min and bytes.fromhex(cm["max"]).decode() == st.max
            snappy_pages += cm["snappy_pages_decompressed"]
    assert (snappy_pages > 0) == (compression == "SNAPPY")
    # a damaged footer is an error message, not a crash
    raw = bytearray(open(path, "rb").read())
    raw[-30] ^= 0xFF
    raw[-40] ^= 0xFF
    bad = str(tmp_path / "bad.parquet")
    open(bad, "wb").write(bytes(raw))
    rc = L.auron_b200_parquet_describe(bad.encode(), buf, len(buf))
    assert rc == -1 or rc > 0                                                      # rejected with a message, or parsed: never a crash
    assert L.auron_b200_parquet_describe(str(tmp_path / "missing.parquet").encode(), buf, len(buf)) == -1 and b"cannot open" in buf.value


def test_parquet_metadata_reader_survives_damaged_files(tmp_path):
    # footers, page headers and Snappy bodies with flipped bytes: the reader answers with an error or a description, never with a crash
    # (a corrupt file must fail one task, not the executor)
    import ctypes as C
    import random
    import numpy as np
    import pyarrow.parquet as pq
    rng = np.random.default_rng(1)
    n = 20_000
    t = pa.table({"i": pa.array(rng.integers(0, 500, n), type=pa.int32(), mask=rng.random(n) < 0.05),
                  "s": pa.array([f"v{int(x) % 300}" for x in rng.integers(0, 10**6, n)]), "f": pa.array(rng.standard_normal(n))})
    L = runtime.lib()
    L.auron_b200_parquet_describe.restype = C.c_int64
    L.auron_b200_parquet_describe.argtypes = [C.c_char_p, C.c_char_p, C.c_int64]
    buf = C.create_string_buffer(1 << 22)
    random.seed(99)
    ok = bad = 0
    for comp, ver in (("SNAPPY", "1.0"), ("SNAPPY", "2.0"), ("ZSTD", "2.0")):
        src = str(tmp_path / f"{comp}{ver}.parquet")
        pq.write_table(t, src, compression=comp, data_page_version=ver, row_group_size=5000, data_page_size=4096)
        raw = open(src, "rb").read()
        mut = str(tmp_path / "mut.parquet")
        for _ in range(300):
            b = bytearray(raw)
            for _ in range(random.randint(1, 4)):
                r = random.random()
                pos = len(b) - 8 - random.randint(1, 900) if r < 0.5 else (random.randint(4, len(b) - 9) if r < 0.9 else len(b) - random.randint(1, 8))
                b[pos] = random.randint(0, 255)
            open(mut, "wb").write(bytes(b))
            rc = L.auron_b200_parquet_describe(mut.encode(), buf, len(buf))
            ok += rc > 0
            bad += rc < 0
    assert ok > 50 and bad > 50


def test_delta_and_split_walkers_survive_damaged_page_bodies(tmp_path):
    # the scan's host-side page walkers (DELTA_BINARY_PACKED / DELTA_LENGTH_BYTE_ARRAY / DELTA_BYTE_ARRAY decoders, the Snappy tag walk)
    # on files whose PAGE BODIES are damaged: every length, offset and bit width comes from untrusted bytes, so the answer must be an
    # error or a description -- never a crash, a hang or an allocation of the advertised size
    import ctypes as C
    import random
    import numpy as np
    import pyarrow.parquet as pq
    rng = np.random.default_rng(2)
    n = 30_000
    t = pa.table({"a": pa.array(np.sort(rng.integers(-2**31, 2**31 - 1, n)).astype(np.int32), mask=rng.random(n) < 0.03),
                  "b": pa.array(rng.integers(-2**63, 2**63 - 1, n, dtype=np.int64)),
                  "dl": pa.array([f"w{int(i) % 977}" for i in rng.integers(0, 10**6, n)], mask=rng.random(n) < 0.05),
                  "db": pa.array(sorted(f"key-{int(x):09d}" for x in rng.integers(0, 10**9, n))),
                  "noise": pa.array(rng.integers(-2**31, 2**31 - 1, n).astype(np.int32), mask=rng.random(n) < 0.05)})
    enc = {"a": "DELTA_BINARY_PACKED", "b": "DELTA_BINARY_PACKED", "dl": "DELTA_LENGTH_BYTE_ARRAY", "db": "DELTA_BYTE_ARRAY", "noise": "PLAIN"}
    L = runtime.lib()
    L.auron_b200_parquet_describe.restype = C.c_int64
    L.auron_b200_parquet_describe.argtypes = [C.c_char_p, C.c_char_p, C.c_int64]
    buf = C.create_string_buffer(1 << 22)
    random.seed(7)
    ok = bad = 0
    for comp, ver in (("NONE", "1.0"), ("NONE", "2.0"), ("SNAPPY", "1.0"), ("SNAPPY", "2.0")):
        src = str(tmp_path / f"{comp}{ver}.parquet")
        pq.write_table(t, src, compression=comp, data_page_version=ver, use_dictionary=False, column_encoding=enc, row_group_size=n, data_page_size=256 * 1024)
        raw = open(src, "rb").read()
        md = pq.ParquetFile(src).metadata.row_group(0)
        spans = [(md.column(c).data_page_offset, md.column(c).data_page_offset + md.column(c).total_compressed_size) for c in range(md.num_columns)]
        mut = str(tmp_path / "mut.parquet")
        for _ in range(250):
            b = bytearray(raw)
            for _ in range(random.randint(1, 3)):
                lo, hi = random.choice(spans)
                r = random.random()
                pos = random.randint(lo, min(hi - 1, lo + 200)) if r < 0.5 else random.randint(lo, hi - 1)   # page / stream headers sit at the front
                b[pos] = random.randint(0, 255)
            open(mut, "wb").write(bytes(b))
            rc = L.auron_b200_parquet_describe(mut.encode(), buf, len(buf))
            ok += rc > 0
            bad += rc < 0
    assert ok > 50 and bad > 50


def _jni_function_table():
    """JNINativeInterface_ in declaration order, generated from the structure of the JNI specification's "Interface Function
    Table" (4 reserved slots, then the functions; the Call*Method families come as plain / V / A triples per result type)."""
    types = ["Object", "Boolean", "Byte", "Char", "Short", "Int", "Long", "Float", "Double"]
    prims = types[1:]
    t = ["reserved0", "reserved1", "reserved2", "reserved3", "GetVersion", "DefineClass", "FindClass", "FromReflectedMethod", "FromReflectedField",
         "ToReflectedMethod", "GetSuperclass", "IsAssignableFrom", "ToReflectedField", "Throw", "ThrowNew", "ExceptionOccurred", "ExceptionDescribe",
         "ExceptionClear", "FatalError", "PushLocalFrame", "PopLocalFrame", "NewGlobalRef", "DeleteGlobalRef", "DeleteLocalRef", "IsSameObject",
         "NewLocalRef", "EnsureLocalCapacity", "AllocObject", "NewObject", "NewObjectV", "NewObjectA", "GetObjectClass", "IsInstanceOf", "GetMethodID"]
    calls = lambda prefix: [f"{prefix}{ty}Method{suffix}" for ty in types + ["Void"] for suffix in ("", "V", "A")]
    t += calls("Call") + calls("CallNonvirtual")
    t += ["GetFieldID"] + [f"Get{ty}Field" for ty in types] + [f"Set{ty}Field" for ty in types]
    t += ["GetStaticMethodID"] + calls("CallStatic")
    t += ["GetStaticFieldID"] + [f"GetStatic{ty}Field" for ty in types] + [f"SetStatic{ty}Field" for ty in types]
    t += ["NewString", "GetStringLength", "GetStringChars", "ReleaseStringChars", "NewStringUTF", "GetStringUTFLength", "GetStringUTFChars",
          "ReleaseStringUTFChars", "GetArrayLength", "NewObjectArray", "GetObjectArrayElement", "SetObjectArrayElement"]
    t += [f"New{p}Array" for p in prims] + [f"Get{p}ArrayElements" for p in prims] + [f"Release{p}ArrayElements" for p in prims]
    t += [f"Get{p}ArrayRegion" for p in prims] + [f"Set{p}ArrayRegion" for p in prims]
    t += ["RegisterNatives", "UnregisterNatives", "MonitorEnter", "MonitorExit", "GetJavaVM", "GetStringRegion", "GetStringUTFRegion",
          "GetPrimitiveArrayCritical", "ReleasePrimitiveArrayCritical", "GetStringCritical", "ReleaseStringCritical", "NewWeakGlobalRef",
          "DeleteWeakGlobalRef", "ExceptionCheck", "NewDirectByteBuffer", "GetDirectBufferAddress", "GetDirectBufferCapacity", "GetObjectRefType",
          "GetModule"]
    return {name: i for i, name in enumerate(t)}


def test_jni_function_indices_follow_the_specification_table():
    # no jni.h in this image: jni_face.cc (product) and tests/jni_mock/mock_jvm.cc address JNIEnv functions by slot number.  Both are
    # checked here against the table generated from the specification's declaration order, so that a wrong slot cannot hide in an
    # agreement between the two files.
    table = _jni_function_table()
    assert len(table) == 234 and table["GetVersion"] == 4 and table["ExceptionCheck"] == 228 and table["GetModule"] == 233
    face = open(os.path.join(ROOT, "auron_b200", "csrc", "jni_face.cc")).read()
    used = re.findall(r"\bFN_(\w+) = (\d+)", face)
    assert len(used) >= 30
    for name, slot in used:
        assert table[name] == int(slot), f"jni_face.cc: {name} is slot {table[name]}, not {slot}"
    vm = dict(re.findall(r"\bVM_(\w+) = (\d+)", face))
    assert vm == {"DetachCurrentThread": "5", "GetEnv": "6", "AttachCurrentThreadAsDaemon": "7"}   # JNIInvokeInterface_: 3 reserved, Destroy, Attach, ...
    mock = open(os.path.join(ROOT, "tests", "jni_mock", "mock_jvm.cc")).read()
    slots = re.findall(r"g_table\[(\d+)\] = \(void\*\)f_(\w+);", mock)
    assert len(slots) >= 30
    for slot, name in slots:
        assert table[name] == int(slot), f"mock_jvm.cc: {name} is slot {table[name]}, not {slot}"
    assert {name for name, _ in used} == {name for _, name in slots}      # the mock provides exactly what the product uses


def test_time_zone_tables_match_python_zoneinfo():
    # the UTC-offset tables the device searches (tzdb.cc: TZif transitions + the footer's POSIX rule expanded to 2200) against
    # Python's independent reader of the same tz database, 1875 .. 2191
    import ctypes as C
    import datetime as dt
    import random
    import zoneinfo
    L = runtime.lib()
    L.auron_b200_tz_offset.argtypes = [C.c_char_p, C.c_int64, C.POINTER(C.c_int32)]
    random.seed(7)
    for zone in ["UTC", "America/New_York", "Asia/Shanghai", "Asia/Kolkata", "Asia/Kathmandu", "Europe/Dublin", "Australia/Lord_Howe",
                 "America/Sao_Paulo", "Pacific/Apia", "Antarctica/Troll", "Etc/GMT+8", "America/St_Johns"]:
        tz = zoneinfo.ZoneInfo(zone)
        for _ in range(1500):
            s = random.randint(-3_000_000_000, 7_000_000_000)
            exp = dt.datetime.fromtimestamp(s, tz=dt.timezone.utc).astimezone(tz).utcoffset().total_seconds()
            got = C.c_int32()
            assert L.auron_b200_tz_offset(zone.encode(), s, C.byref(got)) == 0
            assert got.value == exp, (zone, s)
    for bad in ["Mars/Olympus", "../etc/passwd", "+08:00", "", "posix/UTC"]:
        assert L.auron_b200_tz_offset(bad.encode(), 0, C.byref(C.c_int32())) == -1


def _parse(buf: bytes):
    """tiny generic proto reader: [(field, wire, value)]"""
    out, i = [], 0

    def varint():
        nonlocal i
        v = s = 0
        while True:
            b = buf[i]
            i += 1
            v |= (b & 0x7F) << s
            if not b & 0x80:
                return v
            s += 7

    while i < len(buf):
        t = varint()
        f, w = t >> 3, t & 7
        if w == 0:
            out.append((f, w, varint()))
        elif w == 2:
            n = varint()
            out.append((f, w, buf[i:i + n]))
            i += n
        else:
            raise AssertionError("unexpected wire type")
    return out


def test_plan_encoding_roundtrip():
    s = pa.schema([("a", pa.int64()), ("s", pa.string()), ("d", pa.decimal128(7, 2))])
    plan = P.filter_(P.ffi_reader(s, "rid"), [P.binary("Gt", P.col("a"), P.lit(5, pa.int64()))])
    td = P.task_definition(plan, stage_id=3, partition_id=4, task_id=5)
    top = _parse(td)
    assert [f for f, _, _ in top] == [1, 2]
    assert _parse(top[0][2]) == [(2, 0, 3), (4, 0, 4), (5, 0, 5)]       # PartitionId{stage_id=2, partition_id=4, task_id=5}
    node = _parse(top[1][2])
    assert node[0][0] == 8                                               # PhysicalPlanNode.filter
    flt = _parse(node[0][2])
    assert [f for f, _, _ in flt] == [1, 2]
    ffi = _parse(_parse(flt[0][2])[0][2])
    assert ffi[2] == (3, 2, b"rid")
    fields = [_parse(x[2]) for x in _parse(ffi[1][2])]
    assert [f[0][2] for f in fields] == [b"a", b"s", b"d"]
    assert _parse(fields[2][1][2])[0][0] == 24                           # ArrowType.DECIMAL
    # literal = Arrow IPC stream (schema + 1-row batch)
    expr = _parse(flt[1][2])
    assert expr[0][0] == 4
    lit = _parse(_parse(expr[0][2])[1][2])
    ipc = _parse(lit[0][2])[0][2]
    assert pa.ipc.open_stream(ipc).read_all().column(0).to_pylist() == [5]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    t = pa.table({"a": pa.array([1, 2, 3], type=pa.int64())})
    with pytest.raises(runtime.AuronError):
        runtime.run_task(P.task_definition(P.ffi_reader(t.schema, "t")), {"t": t.to_batches()})
    with pytest.raises(runtime.AuronError):
        runtime.k_partition_ids(t.to_batches()[0], [0], 8)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "auron_b200")):
        for f in files:
            if f.endswith((".py", ".cc", ".cu", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in src and "auron_oracle" not in src, f


def _describe(path):
    import ctypes as C
    import json
    L = runtime.lib()
    L.auron_b200_parquet_describe.restype = C.c_int64
    L.auron_b200_parquet_describe.argtypes = [C.c_char_p, C.c_char_p, C.c_int64]
    buf = C.create_string_buffer(8 << 20)
    assert L.auron_b200_parquet_describe(path.encode(), buf, len(buf)) > 0, buf.value
    return json.loads(buf.value.decode())


@pytest.mark.parametrize("compression", ["NONE", "SNAPPY"])
@pytest.mark.parametrize("version", ["1.0", "2.0"])
def test_delta_decoders_of_the_scan_against_arrow_on_the_cpu(tmp_path, compression, version):
    # the host decoders the scan uses for DELTA_LENGTH_BYTE_ARRAY / DELTA_BYTE_ARRAY pages (rewritten as PLAIN before upload) and the
    # host restatement of DELTA_BINARY_PACKED (the device kernel walks the same block structure): every value of every page, checked
    # through count, wrapping sum (integers) and byte total + FNV-1a (strings) against the columns as Arrow C++ reads them back
    import numpy as np
    import pyarrow.parquet as pq
    rng = np.random.default_rng(5)
    n = 70_001
    t = pa.table({"sorted32": pa.array(np.sort(rng.integers(-2**31, 2**31 - 1, n)).astype(np.int32), mask=rng.random(n) < 0.03),
                  "noise64": pa.array(rng.integers(-2**63, 2**63 - 1, n, dtype=np.int64), mask=rng.random(n) < 0.02),
                  "req64": pa.array(np.cumsum(rng.integers(0, 1000, n)).astype(np.int64)),
                  "dl": pa.array([f"w{int(i) * 37 % 1013}" for i in rng.integers(0, 10**6, n)], mask=rng.random(n) < 0.05),
                  "db": pa.array(sorted(f"key-{int(x):09d}" for x in rng.integers(0, 10**9, n)))},
                 schema=pa.schema([("sorted32", pa.int32()), ("noise64", pa.int64()), pa.field("req64", pa.int64(), nullable=False), ("dl", pa.string()), ("db", pa.string())]))
    enc = {"sorted32": "DELTA_BINARY_PACKED", "noise64": "DELTA_BINARY_PACKED", "req64": "DELTA_BINARY_PACKED", "dl": "DELTA_LENGTH_BYTE_ARRAY", "db": "DELTA_BYTE_ARRAY"}
    path = str(tmp_path / "delta.parquet")
    pq.write_table(t, path, compression=compression, use_dictionary=False, column_encoding=enc, data_page_version=version, row_group_size=30_000, data_page_size=32 * 1024)
    d = _describe(path)
    back = pq.read_table(path)
    got = {name: [0, 0, 0, 1469598103934665603, 0] for name in t.column_names}     # values, sum, bytes, fnv, pages
    for rg in d["row_groups"]:
        for cm in rg["columns"]:
            g = got[cm["path"]]
            g[0] += cm["delta_values"]
            g[1] = (g[1] + int(cm["delta_sum"])) % 2**64
            g[2] += cm["delta_string_bytes"]
            g[4] += cm["delta_pages"]
            assert cm["delta_pages"] == cm["data_pages"] >= 1
    for name in ("sorted32", "noise64", "req64"):
        col = back[name].combine_chunks()
        vals = col.drop_null().to_numpy(zero_copy_only=False).astype(np.int64)
        assert got[name][0] == len(vals) and got[name][1] == int(vals.astype(np.uint64).sum(dtype=np.uint64)), name
    for name in ("dl", "db"):
        vals = [v.encode() for v in back[name].to_pylist() if v is not None]
        assert got[name][0] == len(vals) and got[name][2] == sum(len(v) for v in vals), name
    # the per-chunk FNV values chain per column only inside one chunk: check them chunk by chunk
    for g_i, rg in enumerate(d["row_groups"]):
        lo = sum(x["num_rows"] for x in d["row_groups"][:g_i])
        for cm in rg["columns"]:
            if cm["path"] not in ("dl", "db"):
                continue
            h = 1469598103934665603
            for v in back[cm["path"]].slice(lo, rg["num_rows"]).to_pylist():
                if v is None:
                    continue
                for b in v.encode():
                    h = ((h ^ b) * 1099511628211) % 2**64
                h = ((h ^ 0xFF) * 1099511628211) % 2**64
            assert int(cm["delta_fnv"]) == h, cm["path"]


def test_snappy_bodies_are_split_into_head_and_literal_pieces_on_the_cpu(tmp_path):
    # pq::snappy_split (the tag walk the scan uses to turn the literal chain behind the last back reference of a page body into
    # independent stored-copy jobs): auron_b200_parquet_describe checks every split against the fully decompressed body; here the
    # shapes -- incompressible 1 MB pages split (a short head for the level bytes), compressible pages do not
    import numpy as np
    import pyarrow.parquet as pq
    rng = np.random.default_rng(8)
    n = 600_000
    t = pa.table({"noise": pa.array(rng.integers(-2**31, 2**31 - 1, n).astype(np.int32), mask=rng.random(n) < 0.05),
                  "runs": pa.array(np.repeat(rng.integers(0, 50, n // 1000 + 1), 1000)[:n].astype(np.int32))})
    path = str(tmp_path / "split.parquet")
    pq.write_table(t, path, compression="SNAPPY", use_dictionary=False, row_group_size=n, write_page_index=False)
    d = _describe(path)
    noise, runs = d["row_groups"][0]["columns"]
    assert noise["snappy_split_pages"] == noise["data_pages"] >= 1
    assert noise["snappy_split_stored_bytes"] > 0.95 * 4 * n * 0.95 and noise["snappy_split_head_bytes"] < 0.05 * 4 * n
    assert runs["snappy_split_pages"] == 0 or runs["snappy_split_stored_bytes"] < 0.2 * 4 * n
