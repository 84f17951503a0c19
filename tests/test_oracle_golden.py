"""Pin the CPU oracle against the golden vectors held by the reference's own unit tests.

Every vector below is restated from a `#[test]` in apache/auron (paths relative to
native-engine/); the oracle is only trusted as a checker because it reproduces them.
"""
import math

import numpy as np
import pyarrow as pa
import pytest

import oracle


def u32s(vals):
    return [v - (1 << 32) if v >= (1 << 31) else v for v in vals]


# datafusion-ext-commons/src/hash/mur.rs:94-103
def test_murmur3_raw_bytes():
    got = [oracle.murmur3_bytes(s.encode(), 42) for s in ["", "a", "ab", "abc", "abcd", "abcde"]]
    assert got == [142593372, 1485273170, -97053317, 1322437556, -396302900, 814637928]


# datafusion-ext-commons/src/hash/xxhash.rs:127-150
def test_xxhash64_raw_bytes():
    got = [oracle.xxhash64_bytes(s.encode(), 42)
           for s in ["", "a", "ab", "abc", "abcd", "abcde", "abcdefghijklmnopqrstuvwxyz"]]
    assert got == [-7444071767201028348, -8582455328737087284, 2710560539726725091, 1423657621850124518,
                   -6810745876291105281, -990457398947679591, -3265757659154784300]


# datafusion-ext-commons/src/spark_hash.rs:415-439 (test_i8)
def test_hash_i8():
    a = pa.array([1, 0, -1, 127, -128], type=pa.int8())
    assert oracle.hash_columns([a]).tolist() == u32s([0xdea578e3, 0x379fae8f, 0xa0590e3d, 0x43b4d8ed, 0x422a1365])


# spark_hash.rs:441-458 (test_i32)
def test_hash_i32():
    for v, e in [(1, -559580957), (2, 1765031574), (3, -1823081949), (4, -397064898)]:
        assert oracle.hash_columns([pa.array([v], type=pa.int32())]).tolist() == [e]


# spark_hash.rs:460-497 (test_i64)
def test_hash_i64():
    a = pa.array([1, 0, -1, 2**63 - 1, -2**63], type=pa.int64())
    assert oracle.hash_columns([a]).tolist() == u32s([0x99f0149d, 0x9c67b85d, 0xc8008529, 0xa05b5d7b, 0xcd1e64fb])
    assert oracle.hash_columns([a], "xxhash64").tolist() == [
        -7001672635703045582, -5252525462095825812, 3858142552250413010, -3246596055638297850, -8619748838626508300]


# spark_hash.rs:499-521 (test_str)
def test_hash_str():
    a = pa.array(["hello", "bar", "", "😁", "天地"])
    assert oracle.hash_columns([a]).tolist() == u32s([3286402344, 2486176763, 142593372, 885025535, 2395000894])
    assert oracle.hash_columns([a], "xxhash64").tolist() == [
        -4367754540140381902, -1798770879548125814, -7444071767201028348, -6337236088984028203, -235771157374669727]


def test_hash_null_leaves_seed_and_chains():
    # spark_hash.rs:52-57,78-84: NULL leaves the running hash unchanged; columns chain
    a = pa.array([1, None, 3], type=pa.int32())
    b = pa.array(["x", "y", None])
    h = oracle.hash_columns([a, b])
    assert h[1] == oracle.murmur3_bytes(b"y", 42)
    assert h[2] == oracle.hash_columns([pa.array([3], type=pa.int32())])[0]
    assert h[0] == oracle.murmur3_bytes(b"x", int(oracle.hash_columns([pa.array([1], type=pa.int32())])[0]))


def test_pmod():
    # shuffle/mod.rs:178-188 rem_euclid
    ids = oracle.partition_ids([pa.array([1, 2, 3, 4], type=pa.int32())], 7)
    assert ids.tolist() == [(-559580957) % 7, 1765031574 % 7, (-1823081949) % 7, (-397064898) % 7]


# datafusion-ext-commons/src/io/mod.rs:61-84
@pytest.mark.parametrize("n", [0, 1, 127, 128, 129, 16383, 16384, 10000, 2**31, 2**40 + 12345])
def test_varint_roundtrip(n):
    b = oracle.write_len(n)
    assert oracle.read_len(b)[0] == n
    if n < 128:
        assert b == bytes([n])
    if n == 128:
        assert b == bytes([128, 1])


# datafusion-ext-commons/src/io/batch_serde.rs:676-760 (round trip utf8/u64/bool with nulls, sliced)
def test_serde_roundtrip_with_nulls_and_slices():
    batch = pa.record_batch({
        "s": pa.array([None, "a", "bcd", None, "", "efghij", None, "k"]),
        "u": pa.array([None, 1, 2, None, 4, 5, None, 7], type=pa.int64()),
        "b": pa.array([None, True, False, None, True, True, None, False]),
        "d": pa.array([None, 1, -2, None, 12345678901234567890123, 0, None, -7], type=pa.decimal128(38, 2)),
    })
    for sl in (batch, batch.slice(1, 5), batch.slice(3, 0), batch.slice(2, 6)):
        raw = oracle.serde_write_batch(sl)
        back, pos = oracle.serde_read_batch(raw, sl.schema)
        assert pos == len(raw)
        assert back.equals(pa.record_batch(sl.to_pydict(), schema=sl.schema))


def test_serde_byte_plane_layout():
    # batch_serde.rs:292-305: values are byte-plane transposed (all byte-0s, then byte-1s, ...)
    batch = pa.record_batch({"a": pa.array([0x0102, 0x0304, 0x0506], type=pa.int32())})
    raw = oracle.serde_write_batch(batch)
    assert raw == bytes([3, 0, 0x02, 0x04, 0x06, 0x01, 0x03, 0x05, 0, 0, 0, 0, 0, 0])


# datafusion-ext-commons/src/arrow/cast.rs:553-579 (test_float_to_int)
def test_float_to_int_saturating():
    vals = [123.456, 987.654, 2147483647 + 10000.0, -2147483648 - 10000.0, math.inf, -math.inf, math.nan]
    assert [oracle.f64_to_int(v, 32) for v in vals] == [123, 987, 2147483647, -2147483648, 2147483647, -2147483648, 0]


# cast.rs:691-717 (test_string_to_bigint)
def test_string_to_bigint():
    ins = ["123", "987", "987.654", "123456789012345", "-123456789012345", "999999999999999999999999999999999"]
    assert [oracle.str_to_int(s.encode(), 64) for s in ins] == [123, 987, 987, 123456789012345, -123456789012345, None]
    assert oracle.str_to_int(b"-9223372036854775808", 64) == -2**63
    assert oracle.str_to_int(b"9223372036854775808", 64) is None
    assert oracle.str_to_int(b"+", 64) is None and oracle.str_to_int(b"", 64) is None
    assert oracle.str_to_int(b"12.x", 32) is None and oracle.str_to_int(b"128", 8) is None
    assert oracle.str_to_int(b"-128", 8) == -128 and oracle.str_to_int(b"127", 8) == 127


# cast.rs:719-752 (test_string_to_date)
def test_string_to_date():
    import datetime as dt
    epoch = dt.date(1970, 1, 1)
    exp = {"2001-02-03": dt.date(2001, 2, 3), "2001-03-04": dt.date(2001, 3, 4), "2001-04-05T06:07:08": dt.date(2001, 4, 5),
           "2001-04": dt.date(2001, 4, 1), "2002": dt.date(2002, 1, 1), "2001-00": None, "2001-13": None, "9999-99": None,
           "99999-01": None}
    for s, d in exp.items():
        got = oracle.str_to_date(s.encode())
        assert got == (None if d is None else (d - epoch).days), s
    assert oracle.str_to_date(b"2001-02-30") is None
    assert oracle.str_to_date(b" 2000-02-29 ") == (dt.date(2000, 2, 29) - epoch).days


# datafusion-ext-commons/src/algorithm/rdx_sort.rs:83-114 + shuffle/buffered_data.rs:285-353
def test_partition_rows_groups_by_partition():
    rng = np.random.default_rng(1)
    pid = rng.integers(0, 13, size=5000).astype(np.int32)
    rows, offs = oracle.partition_rows(pid, 13)
    assert offs[0] == 0 and offs[-1] == 5000
    for p in range(13):
        seg = rows[offs[p]:offs[p + 1]]
        assert (pid[seg] == p).all()
    assert sorted(rows.tolist()) == list(range(5000))


# datafusion-ext-plans/src/agg_exec.rs:716-843 (fuzz: SUM/COUNT vs hash map)
def test_agg_sum_count_vs_dict():
    rng = np.random.default_rng(7)
    n = 200_000
    k = rng.integers(-50, 5000, size=n)
    v = rng.integers(-10**6, 10**6, size=n)
    kmask = rng.random(n) < 0.01
    vmask = rng.random(n) < 0.02
    t = oracle.agg_sum_count_i64(pa.array(k, mask=kmask), pa.array(v, mask=vmask))
    exp = {}
    for ki, vi, km, vm in zip(k.tolist(), v.tolist(), kmask.tolist(), vmask.tolist()):
        key = None if km else ki
        s, c = exp.get(key, (0, 0))
        exp[key] = (s + (0 if vm else vi), c + (0 if vm else 1))
    got = {r["k"]: (r["sum"] if r["sum"] is not None else 0, r["cnt"]) for r in t.to_pylist()}
    assert got == exp


# datafusion-ext-functions/src/spark_round.rs:225-447 and spark_bround.rs:256-513
def test_spark_round_and_bround_goldens():
    R = oracle.spark_round
    assert [R(v, 1, "decimal", in_scale=2) for v in (12345, -67895)] == [12350, -67900]                       # round.rs:225-245
    assert [R(v, -1, "f64") for v in (123.45, -678.9)] == [120.0, -680.0]                                     # :250-264
    assert [R(v, 2, "f64") for v in (1.2345, -2.3456, 0.5, -0.5, None)] == [1.23, -2.35, 0.5, -0.5, None]     # :268-292
    assert R(-1.5, 0, "f64") == -2.0                                                                          # :297-306
    scales = range(-6, 7)
    assert [R(31415, s, "int16") for s in scales] == [0, 0, 30000, 31000, 31400, 31420] + [31415] * 7          # :309-327
    assert [R(314159265, s, "int32") for s in scales] == [314000000, 314200000, 314160000, 314159000, 314159300, 314159270] + [314159265] * 7
    pi = [0.0] * 6 + [3.0, 3.1, 3.14, 3.142, 3.1416, 3.14159, 3.141593]
    assert all(abs(R(math.pi, s, "f64") - e) < 1e-9 for s, e in zip(scales, pi))                              # :356-380
    long_pi = 31415926535897932
    assert [R(long_pi, s, "decimal") for s in range(-6, 1)] == [31415926536000000, 31415926535900000, 31415926535900000, 31415926535898000,
                                                               31415926535897900, 31415926535897930, long_pi]  # :410-447
    B = lambda *a, **k: oracle.spark_round(*a, half_even=True, **k)
    assert [B(v, 0, "f64") for v in (1.5, 2.5, -0.5, -1.5, 0.5)] == [2.0, 2.0, 0.0, -2.0, 0.0]                # bround.rs:256-279
    assert [B(v, -1, "f64") for v in (125.0, 135.0, 145.0, 155.0)] == [120.0, 140.0, 140.0, 160.0]            # :282-298
    assert [B(v, 1, "decimal", in_scale=2) for v in (12345, 67895)] == [12340, 67900]                         # :302-316
    assert [B(31415, s, "int16") for s in scales] == [0, 0, 30000, 31000, 31400, 31420] + [31415] * 7          # :377-401
    assert [B(314159265, s, "int32") for s in scales] == [314000000, 314200000, 314160000, 314159000, 314159300, 314159260] + [314159265] * 7
    assert [B(long_pi, s, "decimal") for s in range(-6, 1)] == [31415926536000000, 31415926535900000, 31415926535900000, 31415926535898000,
                                                               31415926535897900, 31415926535897930, long_pi]  # :433-464
    assert [(B(v, s, "f64")) for v, s in ((2.5, 0), (3.5, 0), (-2.5, 0), (-3.5, 0), (-0.35, 1), (-35.0, -1))] == [2.0, 4.0, -2.0, -4.0, -0.4, -40.0]
    assert all(abs(B(math.pi, s, "f64") - e) < 1e-9 for s, e in zip(scales, pi))                              # :325-348


# datafusion-ext-functions/src/spark_dates.rs:661-952, 1077-1153
def test_spark_time_parts_goldens():
    import datetime as dt
    T = oracle.spark_time_part
    utc_ms = lambda *a: int(dt.datetime(*a, tzinfo=dt.timezone.utc).timestamp() * 1000)
    hms = (1 * 3600 + 23 * 60 + 45) * 1000
    assert [T(v, "ms", "hour") for v in (0, hms, None)] == [0, 1, None]                                       # :661-688
    assert [T(v, "ms", "minute") for v in (0, hms)] == [0, 23] and [T(v, "ms", "second") for v in (0, hms)] == [0, 45]
    assert [T(d, "date32", w) for d in (0, 1) for w in ("hour", "minute", "second")] == [0] * 6               # :691-711
    assert (T(-1000, "ms", "hour"), T(-1000, "ms", "minute"), T(-1000, "ms", "second")) == (23, 59, 59)       # :746-764
    assert T(0, "ms", "hour", "Asia/Shanghai") == 8                                                           # :784-802
    assert (T(0, "ms", "minute", "Asia/Kolkata"), T(0, "ms", "second", "Asia/Kolkata")) == (30, 0)            # :835-881
    assert T(0, "ms", "minute", "Asia/Kathmandu") == 30                                                       # :884-904
    t1, t2 = utc_ms(2019, 3, 10, 6, 59, 59), utc_ms(2019, 3, 10, 7, 0, 0)                                     # :907-937
    assert [T(t, "ms", w, "America/New_York") for t in (t1, t2) for w in ("minute", "second")] == [59, 59, 0, 0]
    e = utc_ms(2021, 1, 4, 4, 30, 0)                                                                          # :1077-1110
    assert [T(e, "ms", w, "America/New_York") for w in ("year", "month", "day", "dayofweek")] == [2021, 1, 3, 1]
    assert T(utc_ms(2021, 4, 1, 3, 0, 0), "ms", "quarter", "America/New_York") == 1                           # :1113-1123
    e = utc_ms(2021, 12, 31, 17, 0, 0)                                                                        # :1126-1153
    assert [T(e, "ms", w, "Asia/Shanghai") for w in ("year", "month", "day", "quarter", "dayofweek")] == [2022, 1, 1, 1, 7]
    assert T(0, "ms", "hour", "Mars/Olympus") == 0                                                            # unknown zone = none
