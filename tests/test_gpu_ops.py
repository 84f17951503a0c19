"""GPU parity tests: every case runs the CUDA path through the C ABI (libauron_b200.so) and compares it
with the oracle / the reference's own golden tables.  Golden tables are restated from the reference's
unit tests (file:line cited per test, paths relative to native-engine/)."""
import datetime as dt
import decimal
import math
import os

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

import oracle
from auron_b200 import proto as P
from auron_b200 import runtime
from helpers import assert_same_rows, batches, canon, i32, run, table_i32

pytestmark = pytest.mark.gpu


# =============================================================================== hashing (S3)
def test_hash_golden_vectors():
    # datafusion-ext-commons/src/spark_hash.rs:415-520
    b = pa.record_batch({"a": pa.array([1, 0, -1, 127, -128], type=pa.int8())})
    assert runtime.k_hash(b, [0]).to_pylist() == [v - (1 << 32) if v >= 1 << 31 else v for v in
                                                  [0xdea578e3, 0x379fae8f, 0xa0590e3d, 0x43b4d8ed, 0x422a1365]]
    b = pa.record_batch({"a": pa.array([1, 2, 3, 4], type=pa.int32())})
    assert runtime.k_hash(b, [0]).to_pylist() == [-559580957, 1765031574, -1823081949, -397064898]
    b = pa.record_batch({"a": pa.array([1, 0, -1, 2**63 - 1, -2**63], type=pa.int64())})
    assert runtime.k_hash(b, [0], "xxhash64").to_pylist() == [-7001672635703045582, -5252525462095825812, 3858142552250413010,
                                                              -3246596055638297850, -8619748838626508300]
    b = pa.record_batch({"s": pa.array(["hello", "bar", "", "😁", "天地"])})
    assert runtime.k_hash(b, [0]).to_pylist() == [v - (1 << 32) if v >= 1 << 31 else v for v in
                                                  [3286402344, 2486176763, 142593372, 885025535, 2395000894]]
    assert runtime.k_hash(b, [0], "xxhash64").to_pylist() == [-4367754540140381902, -1798770879548125814, -7444071767201028348,
                                                              -6337236088984028203, -235771157374669727]


def _random_table(n, seed, nulls=0.05):
    rng = np.random.default_rng(seed)

    def m():
        return rng.random(n) < nulls

    words = ["", "a", "ab", "abc", "abcd", "hello world", "天地玄黄", "x" * 37, "😁", "spark-b200"]
    return pa.table({
        "i8": pa.array(rng.integers(-128, 128, n), type=pa.int8(), mask=m()),
        "i16": pa.array(rng.integers(-2**15, 2**15, n), type=pa.int16(), mask=m()),
        "i32": pa.array(rng.integers(-2**31, 2**31, n), type=pa.int32(), mask=m()),
        "i64": pa.array(rng.integers(-2**63, 2**63 - 1, n), type=pa.int64(), mask=m()),
        "f32": pa.array(rng.standard_normal(n).astype(np.float32), mask=m()),
        "f64": pa.array(rng.standard_normal(n), mask=m()),
        "b": pa.array(rng.random(n) < 0.5, mask=m()),
        "s": pa.array([None if x else words[int(k)] for x, k in zip(m(), rng.integers(0, len(words), n))]),
        "d32": pa.array(rng.integers(-30000, 60000, n).astype(np.int32), type=pa.date32(), mask=m()),
        "dec": pa.array([None if x else decimal.Decimal(int(v)) / 100 for x, v in zip(m(), rng.integers(-10**12, 10**12, n))],
                        type=pa.decimal128(17, 2)),
    })


@pytest.mark.parametrize("n", [1, 31, 1000, 100_003])
def test_hash_random_all_types(n):
    t = _random_table(n, seed=n)
    b = t.to_batches()[0]
    for kind in ("murmur3", "xxhash64"):
        for cols in ([0], [1], [2], [3], [4], [5], [6], [7], [8], [9], [2, 7, 9], list(range(10))):
            got = runtime.k_hash(b, cols, kind).to_numpy()
            exp = oracle.hash_columns([t.column(c).combine_chunks() for c in cols], kind)
            assert (got == exp).all(), (kind, cols)


def test_partition_ids_and_empty():
    t = _random_table(50_000, seed=3)
    b = t.to_batches()[0]
    for nparts in (1, 7, 200, 4096):
        got = runtime.k_partition_ids(b, [3, 7], nparts).to_numpy()
        exp = oracle.partition_ids([t["i64"].combine_chunks(), t["s"].combine_chunks()], nparts)
        assert (got == exp).all()
        assert got.min() >= 0 and got.max() < nparts
    empty = t.slice(0, 0).combine_chunks().to_batches() or [pa.RecordBatch.from_pydict({"i64": pa.array([], type=pa.int64())})]
    assert len(runtime.k_hash(pa.record_batch({"a": pa.array([], type=pa.int64())}), [0])) == 0


# =============================================================================== filter / project (F1-F3, E1-E5)
def _cfg1_table(n=200_000, seed=42):
    rng = np.random.default_rng(seed)
    vocab = ["".join(chr(97 + int(c)) for c in rng.integers(0, 26, int(l))) for l in rng.integers(4, 25, 1000)]
    a = rng.integers(0, 1_000_000, n)
    return pa.table({
        "a": pa.array(a, type=pa.int64(), mask=rng.random(n) < 0.01),
        "s": pa.array([vocab[int(i)] for i in rng.integers(0, 1000, n)], mask=rng.random(n) < 0.01),
    })


def test_filter_project_config1():
    # BASELINE config 1: Project[a+1, substr(s,1,4)] <- Filter[a > 500000 AND s LIKE 'ab%']
    t = _cfg1_table()
    src = P.ffi_reader(t.schema, "t")
    flt = P.filter_(src, [P.binary("Gt", P.col("a"), P.lit(500000, pa.int64())), P.like(P.col("s"), P.lit("a%", pa.string()))])
    plan = P.projection(flt, [P.binary("Plus", P.col("a"), P.lit(1, pa.int64())),
                              P.scalar_fn("Substr", [P.col("s"), P.lit(1, pa.int64()), P.lit(4, pa.int64())], pa.string())],
                        ["a1", "s4"], [pa.int64(), pa.string()])
    got = run(plan, {"t": t}, chunk=10_000)
    mask = pc.and_kleene(pc.greater(t["a"], 500000), pc.match_like(t["s"], "a%"))
    ft = t.filter(mask)   # null mask -> dropped
    exp = pa.table({"a1": pc.add(ft["a"], 1), "s4": pc.utf8_slice_codeunits(ft["s"], 0, 4)})
    assert got.schema.names == ["a1", "s4"]
    assert got.num_rows == exp.num_rows and got.num_rows > 0
    assert got.column(0).to_pylist() == exp.column(0).to_pylist()      # filter/project keep row order
    assert got.column(1).to_pylist() == exp.column(1).to_pylist()


def test_conjunctive_compare_fast_paths_match_vm_and_arrow():
    # column-vs-literal conjunctions take the interval / simple-term kernels; they must agree with the interpreter and Arrow
    import os
    rng = np.random.default_rng(77)
    n = 70_001
    t = pa.table({"i": pa.array(rng.integers(-50, 50, n), type=pa.int32(), mask=rng.random(n) < 0.05),
                  "l": pa.array(rng.integers(-2**62, 2**62, n), type=pa.int64(), mask=rng.random(n) < 0.05),
                  "f": pa.array(np.where(rng.random(n) < 0.02, np.nan, rng.standard_normal(n)), mask=rng.random(n) < 0.05),
                  "d": pa.array([None if x else decimal.Decimal(int(v)) / 100 for x, v in zip(rng.random(n) < 0.05, rng.integers(-5000, 5000, n))], type=pa.decimal128(7, 2)),
                  "b": pa.array(rng.random(n) < 0.5, mask=rng.random(n) < 0.05),
                  "row": pa.array(np.arange(n), type=pa.int64())})
    I, Lc, F, D, B = P.col("i"), P.col("l"), P.col("f"), P.col("d"), P.col("b")
    li = lambda v: P.lit(v, pa.int32())
    cases = [
        ([P.binary("GtEq", I, li(-10)), P.binary("Lt", I, li(20))], pc.and_kleene(pc.greater_equal(t["i"], -10), pc.less(t["i"], 20))),
        ([P.binary("Gt", li(5), I)], pc.less(t["i"], 5)),                                       # literal on the left
        ([P.binary("Eq", I, li(7)), P.is_not_null(Lc)], pc.and_kleene(pc.equal(t["i"], 7), pc.is_valid(t["l"]))),
        ([P.binary("NotEq", I, li(0)), P.binary("LtEq", Lc, P.lit(0, pa.int64()))], pc.and_kleene(pc.not_equal(t["i"], 0), pc.less_equal(t["l"], 0))),
        ([P.is_null(I), P.binary("Gt", Lc, P.lit(-2**61, pa.int64()))], pc.and_kleene(pc.is_null(t["i"]), pc.greater(t["l"], -2**61))),
        ([P.binary("Lt", Lc, P.lit(-2**63, pa.int64()))], pc.less(t["l"], -2**63)),             # empty interval
        ([P.binary("GtEq", Lc, P.lit(-2**61, pa.int64())), P.binary("Lt", P.col("row"), P.lit(60_000, pa.int64()))],     # two int64 columns (vector kernel)
         pc.and_kleene(pc.greater_equal(t["l"], -2**61), pc.less(t["row"], 60_000))),
        ([P.binary("Gt", I, P.lit(2**31 - 1, pa.int32()))], pc.greater(t["i"], 2**31 - 1)),     # empty after clamping to the int32 domain
        ([P.binary("Gt", F, P.lit(0.25, pa.float64())), P.binary("Eq", B, P.lit(True, pa.bool_()))], None),   # NaN > x under totalOrder
        ([P.binary("GtEq", D, P.lit(decimal.Decimal("-1.50"), pa.decimal128(7, 2))), P.binary("Lt", D, P.lit(decimal.Decimal("12.34"), pa.decimal128(7, 2)))],
         pc.and_kleene(pc.greater_equal(t["d"], decimal.Decimal("-1.50")), pc.less(t["d"], decimal.Decimal("12.34")))),
    ]
    for preds, arrow_mask in cases:
        plan = P.filter_(P.ffi_reader(t.schema, "t"), preds)
        fast = run(plan, {"t": t})["row"].to_pylist()
        os.environ["AURON_DISABLE_SIMPLE_PREDICATE"] = "1"
        try:
            vm = run(plan, {"t": t})["row"].to_pylist()
        finally:
            os.environ.pop("AURON_DISABLE_SIMPLE_PREDICATE", None)
        assert fast == vm
        if arrow_mask is not None:
            assert fast == t.filter(arrow_mask)["row"].to_pylist()
        else:
            f = t["f"].to_pylist()
            b = t["b"].to_pylist()
            exp = [i for i in range(n) if f[i] is not None and b[i] and (math.isnan(f[i]) or f[i] > 0.25)]
            assert fast == exp


def _eval(t: pa.Table, exprs, names, types):
    plan = P.projection(P.ffi_reader(t.schema, "t"), exprs, names, types)
    return run(plan, {"t": t})


def test_arithmetic_wrap_null_and_division():
    t = pa.table({"x": pa.array([2**31 - 1, -2**31, 7, None, 0], type=pa.int32()), "y": pa.array([1, -1, 0, 5, 3], type=pa.int32()),
                  "f": pa.array([1.5, -0.0, float("nan"), None, float("inf")])})
    got = _eval(t, [P.binary("Plus", P.col("x"), P.col("y")), P.binary("Multiply", P.col("x"), P.col("y")),
                    P.binary("Divide", P.col("x"), P.scalar_fn("Spark_NullIfZero", [P.col("y")], pa.int32())),
                    P.binary("Modulo", P.col("x"), P.col("y")), P.binary("Plus", P.col("f"), P.lit(1.0, pa.float64())),
                    P.negative(P.col("x"))],
                ["add", "mul", "div", "mod", "fadd", "neg"], [pa.int32()] * 4 + [pa.float64(), pa.int32()])
    assert got["add"].to_pylist() == [-2**31, 2**31 - 1, 7, None, 3]             # wraps (arrow *_wrapping)
    assert got["mul"].to_pylist() == [2**31 - 1, -2**31, 0, None, 0]
    assert got["div"].to_pylist() == [2**31 - 1, -2**31, None, None, 0]           # x/0 -> NULL, MIN/-1 wraps
    assert got["mod"].to_pylist() == [0, 0, None, None, 0]
    f = got["fadd"].to_pylist()
    assert f[0] == 2.5 and f[1] == 1.0 and math.isnan(f[2]) and f[3] is None and f[4] == float("inf")
    assert got["neg"].to_pylist() == [-(2**31 - 1), -2**31, -7, None, 0]


def test_comparisons_kleene_case_in():
    t = pa.table({"x": pa.array([1, 5, None, 9, 3], type=pa.int64()), "b": pa.array([True, None, False, True, None]),
                  "f": pa.array([float("nan"), 1.0, -0.0, None, 2.0])})
    x, b = P.col("x"), P.col("b")
    gt3 = P.binary("Gt", x, P.lit(3, pa.int64()))
    got = _eval(t, [gt3, P.binary("And", gt3, b), P.binary("Or", gt3, b), P.not_(b), P.is_null(x), P.is_not_null(x),
                    P.case([(gt3, P.lit(100, pa.int64())), (P.binary("Eq", x, P.lit(1, pa.int64())), P.lit(200, pa.int64()))], P.lit(-1, pa.int64())),
                    P.case([(gt3, x)]),
                    P.in_list(x, [P.lit(1, pa.int64()), P.lit(9, pa.int64())]), P.in_list(x, [P.lit(1, pa.int64()), P.lit(None, pa.int64())]),
                    P.in_list(x, [P.lit(5, pa.int64())], negated=True),
                    P.binary("Eq", P.col("f"), P.col("f")), P.binary("IsNotDistinctFrom", x, P.lit(None, pa.int64())),
                    P.scalar_fn("Coalesce", [x, P.lit(0, pa.int64())], pa.int64())],
                ["gt", "and", "or", "not", "isn", "isnn", "case", "case2", "in", "in_null", "notin", "feq", "nseq", "coal"],
                [pa.bool_()] * 6 + [pa.int64(), pa.int64()] + [pa.bool_()] * 5 + [pa.int64()])
    assert got["gt"].to_pylist() == [False, True, None, True, False]
    assert got["and"].to_pylist() == [False, None, False, True, False]            # Kleene
    assert got["or"].to_pylist() == [True, True, None, True, None]
    assert got["not"].to_pylist() == [False, None, True, False, None]
    assert got["isn"].to_pylist() == [False, False, True, False, False]
    assert got["isnn"].to_pylist() == [True, True, False, True, True]
    assert got["case"].to_pylist() == [200, 100, -1, 100, -1]
    assert got["case2"].to_pylist() == [None, 5, None, 9, None]
    assert got["in"].to_pylist() == [True, False, None, True, False]
    assert got["in_null"].to_pylist() == [True, None, None, None, None]
    assert got["notin"].to_pylist() == [True, False, None, True, True]
    assert got["feq"].to_pylist() == [True, True, True, None, True]               # totalOrder: NaN == NaN (arrow-ord cmp)
    assert got["nseq"].to_pylist() == [False, False, True, False, False]
    assert got["coal"].to_pylist() == [1, 5, 0, 9, 3]


def test_casts_golden():
    # datafusion-ext-commons/src/arrow/cast.rs:553-752
    f = pa.table({"f": pa.array([None, 123.456, 987.654, 2147483647 + 10000.0, -2147483648 - 10000.0, math.inf, -math.inf, math.nan])})
    got = _eval(f, [P.try_cast(P.col("f"), pa.int32())], ["i"], [pa.int32()])
    assert got["i"].to_pylist() == [None, 123, 987, 2147483647, -2147483648, 2147483647, -2147483648, 0]
    i = pa.table({"i": pa.array([None, 123, 987, 2**31 - 1, -2**31], type=pa.int32())})
    got = _eval(i, [P.try_cast(P.col("i"), pa.float64()), P.try_cast(P.col("i"), pa.decimal128(38, 18)), P.try_cast(P.col("i"), pa.int8()),
                    P.try_cast(P.col("i"), pa.int64())], ["f", "d", "i8", "i64"], [pa.float64(), pa.decimal128(38, 18), pa.int8(), pa.int64()])
    assert got["f"].to_pylist() == [None, 123.0, 987.0, float(2**31 - 1), float(-2**31)]
    assert [None if v is None else int(v.scaleb(18)) for v in got["d"].to_pylist()] == [
        None, 123 * 10**18, 987 * 10**18, (2**31 - 1) * 10**18, (-2**31) * 10**18]
    assert got["i8"].to_pylist() == [None, 123, None, None, None]                 # arrow safe cast: out of range -> NULL
    assert got["i64"].to_pylist() == [None, 123, 987, 2**31 - 1, -2**31]
    s = pa.table({"s": pa.array([None, "123", "987", "987.654", "123456789012345", "-123456789012345", "999999999999999999999999999999999",
                                 "12x", "", "+", "-5.", "2001-02-03"])})
    got = _eval(s, [P.try_cast(P.col("s"), pa.int64()), P.try_cast(P.col("s"), pa.int32())], ["l", "i"], [pa.int64(), pa.int32()])
    assert got["l"].to_pylist() == [None, 123, 987, 987, 123456789012345, -123456789012345, None, None, None, None, -5, None]
    assert got["i"].to_pylist() == [None, 123, 987, 987, None, None, None, None, None, None, -5, None]
    d = pa.table({"s": pa.array([None, "2001-02-03", "2001-03-04", "2001-04-05T06:07:08", "2001-04", "2002", "2001-00", "2001-13", "9999-99",
                                 "99999-01", " 2000-02-29 ", "2001-02-30"])})
    got = _eval(d, [P.try_cast(P.col("s"), pa.date32())], ["d"], [pa.date32()])
    assert got["d"].to_pylist() == [None, dt.date(2001, 2, 3), dt.date(2001, 3, 4), dt.date(2001, 4, 5), dt.date(2001, 4, 1), dt.date(2002, 1, 1),
                                    None, None, None, None, dt.date(2000, 2, 29), None]
    dec = pa.table({"d": pa.array([None, decimal.Decimal("123.456"), decimal.Decimal("-0.005"), decimal.Decimal("99999.995"), decimal.Decimal("-7.5")],
                                  type=pa.decimal128(10, 3))})
    got = _eval(dec, [P.try_cast(P.col("d"), pa.decimal128(12, 5)), P.try_cast(P.col("d"), pa.decimal128(7, 2)), P.try_cast(P.col("d"), pa.int32()),
                      P.try_cast(P.col("d"), pa.float64())], ["up", "down", "i", "f"],
                [pa.decimal128(12, 5), pa.decimal128(7, 2), pa.int32(), pa.float64()])
    D = decimal.Decimal
    assert got["up"].to_pylist() == [None, D("123.45600"), D("-0.00500"), D("99999.99500"), D("-7.50000")]
    assert got["down"].to_pylist() == [None, D("123.46"), D("-0.01"), None, D("-7.50")]   # half away from zero; 100000.00 overflows (7,2)
    assert got["i"].to_pylist() == [None, 123, 0, 99999, -7]
    assert got["f"].to_pylist() == [None, 123.456, -0.005, 99999.995, -7.5]


def test_cast_string_to_decimal_golden():
    # datafusion-ext-commons/src/arrow/cast.rs:629-658 (scientific notation is rewritten to plain digits first, :328-351)
    g = pa.table({"s": pa.array([None, "1e-8", "1.012345678911111111e10", "1.42e-6", "0.00000142", "123.456", "987.654",
                                 "123456789012345.678901234567890", "-123456789012345.678901234567890"])})
    got = _eval(g, [P.try_cast(P.col("s"), pa.decimal128(38, 18))], ["d"], [pa.decimal128(38, 18)])
    with decimal.localcontext() as ctx:
        ctx.prec = 60                                    # the default 28 digits would round the 33-digit values
        unscaled = [None if v is None else int(v.scaleb(18)) for v in got["d"].to_pylist()]
    assert unscaled == [
        None, 10000000000, 10123456789111111110000000000, 1420000000000, 1420000000000, 123456000000000000000, 987654000000000000000,
        123456789012345678901234567890000, -123456789012345678901234567890000]
    # arrow's parser behind it: fraction digits past the scale are dropped, too many digits / malformed input -> NULL, no trimming
    e = pa.table({"s": pa.array(["987.6549999", "-0.019", "+5", ".5", "1.", "12345.67", "123456.7", "", ".", "-", "1e", "1e+2", "12x", " 1", "0.000000", "1E2"])})
    got = _eval(e, [P.try_cast(P.col("s"), pa.decimal128(7, 2))], ["d"], [pa.decimal128(7, 2)])
    D = decimal.Decimal
    assert got["d"].to_pylist() == [D("987.65"), D("-0.01"), D("5.00"), D("0.50"), D("1.00"), D("12345.67"), None, None, None, None, None, D("100.00"), None, None,
                                    D("0.00"), D("100.00")]


def test_casts_to_string_golden_and_vs_arrow():
    # cast.rs:536-552 (bool -> "true"/"false"), cast.rs:660-690 (decimal -> plain string with `scale` fraction digits); integers and
    # dates go through arrow's cast in the reference: compared with Arrow C++ here
    b = pa.table({"b": pa.array([None, True, False])})
    assert _eval(b, [P.try_cast(P.col("b"), pa.string())], ["s"], [pa.string()])["s"].to_pylist() == [None, "true", "false"]
    d = pa.table({"d": pa.array([None, decimal.Decimal("123.000000"), decimal.Decimal("-987.654321"), decimal.Decimal("0.000005"), decimal.Decimal("-0.500000"),
                                 decimal.Decimal("999999999999.999999")], type=pa.decimal128(18, 6)),
                  "z": pa.array([None, decimal.Decimal(7), decimal.Decimal(-12345), decimal.Decimal(0), decimal.Decimal(10), decimal.Decimal(-1)], type=pa.decimal128(9, 0))})
    got = _eval(d, [P.try_cast(P.col("d"), pa.string()), P.try_cast(P.col("z"), pa.string())], ["s", "t"], [pa.string(), pa.string()])
    assert got["s"].to_pylist() == [None, "123.000000", "-987.654321", "0.000005", "-0.500000", "999999999999.999999"]
    assert got["t"].to_pylist() == [None, "7", "-12345", "0", "10", "-1"]
    rng = np.random.default_rng(4)
    n = 20_000
    t = pa.table({"i8": pa.array(rng.integers(-128, 128, n), type=pa.int8(), mask=rng.random(n) < 0.05),
                  "i32": pa.array(rng.integers(-2**31, 2**31, n), type=pa.int32(), mask=rng.random(n) < 0.05),
                  "i64": pa.array(np.concatenate([[-2**63, 2**63 - 1, 0, -1], rng.integers(-2**62, 2**62, n - 4)]), type=pa.int64()),
                  "dt": pa.array(rng.integers(-200_000, 200_000, n).astype(np.int32), type=pa.date32(), mask=rng.random(n) < 0.05)})
    got = _eval(t, [P.try_cast(P.col(c), pa.string()) for c in t.column_names] + [P.try_cast(P.binary("Plus", P.col("i32"), P.col("i32")), pa.string())],
                t.column_names + ["sum"], [pa.string()] * 5)
    for c in ("i8", "i32", "i64"):
        assert got[c].to_pylist() == t[c].cast(pa.string()).to_pylist(), c
    exp_dt = [None if v is None else (f"{v.year:04d}-{v.month:02d}-{v.day:02d}") for v in t["dt"].to_pylist()]
    assert got["dt"].to_pylist() == exp_dt
    exp_sum = [None if v is None else str((int(v) * 2 + 2**31) % 2**32 - 2**31) for v in t["i32"].to_pylist()]
    assert got["sum"].to_pylist() == exp_sum                                    # wrapping int32 addition, then formatted


def test_string_predicates_golden():
    # datafusion-ext-exprs/src/string_starts_with.rs:137-165, string_ends_with.rs:137-168, string_contains.rs:136-168
    s1 = pa.table({"s": pa.array([None, "rabaok", "rraara", "s_skdo[]ra.,?';,{}\ra", " raefuwidn"])})
    assert _eval(s1, [P.starts_with(P.col("s"), "ra")], ["r"], [pa.bool_()])["r"].to_pylist() == [None, True, False, False, False]
    s2 = pa.table({"s": pa.array(["abrrbrr", "rrjndebcsabdji", None, "rr", "roser r"])})
    assert _eval(s2, [P.ends_with(P.col("s"), "rr")], ["r"], [pa.bool_()])["r"].to_pylist() == [True, False, None, True, False]
    s3 = pa.table({"s": pa.array(["abrr", "barr", "rnba", "nbar", None])})
    assert _eval(s3, [P.contains(P.col("s"), "ba")], ["r"], [pa.bool_()])["r"].to_pylist() == [False, True, True, True, None]


def test_like_substr_and_friends_vs_arrow():
    rng = np.random.default_rng(5)
    alphabet = list("ab_%x天")
    vals = ["".join(alphabet[int(i)] for i in rng.integers(0, len(alphabet), int(l))) for l in rng.integers(0, 9, 3000)]
    t = pa.table({"s": pa.array(vals, mask=rng.random(3000) < 0.05)})
    for pat in ["a%", "%b", "%ab%", "a_b%", "_", "%", "", "a\\%%", "%天_", "ab", "%a%b%"]:
        got = _eval(t, [P.like(P.col("s"), P.lit(pat, pa.string())), P.like(P.col("s"), P.lit(pat, pa.string()), negated=True)],
                    ["l", "nl"], [pa.bool_(), pa.bool_()])
        exp = pc.match_like(t["s"], pat)
        assert got["l"].to_pylist() == exp.to_pylist(), pat
        assert got["nl"].to_pylist() == pc.invert(exp).to_pylist(), pat
    got = _eval(t, [P.scalar_fn("Substr", [P.col("s"), P.lit(2, pa.int64()), P.lit(3, pa.int64())], pa.string()),
                    P.scalar_fn("Substr", [P.col("s"), P.lit(3, pa.int64())], pa.string()),
                    P.scalar_fn("CharacterLength", [P.col("s")], pa.int32()), P.scalar_fn("Upper", [P.col("s")], pa.string()),
                    P.case([(P.starts_with(P.col("s"), "a"), P.lit("A!", pa.string()))], P.col("s"))],
                ["sub", "sub2", "len", "up", "cs"], [pa.string(), pa.string(), pa.int32(), pa.string(), pa.string()])
    assert got["sub"].to_pylist() == pc.utf8_slice_codeunits(t["s"], 1, 4).to_pylist()
    assert got["sub2"].to_pylist() == pc.utf8_slice_codeunits(t["s"], 2).to_pylist()
    assert got["len"].to_pylist() == pc.utf8_length(t["s"]).to_pylist()
    assert got["up"].to_pylist() == pc.ascii_upper(t["s"]).to_pylist()
    assert got["cs"].to_pylist() == [None if v is None else ("A!" if v.startswith("a") else v) for v in t["s"].to_pylist()]


def test_date_parts_vs_python():
    days = [-719162, -1, 0, 1, 59, 10957, 11016, 18321, 19000, 2932896, None]
    t = pa.table({"d": pa.array(days, type=pa.int32()).cast(pa.date32())})
    fns = ["Spark_Year", "Spark_Month", "Spark_Day", "Spark_DayOfWeek", "Spark_Quarter", "Spark_WeekOfYear"]
    got = _eval(t, [P.scalar_fn(f, [P.col("d")], pa.int32()) for f in fns], fns, [pa.int32()] * len(fns))
    epoch = dt.date(1970, 1, 1)
    for i, d in enumerate(days):
        if d is None:
            assert all(got[f][i].as_py() is None for f in fns)
            continue
        x = epoch + dt.timedelta(days=d)
        exp = [x.year, x.month, x.day, (x.isoweekday() % 7) + 1, (x.month - 1) // 3 + 1, x.isocalendar()[1]]
        assert [got[f][i].as_py() for f in fns] == exp, x


def _utc_ms(y, mo, d, h, mi, s):
    return int(dt.datetime(y, mo, d, h, mi, s, tzinfo=dt.timezone.utc).timestamp() * 1000)


def test_time_parts_and_session_time_zone_goldens():
    # datafusion-ext-functions/src/spark_dates.rs tests :661-952 (hour / minute / second) and :1077-1175 (date parts with a zone)
    S, I = pa.string(), pa.int32()
    hms = lambda h, m, s: (h * 3600 + m * 60 + s) * 1000
    ts = pa.table({"t": pa.array([0, hms(1, 23, 45), None, hms(12, 34, 56), -1000, _utc_ms(2019, 3, 10, 6, 59, 59), _utc_ms(2019, 3, 10, 7, 0, 0)],
                                 type=pa.timestamp("ms")),
                   "d": pa.array([0, 1, None, 100, -1, 17965, 17965], type=pa.int32()).cast(pa.date32())})
    f = lambda name, col, tz=None: P.scalar_fn(name, [P.col(col)] + ([P.lit(tz, S)] if tz is not None else []), I)
    exprs = {"h": f("Spark_Hour", "t"), "m": f("Spark_Minute", "t"), "s": f("Spark_Second", "t"),
             "hd": f("Spark_Hour", "d"), "md": f("Spark_Minute", "d"), "sd": f("Spark_Second", "d"),
             "h_utc": f("Spark_Hour", "t", "UTC"), "h_sh": f("Spark_Hour", "t", "Asia/Shanghai"),
             "m_kol": f("Spark_Minute", "t", "Asia/Kolkata"), "s_kol": f("Spark_Second", "t", "Asia/Kolkata"),
             "m_kat": f("Spark_Minute", "t", "Asia/Kathmandu"), "m_ny": f("Spark_Minute", "t", "America/New_York"),
             "s_ny": f("Spark_Second", "t", "America/New_York"), "h_ny": f("Spark_Hour", "t", "America/New_York"),
             "h_bad": f("Spark_Hour", "t", "Mars/Olympus")}
    got = _eval(ts, list(exprs.values()), list(exprs), [I] * len(exprs))
    N = None
    assert got["h"].to_pylist() == [0, 1, N, 12, 23, 6, 7]                 # :661-688, :714-744, :746-764 (-1000 ms -> 23:59:59)
    assert got["m"].to_pylist() == [0, 23, N, 34, 59, 59, 0]
    assert got["s"].to_pylist() == [0, 45, N, 56, 59, 59, 0]
    assert got["hd"].to_pylist() == [0, 0, N, 0, 0, 0, 0] == got["md"].to_pylist() == got["sd"].to_pylist()   # :691-711
    assert got["h_utc"].to_pylist() == got["h"].to_pylist()                # :805-833
    assert got["h_sh"].to_pylist()[0] == 8                                 # :784-802
    assert (got["m_kol"].to_pylist()[0], got["s_kol"].to_pylist()[0]) == (30, 0)    # :835-881
    assert got["m_kat"].to_pylist()[0] == 30                               # :884-904 (UTC+5:30 in 1970)
    assert got["m_ny"].to_pylist()[5:] == [59, 0] and got["s_ny"].to_pylist()[5:] == [59, 0]   # :907-937 spring forward
    assert got["h_ny"].to_pylist()[5:] == [1, 3]                           # 01:59:59 EST -> 03:00:00 EDT
    assert got["h_bad"].to_pylist() == got["h"].to_pylist()                # a zone chrono-tz cannot parse counts as none (:97-102)
    # date parts of timestamps in a zone
    t2 = pa.table({"t": pa.array([_utc_ms(2021, 1, 4, 4, 30, 0), _utc_ms(2021, 1, 4, 5, 30, 0), _utc_ms(2021, 4, 1, 3, 0, 0),
                                  _utc_ms(2021, 12, 31, 17, 0, 0), None], type=pa.timestamp("ms")),
                   "d": pa.array([0, 100, None, 4017, 16801], type=pa.int32()).cast(pa.date32())})
    names = ["Spark_Year", "Spark_Month", "Spark_Day", "Spark_DayOfWeek", "Spark_Quarter", "Spark_WeekOfYear"]
    ny = {n: P.scalar_fn(n, [P.col("t"), P.lit("America/New_York", S)], I) for n in names}
    sh = {n: P.scalar_fn(n, [P.col("t"), P.lit("Asia/Shanghai", S)], I) for n in names}
    g_ny = _eval(t2, list(ny.values()), names, [I] * 6)
    g_sh = _eval(t2, list(sh.values()), names, [I] * 6)
    row = lambda g, i: [g[n][i].as_py() for n in names]
    assert row(g_ny, 0) == [2021, 1, 3, 1, 1, 53]                          # :1077-1110, :569-585: 23:30 on Sunday Jan 3, ISO week 53
    assert row(g_ny, 1)[5] == 1                                            # 00:30 on Jan 4 local: ISO week 1
    assert row(g_ny, 2)[4] == 1                                            # :1113-1123: 23:00 on Mar 31 -> Q1
    assert row(g_sh, 3) == [2022, 1, 1, 7, 1, 52]                          # :1126-1153: 01:00 on Sat Jan 1 2022
    assert row(g_ny, 4) == [N] * 6
    # NULL zone literal == no zone (:1156-1175); dates keep their plain parts
    nz = {n: P.scalar_fn(n, [P.col("d"), P.lit(None, S)], I) for n in names}
    pl = {n: P.scalar_fn(n, [P.col("d")], I) for n in names}
    a, b = _eval(t2, list(nz.values()), names, [I] * 6), _eval(t2, list(pl.values()), names, [I] * 6)
    assert all(a[n].to_pylist() == b[n].to_pylist() for n in names)
    assert b["Spark_WeekOfYear"].to_pylist() == [1, 15, N, 1, 53]          # :546-566 (0, 4017, 16801 of the golden list)
    with pytest.raises(runtime.AuronError, match="invalid timezone"):      # :588-600
        _eval(t2, [P.scalar_fn("Spark_WeekOfYear", [P.col("t"), P.lit("Mars/Olympus", S)], I)], ["w"], [I])


@pytest.mark.parametrize("zone", ["America/New_York", "Europe/London", "Asia/Kolkata", "Australia/Lord_Howe", "America/Sao_Paulo"])
def test_time_zone_functions_vs_python_zoneinfo(zone):
    import zoneinfo
    rng = np.random.default_rng(41)
    n = 20_000
    secs = rng.integers(-2_000_000_000, 6_000_000_000, n)                  # 1906 .. 2160: past the tz database's explicit transitions
    us = secs * 1_000_000 + rng.integers(0, 1_000_000, n)
    t = pa.table({"t": pa.array(us, type=pa.timestamp("us"), mask=rng.random(n) < 0.02)})
    S, I = pa.string(), pa.int32()
    names = ["Spark_Hour", "Spark_Minute", "Spark_Second", "Spark_Year", "Spark_Month", "Spark_Day", "Spark_DayOfWeek", "Spark_Quarter"]
    got = _eval(t, [P.scalar_fn(f, [P.col("t"), P.lit(zone, S)], I) for f in names], names, [I] * len(names))
    tz = zoneinfo.ZoneInfo(zone)
    cols = [got[f].to_pylist() for f in names]
    valid = t["t"].is_valid().to_pylist()
    for i in range(n):
        if not valid[i]:
            assert all(c[i] is None for c in cols)
            continue
        # the reference first casts to Timestamp(Millisecond) (spark_dates.rs:348-351); arrow's unit cast divides toward zero,
        # so a pre-epoch instant with a sub-millisecond part lands on the millisecond above it
        u = int(us[i])
        ms = u // 1000 if u >= 0 else -((-u) // 1000)
        x = dt.datetime.fromtimestamp(ms // 1000, tz=dt.timezone.utc).astimezone(tz)
        exp = [x.hour, x.minute, x.second, x.year, x.month, x.day, (x.isoweekday() % 7) + 1, (x.month - 1) // 3 + 1]
        assert [c[i] for c in cols] == exp, (int(us[i]), x)


def _round_f32_array_branch(x, sc):
    x, f = np.float32(x), np.float32(1)
    for _ in range(abs(sc)):
        f = np.float32(f * np.float32(10))
    if sc < 0:
        f = np.float32(np.float32(1) / f)
    y = np.float32(x * f)
    r = np.floor(np.float32(y + np.float32(0.5))) if y >= 0 else np.ceil(np.float32(y - np.float32(0.5)))
    return float(np.float32(np.float32(r) / f))


def test_spark_round_goldens():
    # datafusion-ext-functions/src/spark_round.rs:225-447 (HALF_UP; decimals keep precision and scale; integers round at negative scales)
    I, D = pa.int32(), decimal.Decimal
    t = pa.table({"d": pa.array([D("123.45"), D("-678.95"), None], type=pa.decimal128(10, 2)),
                  "x": pa.array([123.45, -678.9, None]), "y": pa.array([1.2345, -2.3456, 0.5]), "z": pa.array([-0.5, -1.5, float("nan")]),
                  "h": pa.array([31415, 31415, None], type=pa.int16()), "i": pa.array([314159265, -314159265, None], type=pa.int32()),
                  "f": pa.array([3.1415, float("inf"), None], type=pa.float32()), "p": pa.array([math.pi, -math.pi, None]),
                  "l": pa.array([D(31415926535897932), D(-31415926535897932), None], type=pa.decimal128(38, 0))})
    rnd = lambda c, sc, ty: P.scalar_fn("Spark_Round", [P.col(c), P.lit(sc, I)], ty)
    got = _eval(t, [rnd("d", 1, pa.decimal128(10, 2)), rnd("x", -1, pa.float64()), rnd("y", 2, pa.float64()), rnd("z", 0, pa.float64())],
                ["d", "x", "y", "z"], [pa.decimal128(10, 2), pa.float64(), pa.float64(), pa.float64()])
    assert got["d"].to_pylist() == [D("123.50"), D("-679.00"), None]                     # :225-245
    assert got["x"].to_pylist() == [120.0, -680.0, None]                                 # :250-264
    assert got["y"].to_pylist() == [1.23, -2.35, 0.5]                                    # :268-292
    z = got["z"].to_pylist()
    assert z[:2] == [-1.0, -2.0] and math.isnan(z[2])                                    # -0.5 -> -1 (HALF_UP), :297-306; NaN passes through
    scales = list(range(-6, 7))
    for col, ty, exp, tol in [
            ("h", pa.int16(), [0, 0, 30000, 31000, 31400, 31420] + [31415] * 7, 0),                                                   # :309-327
            ("i", pa.int32(), [314000000, 314200000, 314160000, 314159000, 314159300, 314159270] + [314159265] * 7, 0),                # :383-408
            # :330-353 runs the scalar branch, which rounds a Float32 in f64 arithmetic (:146-151); a Float32 COLUMN takes the array
            # branch in f32 arithmetic (:100-113), where 3.1415f * 1000f is exactly 3141.5 and rounds up: restated with numpy float32
            ("f", pa.float32(), [_round_f32_array_branch(3.1415, sc) for sc in range(-6, 7)], 1e-7),
            ("p", pa.float64(), [0.0] * 6 + [3.0, 3.1, 3.14, 3.142, 3.1416, 3.14159, 3.141593], 1e-9),                                # :356-380
            # :410-447 up to scale 0; the golden runs the scalar branch, a COLUMN of decimals takes the array branch (:62-81), which for
            # a scale above the stored one multiplies the unscaled value instead (the declared scale is kept): mirrored as is
            ("l", pa.decimal128(38, 0), [31415926536000000, 31415926535900000, 31415926535900000, 31415926535898000, 31415926535897900,
                                          31415926535897930, 31415926535897932] + [31415926535897932 * 10**k for k in range(1, 7)], 0)]:
        g = _eval(t, [rnd(col, sc, ty) for sc in scales], [f"s{k}" for k in range(13)], [ty] * 13)
        first = [g[f"s{k}"][0].as_py() for k in range(13)]
        if tol:
            assert all(abs(a - b) < tol for a, b in zip(first, exp)), (col, first)
        else:
            assert [int(v) for v in first] == exp, (col, first)
        assert all(g[f"s{k}"][2].as_py() is None for k in range(13))
        if col in ("i", "l"):      # negative values mirror (HALF_UP rounds away from zero)
            assert [int(g[f"s{k}"][1].as_py()) for k in range(13)] == [-v for v in exp], col
    assert math.isinf(_eval(t, [rnd("f", 2, pa.float32())], ["f"], [pa.float32()])["f"][1].as_py())


def test_spark_bround_goldens():
    # datafusion-ext-functions/src/spark_bround.rs:256-513 (HALF_EVEN, "banker's rounding")
    I, D = pa.int32(), decimal.Decimal
    t = pa.table({"a": pa.array([1.5, 2.5, -0.5, -1.5, 0.5, 3.5, -2.5, -3.5]), "b": pa.array([125.0, 135.0, 145.0, 155.0, -35.0, None, 0.0, 1e300]),
                  "c": pa.array([-0.35] + [None] * 7), "d": pa.array([D("123.45"), D("678.95"), D("-123.45"), D("-678.95")] + [None] * 4, type=pa.decimal128(10, 2)),
                  "h": pa.array([31415] * 8, type=pa.int16()), "i": pa.array([314159265, -314159265] * 4, type=pa.int32()),
                  "p": pa.array([math.pi] * 8), "l": pa.array([D(31415926535897932), D(-31415926535897932)] * 4, type=pa.decimal128(38, 0))})
    br = lambda c, sc, ty: P.scalar_fn("Spark_BRound", [P.col(c), P.lit(sc, I)], ty)
    F = pa.float64()
    got = _eval(t, [br("a", 0, F), br("b", -1, F), br("c", 1, F), br("d", 1, pa.decimal128(10, 2))], ["a", "b", "c", "d"], [F, F, F, pa.decimal128(10, 2)])
    assert got["a"].to_pylist() == [2.0, 2.0, 0.0, -2.0, 0.0, 4.0, -2.0, -4.0]          # :256-279, :487-498 (ties go to the even neighbour)
    assert got["b"].to_pylist() == [120.0, 140.0, 140.0, 160.0, -40.0, None, 0.0, 1e300]   # :282-298
    assert got["c"].to_pylist()[0] == -0.4
    assert got["d"].to_pylist()[:4] == [D("123.40"), D("679.00"), D("-123.40"), D("-679.00")]   # :302-316 (12345 -> 12340: tie to even)
    scales = list(range(-6, 7))
    for col, ty, exp in [("h", pa.int16(), [0, 0, 30000, 31000, 31400, 31420] + [31415] * 7),                                          # :377-401
                         ("i", pa.int32(), [314000000, 314200000, 314160000, 314159000, 314159300, 314159260] + [314159265] * 7),       # :404-430
                         ("l", pa.decimal128(38, 0), [31415926536000000, 31415926535900000, 31415926535900000, 31415926535898000,
                                                      31415926535897900, 31415926535897930, 31415926535897932])]:                       # :433-464 up to scale 0
        g = _eval(t, [br(col, sc, ty) for sc in scales[:len(exp)]], [f"s{k}" for k in range(len(exp))], [ty] * len(exp))
        assert [int(g[f"s{k}"][0].as_py()) for k in range(len(exp))] == exp, col
        if col != "h":
            assert [int(g[f"s{k}"][1].as_py()) for k in range(len(exp))] == [-v for v in exp], col
    g = _eval(t, [br("p", sc, F) for sc in scales], [f"s{k}" for k in range(13)], [F] * 13)
    exp = [0.0] * 6 + [3.0, 3.1, 3.14, 3.142, 3.1416, 3.14159, 3.141593]                 # :325-348
    assert all(abs(g[f"s{k}"][0].as_py() - exp[k]) < 1e-9 for k in range(13))


def test_murmur3_expr_and_misc_functions():
    t = _random_table(2000, seed=11)
    got = _eval(t, [P.scalar_fn("Spark_Murmur3Hash", [P.col("i32"), P.col("s")], pa.int32()),
                    P.scalar_fn("Spark_XxHash64", [P.col("i64"), P.col("dec")], pa.int64()),
                    P.scalar_fn("Spark_IsNaN", [P.col("f64")], pa.bool_()), P.scalar_fn("Abs", [P.col("i32")], pa.int32()),
                    P.scalar_fn("Sqrt", [P.scalar_fn("Abs", [P.col("f64")], pa.float64())], pa.float64())],
                ["m3", "xx", "nan", "abs", "sqrt"], [pa.int32(), pa.int64(), pa.bool_(), pa.int32(), pa.float64()])
    assert (got["m3"].to_numpy() == oracle.hash_columns([t["i32"].combine_chunks(), t["s"].combine_chunks()])).all()
    assert (got["xx"].to_numpy() == oracle.hash_columns([t["i64"].combine_chunks(), t["dec"].combine_chunks()], "xxhash64")).all()
    assert got["nan"].to_pylist() == [False if v is None else math.isnan(v) for v in t["f64"].to_pylist()]
    exp_abs = [None if v is None else (v if v >= 0 else (-v if v != -2**31 else -2**31)) for v in t["i32"].to_pylist()]
    assert got["abs"].to_pylist() == exp_abs
    np.testing.assert_allclose(np.array(got["sqrt"].fill_null(0).to_pylist()), np.sqrt(np.abs(np.array(t["f64"].fill_null(0).to_pylist()))), rtol=1e-12)


# =============================================================================== aggregate (A1-A5)
def test_agg_golden_partial_then_final():
    # datafusion-ext-plans/src/agg_exec.rs:495-682 (collect_* / UDAF columns are out of scope)
    t = table_i32(a=[2, 9, 3, 1, 0, 4, 6], b=[1, 0, 0, 3, 5, 6, 3], c=[7, 8, 7, 8, 9, 2, 5], d=[-7, 86, 71, 83, 90, -2, 5],
                  e=[-7, 86, 71, 83, 90, -2, 5], f=[0, 1, 2, 3, 4, 5, 6], g=[6, 3, 6, 3, 1, 5, 4], h=[6, 3, 6, 3, 1, 5, 4])
    aggs = [("SUM", "a", pa.int64()), ("AVG", "b", pa.float64()), ("MAX", "d", pa.int32()), ("MIN", "e", pa.int32()), ("COUNT", "f", pa.int64()),
            ("FIRST_IGNORES_NULL", "h", pa.int32())]
    names = ["agg_expr_sum", "agg_expr_avg", "agg_expr_max", "agg_expr_min", "agg_expr_count", "agg_agg_firstign"]
    partial = P.agg(P.ffi_reader(t.schema, "t"), [P.col("c")], ["c"], [P.agg_expr(f, [P.col(c)], rt) for f, c, rt in aggs], names, ["PARTIAL"] * 6)
    final = P.agg(partial, [P.col("c")], ["c"], [P.agg_expr(f, [P.lit(None, pa.null())], rt) for f, c, rt in aggs], names, ["FINAL"] * 6)
    got = run(final, {"t": t})
    assert got.schema.names == ["c"] + names
    exp = [(2, 4, 6.0, -2, -2, 1, 5), (5, 6, 3.0, 5, 5, 1, 4), (7, 5, 0.5, 71, -7, 2, 6), (8, 10, 1.5, 86, 83, 2, 3), (9, 0, 5.0, 90, 90, 1, 1)]
    assert canon(got) == exp
    # the partial stage alone exposes the accumulator layout [c, sum, avg.sum, avg.count, max, min, count, first]
    part = run(partial, {"t": t})
    assert part.num_columns == 8 and [f.type for f in part.schema][1:] == [pa.int64(), pa.float64(), pa.int64(), pa.int32(), pa.int32(), pa.int64(), pa.int32()]


@pytest.mark.parametrize("n,card,chunk,scale", [(1000, 7, None, 1), (300_000, 2000, 50_000, 1), (300_000, 250_000, 100_000, 1),
                                                (300_000, 2000, 50_000, 10**12)])
def test_agg_fuzz_sum_count_vs_oracle(n, card, chunk, scale):
    # fuzz test of agg_exec.rs:716-843: SUM/COUNT vs a hash map, nullable keys and values, multi-chunk merge.
    # scale = 1: dense integer keys (direct-address path for the large cases); scale = 10^12: sparse keys (hash table path)
    rng = np.random.default_rng(n + card)
    t = pa.table({"k": pa.array(rng.integers(-card // 2, card // 2, n) * scale, type=pa.int64(), mask=rng.random(n) < 0.01),
                  "v": pa.array(rng.integers(-10**6, 10**6, n), type=pa.int64(), mask=rng.random(n) < 0.03)})
    plan = P.agg(P.ffi_reader(t.schema, "t"), [P.col("k")], ["k"],
                 [P.agg_expr("SUM", [P.col("v")], pa.int64()), P.agg_expr("COUNT", [P.col("v")], pa.int64())], ["s", "c"], ["PARTIAL", "PARTIAL"])
    td = P.task_definition(plan)
    import os
    if chunk:   # force several device chunks so partial results are merged (agg_table.rs partial -> merge)
        os.environ["AURON_GPU_CHUNK_ROWS"] = str(chunk)
    try:
        with runtime.Task(td, {"t": batches(t, chunk)}) as task:
            got = pa.Table.from_batches(list(task), schema=task.schema)
    finally:
        os.environ.pop("AURON_GPU_CHUNK_ROWS", None)
    exp = oracle.agg_sum_count_i64(t["k"].combine_chunks(), t["v"].combine_chunks())
    assert_same_rows(got, exp)


@pytest.mark.parametrize("card", [7, 400, 5000])
def test_agg_low_cardinality_shared_memory_variant(card):
    # few groups => the per-CTA shared-memory pre-aggregation variant; 5000 groups overflow its 2048 slots for some CTAs only
    # when the sample says "low" -- every path must give the exact result
    rng = np.random.default_rng(card)
    n = 2_500_000
    k = np.where(rng.random(n) < 0.5, rng.integers(0, min(card, 300), n), rng.integers(0, card, n))   # skewed: hot keys
    if card == 5000:
        k[: 1 << 18] = rng.integers(0, 100, 1 << 18)    # the sampled prefix looks low-cardinality, the tail is not
    t = pa.table({"k": pa.array(k, type=pa.int32(), mask=rng.random(n) < 0.01), "v": pa.array(rng.integers(-10**6, 10**6, n), type=pa.int64(), mask=rng.random(n) < 0.03),
                  "f": pa.array(rng.standard_normal(n), mask=rng.random(n) < 0.03)})
    plan = P.agg(P.ffi_reader(t.schema, "t"), [P.col("k")], ["k"],
                 [P.agg_expr("SUM", [P.col("v")], pa.int64()), P.agg_expr("COUNT", [P.col("v")], pa.int64()), P.agg_expr("MIN", [P.col("v")], pa.int64()),
                  P.agg_expr("MAX", [P.col("f")], pa.float64())], ["s", "c", "mn", "mx"], ["PARTIAL"] * 4)
    import os
    g = t.group_by("k").aggregate([("v", "sum"), ("v", "count"), ("v", "min"), ("f", "max")])
    exp = pa.table({"k": g["k"], "s": g["v_sum"], "c": g["v_count"], "mn": g["v_min"], "mx": g["f_max"]})
    assert_same_rows(run(plan, {"t": t}), exp)                      # default: L2-atomic kernel
    os.environ["AURON_ENABLE_SMEM_AGG"] = "1"
    try:
        assert_same_rows(run(plan, {"t": t}), exp)                  # opt-in shared-memory variant
    finally:
        os.environ.pop("AURON_ENABLE_SMEM_AGG", None)


def test_agg_filter_fusion_and_expressions():
    rng = np.random.default_rng(9)
    n = 100_000
    t = pa.table({"k": pa.array(rng.integers(0, 300, n), type=pa.int32()), "v": pa.array(rng.integers(-1000, 1000, n), type=pa.int32()),
                  "f": pa.array(rng.integers(0, 100, n), type=pa.int32(), mask=rng.random(n) < 0.02)})
    flt = P.filter_(P.ffi_reader(t.schema, "t"), [P.binary("GtEq", P.col("f"), P.lit(10, pa.int32())), P.binary("Lt", P.col("f"), P.lit(60, pa.int32()))])
    # GROUP BY cast(k as bigint), SUM(v * 2), COUNT(v), MIN(v), MAX(v), AVG(v)
    plan = P.agg(flt, [P.try_cast(P.col("k"), pa.int64())], ["k"],
                 [P.agg_expr("SUM", [P.binary("Multiply", P.col("v"), P.lit(2, pa.int32()))], pa.int64()), P.agg_expr("COUNT", [P.col("v")], pa.int64()),
                  P.agg_expr("MIN", [P.col("v")], pa.int32()), P.agg_expr("MAX", [P.col("v")], pa.int32()), P.agg_expr("AVG", [P.col("v")], pa.float64())],
                 ["s", "c", "mn", "mx", "avg"], ["PARTIAL"] * 5)
    final = P.agg(plan, [P.col("k")], ["k"],
                  [P.agg_expr(f, [P.lit(None, pa.null())], rt) for f, rt in [("SUM", pa.int64()), ("COUNT", pa.int64()), ("MIN", pa.int32()), ("MAX", pa.int32()), ("AVG", pa.float64())]],
                  ["s", "c", "mn", "mx", "avg"], ["FINAL"] * 5)
    got = run(final, {"t": t}, chunk=30_000)
    ft = t.filter(pc.and_kleene(pc.greater_equal(t["f"], 10), pc.less(t["f"], 60)))
    g = pa.table({"k": ft["k"].cast(pa.int64()), "v2": pc.multiply(ft["v"].cast(pa.int64()), 2), "v": ft["v"]}).group_by("k").aggregate(
        [("v2", "sum"), ("v", "count"), ("v", "min"), ("v", "max"), ("v", "mean")])
    exp = pa.table({"k": g["k"], "s": g["v2_sum"], "c": g["v_count"], "mn": g["v_min"], "mx": g["v_max"], "avg": g["v_mean"]})
    assert_same_rows(got, exp, float_tol=1e-9)


@pytest.mark.parametrize("direct", ["default", "force", "off"])
def test_agg_consumes_filter_mask_on_every_kernel_path(direct, monkeypatch):
    # single integer keys normally take the direct-address path (slot = key - min); "off" keeps them on the hash table
    if direct == "force":
        monkeypatch.setenv("AURON_FORCE_DIRECT_AGG", "1")
    elif direct == "off":
        monkeypatch.setenv("AURON_DISABLE_DIRECT_AGG", "1")
        monkeypatch.setenv("AURON_DISABLE_AGG_VALID_FROM_COUNT", "1")
    # Filter -> HashAggregate with plain column arguments: the aggregate reads the filter's bit mask directly (no index vector).
    # Covers the 64-bit fast-key kernel, the general (string + int) row-key kernel, the no-grouping kernel and FIRST positions.
    rng = np.random.default_rng(33)
    n = 123_457
    words = ["x", "yy", "zzz", None]
    t = pa.table({"k": pa.array(rng.integers(0, 50, n), type=pa.int64(), mask=rng.random(n) < 0.03),
                  "s": pa.array([words[int(i)] for i in rng.integers(0, 4, n)]),
                  "v": pa.array(rng.integers(-100, 100, n), type=pa.int64(), mask=rng.random(n) < 0.1),
                  "f": pa.array(rng.integers(0, 10, n), type=pa.int32()),
                  "pos": pa.array(np.arange(n), type=pa.int64())})
    pred = [P.binary("GtEq", P.col("f"), P.lit(3, pa.int32()))]
    ft = t.filter(pc.greater_equal(t["f"], 3))
    aggs = [P.agg_expr("SUM", [P.col("v")], pa.int64()), P.agg_expr("COUNT", [P.col("v")], pa.int64()), P.agg_expr("MIN", [P.col("pos")], pa.int64())]
    finals = [P.agg_expr(f, [P.lit(None, pa.null())], pa.int64()) for f in ("SUM", "COUNT", "MIN")]
    for keys in (["k"], ["s", "k"], []):
        flt = P.filter_(P.ffi_reader(t.schema, "t"), pred)
        part = P.agg(flt, [P.col(k) for k in keys], keys, aggs, ["s_v", "c_v", "first_pos"], ["PARTIAL"] * 3)
        final = P.agg(part, [P.col(k) for k in keys], keys, finals, ["s_v", "c_v", "first_pos"], ["FINAL"] * 3)
        got = run(final, {"t": t}, chunk=40_000)
        if keys:
            g = ft.group_by(keys, use_threads=False).aggregate([("v", "sum"), ("v", "count"), ("pos", "min")])
            exp = pa.table({**{k: g[k] for k in keys}, "s_v": g["v_sum"], "c_v": g["v_count"], "first_pos": g["pos_min"]})
        else:
            exp = pa.table({"s_v": pa.array([pc.sum(ft["v"]).as_py()], type=pa.int64()), "c_v": pa.array([pc.count(ft["v"]).as_py()], type=pa.int64()),
                            "first_pos": pa.array([pc.min(ft["pos"]).as_py()], type=pa.int64())})
        assert_same_rows(got, exp)
    # FIRST keeps the first selected row of each group in input order
    flt = P.filter_(P.ffi_reader(t.schema, "t"), pred)
    got = run(P.agg(flt, [P.col("k")], ["k"], [P.agg_expr("FIRST", [P.col("pos")], pa.int64())], ["fp"], ["PARTIAL"]), {"t": t})
    first = {}
    for k, pos in zip(ft["k"].to_pylist(), ft["pos"].to_pylist()):
        first.setdefault(k, pos)
    got_map = dict(zip(got.column(0).to_pylist(), got.column(1).to_pylist()))
    assert got_map == first


def test_agg_string_and_multi_keys_decimal_sum():
    rng = np.random.default_rng(21)
    n = 50_000
    words = ["", "alpha", "beta", "gamma", "delta-epsilon-zeta", "天地", None]
    t = pa.table({"s": pa.array([words[int(i)] for i in rng.integers(0, len(words), n)]),
                  "k2": pa.array(rng.integers(0, 5, n), type=pa.int16(), mask=rng.random(n) < 0.1),
                  "d": pa.array([None if x else decimal.Decimal(int(v)) / 100 for x, v in zip(rng.random(n) < 0.02, rng.integers(-99999, 99999, n))],
                                type=pa.decimal128(7, 2)),
                  "x": pa.array(rng.standard_normal(n))})
    plan = P.agg(P.ffi_reader(t.schema, "t"), [P.col("s"), P.col("k2")], ["s", "k2"],
                 [P.agg_expr("SUM", [P.col("d")], pa.decimal128(17, 2)), P.agg_expr("AVG", [P.col("d")], pa.decimal128(11, 6)),
                  P.agg_expr("SUM", [P.col("x")], pa.float64()), P.agg_expr("COUNT", [], pa.int64())], ["sd", "ad", "sx", "c"], ["PARTIAL"] * 4)
    final = P.agg(plan, [P.col("s"), P.col("k2")], ["s", "k2"],
                  [P.agg_expr("SUM", [P.lit(None, pa.null())], pa.decimal128(17, 2)), P.agg_expr("AVG", [P.lit(None, pa.null())], pa.decimal128(11, 6)),
                   P.agg_expr("SUM", [P.lit(None, pa.null())], pa.float64()), P.agg_expr("COUNT", [P.lit(None, pa.null())], pa.int64())],
                  ["sd", "ad", "sx", "c"], ["FINAL"] * 4)
    got = run(final, {"t": t}, chunk=20_000)
    # oracle in exact Python arithmetic
    acc = {}
    for s, k2, d, x in zip(t["s"].to_pylist(), t["k2"].to_pylist(), t["d"].to_pylist(), t["x"].to_pylist()):
        a = acc.setdefault((s, k2), [None, 0, 0.0, 0])
        if d is not None:
            a[0] = (a[0] or 0) + int(d.scaleb(2))
            a[1] += 1
        a[2] += x
        a[3] += 1
    rows = []
    for (s, k2), (sd, cd, sx, c) in acc.items():
        avg = None
        if cd:
            scaled = sd * 10**4                      # AVG sum is cast to decimal(11,6): rescale by 10^4 (agg.rs:195)
            avg = decimal.Decimal(scaled // cd).scaleb(-6)   # div_euclid at the same scale (avg.rs:165-170)
        rows.append((s, k2, None if sd is None else decimal.Decimal(sd).scaleb(-2), avg, sx, c))
    exp = pa.table({"s": pa.array([r[0] for r in rows]), "k2": pa.array([r[1] for r in rows], type=pa.int16()),
                    "sd": pa.array([r[2] for r in rows], type=pa.decimal128(17, 2)), "ad": pa.array([r[3] for r in rows], type=pa.decimal128(11, 6)),
                    "sx": pa.array([r[4] for r in rows]), "c": pa.array([r[5] for r in rows], type=pa.int64())})
    assert_same_rows(got, exp, float_tol=1e-6)   # float SUM: 1e-6 relative (BASELINE north_star)


def test_agg_no_grouping_and_empty_input():
    t = pa.table({"v": pa.array([1, None, 3, 4], type=pa.int64())})
    plan = lambda: P.agg(P.ffi_reader(t.schema, "t"), [], [], [P.agg_expr("SUM", [P.col("v")], pa.int64()), P.agg_expr("COUNT", [P.col("v")], pa.int64()),
                                                               P.agg_expr("MAX", [P.col("v")], pa.int64())], ["s", "c", "m"], ["PARTIAL"] * 3)
    assert canon(run(plan(), {"t": t})) == [(8, 3, 4)]
    assert canon(run(plan(), {"t": t.slice(0, 0)})) == [(None, 0, None)]           # one row even without input (agg_exec.rs:280-323)
    grouped = P.agg(P.ffi_reader(t.schema, "t"), [P.col("v")], ["v"], [P.agg_expr("COUNT", [P.col("v")], pa.int64())], ["c"], ["PARTIAL"])
    assert run(grouped, {"t": t.slice(0, 0)}).num_rows == 0


# =============================================================================== joins (J1-J4, M1)
L1 = dict(a1=[1, 2, 3], b1=[4, 5, 5], c1=[7, 8, 9])
R1 = dict(a2=[10, 20, 30], b1=[4, 5, 6], c2=[70, 80, 90])


def _join(left, right, on, jt, impl, chunk=None):
    """left/right: a table, or a list of tables that arrive as separate input batches (build_table_from_batches, test.rs:68-78)"""
    lparts, rparts = (left if isinstance(left, list) else [left]), (right if isinstance(right, list) else [right])
    left, right = lparts[0], rparts[0]
    lsrc, rsrc = P.ffi_reader(left.schema, "l"), P.ffi_reader(right.schema, "r")
    names = [f"l_{n}" for n in left.column_names]
    if jt in ("SEMI", "ANTI"):
        fields = [pa.field(f"l_{f.name}", f.type) for f in left.schema]
    elif jt == "EXISTENCE":
        fields = [pa.field(f"l_{f.name}", f.type) for f in left.schema] + [pa.field("exists", pa.bool_())]
    else:
        fields = [pa.field(f"l_{f.name}", f.type) for f in left.schema] + [pa.field(f"r_{f.name}", f.type) for f in right.schema]
    s = pa.schema(fields)
    onp = [(P.col(l), P.col(r)) for l, r in on]
    if impl == "smj":   # Spark plans a SortExec on the join keys (ascending, nulls first) under each side of a SortMergeJoin
        lsrc = P.sort(lsrc, [P.sort_expr(P.col(l), True, True) for l, _ in on])
        rsrc = P.sort(rsrc, [P.sort_expr(P.col(r), True, True) for _, r in on])
        plan = P.sort_merge_join(s, lsrc, rsrc, onp, jt)
    elif impl.startswith("bhj"):
        plan = P.broadcast_join(s, lsrc, rsrc, onp, jt, "LEFT" if impl.endswith("L") else "RIGHT")
    else:
        plan = P.hash_join(s, lsrc, rsrc, onp, jt, "LEFT" if impl.endswith("L") else "RIGHT")
    td = P.task_definition(plan, stage_id=1, partition_id=0, task_id=7)
    return runtime.run_task(td, {"l": [b for t in lparts for b in batches(t, chunk)], "r": [b for t in rparts for b in batches(t, chunk)]})


IMPLS = ["smj", "bhjL", "bhjR", "shjL", "shjR"]   # joins/test.rs:366-381 runs every scenario on these five


@pytest.mark.parametrize("impl", IMPLS)
def test_join_golden_scenarios(impl):
    # datafusion-ext-plans/src/joins/test.rs:384-856
    N = None
    got = _join(table_i32(**L1), table_i32(**R1), [("b1", "b1")], "INNER", impl)
    assert canon(got) == [(1, 4, 7, 10, 4, 70), (2, 5, 8, 20, 5, 80), (3, 5, 9, 20, 5, 80)]
    got = _join(table_i32(a1=[1, 2, 2], b2=[1, 2, 2], c1=[7, 8, 9]), table_i32(a1=[1, 2, 3], b2=[1, 2, 2], c2=[70, 80, 90]),
                [("a1", "a1"), ("b2", "b2")], "INNER", impl)
    assert canon(got) == [(1, 1, 7, 1, 1, 70), (2, 2, 8, 2, 2, 80), (2, 2, 9, 2, 2, 80)]
    # join_inner_with_nulls :500-538 (NULL keys never match)
    got = _join(table_i32(a1=[1, 1, 2, 2], b2=[N, 1, 2, 2], c1=[1, N, 8, 9]), table_i32(a1=[1, 1, 2, 3], b2=[N, 1, 2, 2], c2=[10, 70, 80, 90]),
                [("a1", "a1"), ("b2", "b2")], "INNER", impl)
    assert canon(got) == [(1, 1, N, 1, 1, 70), (2, 2, 8, 2, 2, 80), (2, 2, 9, 2, 2, 80)]
    l7 = table_i32(a1=[1, 2, 3], b1=[4, 5, 7], c1=[7, 8, 9])
    got = _join(l7, table_i32(**R1), [("b1", "b1")], "LEFT", impl)      # join_left_one :652
    assert canon(got) == sorted([(1, 4, 7, 10, 4, 70), (2, 5, 8, 20, 5, 80), (3, 7, 9, N, N, N)], key=lambda r: tuple((v is not None, v or 0) for v in r))
    got = _join(l7, table_i32(**R1), [("b1", "b1")], "RIGHT", impl)     # join_right_one :686
    assert canon(got) == [(N, N, N, 30, 6, 90), (1, 4, 7, 10, 4, 70), (2, 5, 8, 20, 5, 80)]
    got = _join(table_i32(a1=[1, 2, 2, 3], b1=[4, 5, 5, 7], c1=[7, 8, 80, 9]), table_i32(a2=[10, 20, 20, 30], b2=[4, 5, 5, 6], c2=[70, 80, 800, 90]),
                [("b1", "b2")], "FULL", impl)                             # join_full_one :720
    assert canon(got) == [(N, N, N, 30, 6, 90), (1, 4, 7, 10, 4, 70), (2, 5, 8, 20, 5, 80), (2, 5, 8, 20, 5, 800), (2, 5, 80, 20, 5, 80),
                          (2, 5, 80, 20, 5, 800), (3, 7, 9, N, N, N)]
    got = _join(table_i32(a1=[1, 2, 2, 3, 5], b1=[4, 5, 5, 7, 7], c1=[7, 8, 8, 9, 11]), table_i32(**R1), [("b1", "b1")], "ANTI", impl)   # :757
    assert canon(got) == [(3, 7, 9), (5, 7, 11)]
    got = _join(table_i32(a1=[1, 2, 2, 3], b1=[4, 5, 5, 7], c1=[7, 8, 8, 9]), table_i32(**R1), [("b1", "b1")], "SEMI", impl)            # :790
    assert canon(got) == [(1, 4, 7), (2, 5, 8), (2, 5, 8)]
    got = _join(table_i32(a1=[1, 2, N, 4, 5], b1=[4, 5, 6, N, 8], c1=[7, 8, 9, 10, 11]), table_i32(a2=[10, 20, 30], b1=[4, 5, 7], c2=[70, 80, 90]),
                [("b1", "b1")], "ANTI", impl)                             # join_anti_with_null_keys :824
    assert canon(got) == [(N, 6, 9), (4, N, 10), (5, 8, 11)]
    got = _join(table_i32(a1=[1, 2, 3], b1=[4, 5, 7], c1=[7, 8, 9]), table_i32(**R1), [("b1", "b1")], "EXISTENCE", impl)
    assert canon(got) == [(1, 4, 7, True), (2, 5, 8, True), (3, 7, 9, False)]


@pytest.mark.parametrize("impl", IMPLS)
def test_join_golden_scenarios_part2(impl):
    # the rest of datafusion-ext-plans/src/joins/test.rs: the scenarios test_join_golden_scenarios does not restate
    N = None
    # join_inner_two_two :455-498 (duplicates on both sides of a two-column key)
    got = _join(table_i32(a1=[1, 1, 2], b2=[1, 1, 2], c1=[7, 8, 9]), table_i32(a1=[1, 1, 3], b2=[1, 1, 2], c2=[70, 80, 90]),
                [("a1", "a1"), ("b2", "b2")], "INNER", impl)
    assert canon(got) == [(1, 1, 7, 1, 1, 70), (1, 1, 7, 1, 1, 80), (1, 1, 8, 1, 1, 70), (1, 1, 8, 1, 1, 80)]
    # join_inner_batchsize :540-650 (5 x 7 rows of one key; the result must not depend on how the input is batched)
    lb = table_i32(a1=[1] * 5, b1=[1, 2, 3, 4, 5], c1=[1, 2, 3, 4, 5])
    rb = table_i32(a2=[1] * 7, b2=[1, 2, 3, 4, 5, 6, 7], c2=[1, 2, 3, 4, 5, 6, 7])
    exp = sorted((1, i, i, 1, j, j) for i in range(1, 6) for j in range(1, 8))
    for chunk in (2, 3, 4, 5, 6, 7):
        assert canon(_join(lb, rb, [("a1", "a2")], "INNER", impl, chunk=chunk)) == exp
    # join_with_duplicated_column_names :858-888 (left.a = right.b)
    got = _join(table_i32(a=[1, 2, 3], b=[4, 5, 7], c=[7, 8, 9]), table_i32(a=[10, 20, 30], b=[1, 2, 7], c=[70, 80, 90]), [("a", "b")], "INNER", impl)
    assert canon(got) == [(1, 4, 7, 10, 1, 70), (2, 5, 8, 20, 2, 80)]
    # join_date32 :890-923 / join_date64 :925-959 (every column is a date)
    d32 = lambda **c: pa.table({k: pa.array(v, type=pa.int32()).cast(pa.date32()) for k, v in c.items()})
    got = _join(d32(a1=[1, 2, 3], b1=[19107, 19108, 19108], c1=[7, 8, 9]), d32(a2=[10, 20, 30], b1=[19107, 19108, 19109], c2=[70, 80, 90]),
                [("b1", "b1")], "INNER", impl)
    assert got.schema.types == [pa.date32()] * 6
    day = lambda n: dt.date(1970, 1, 1) + dt.timedelta(days=n)
    assert canon(got) == [tuple(day(v) for v in r) for r in [(1, 19107, 7, 10, 19107, 70), (2, 19108, 8, 20, 19108, 80), (3, 19108, 9, 20, 19108, 80)]]
    assert day(19107) == dt.date(2022, 4, 25)
    d64 = lambda **c: pa.table({k: pa.array(v, type=pa.int64()).cast(pa.date64()) for k, v in c.items()})
    got = _join(d64(a1=[1, 2, 3], b1=[1650703441000, 1650903441000, 1650903441000], c1=[7, 8, 9]),
                d64(a2=[10, 20, 30], b1=[1650703441000, 1650503441000, 1650903441000], c2=[70, 80, 90]), [("b1", "b1")], "INNER", impl)
    assert got.schema.types == [pa.date64()] * 6
    ms = [tuple(r) for r in zip(*[c.cast(pa.int64()).to_pylist() for c in got.columns])]
    assert sorted(ms) == [(1, 1650703441000, 7, 10, 1650703441000, 70), (2, 1650903441000, 8, 30, 1650903441000, 90), (3, 1650903441000, 9, 30, 1650903441000, 90)]
    # join_left_sort_order :961-997 / join_right_sort_order :999-1031
    lo = table_i32(a1=[0, 1, 2, 3, 4, 5], b1=[3, 4, 5, 6, 6, 7], c1=[4, 5, 6, 7, 8, 9])
    ro = table_i32(a2=[0, 10, 20, 30, 40], b2=[2, 4, 6, 6, 8], c2=[50, 60, 70, 80, 90])
    left_exp = [(0, 3, 4, N, N, N), (1, 4, 5, 10, 4, 60), (2, 5, 6, N, N, N), (3, 6, 7, 20, 6, 70), (3, 6, 7, 30, 6, 80), (4, 6, 8, 20, 6, 70),
                (4, 6, 8, 30, 6, 80), (5, 7, 9, N, N, N)]
    got = _join(lo, ro, [("b1", "b2")], "LEFT", impl)
    assert canon(got) == canon(pa.table([pa.array([r[i] for r in left_exp], type=pa.int32()) for i in range(6)], names=list("abcdef")))
    got = _join(table_i32(a1=[0, 1, 2, 3], b1=[3, 4, 5, 7], c1=[6, 7, 8, 9]), table_i32(a2=[0, 10, 20, 30], b2=[2, 4, 5, 6], c2=[60, 70, 80, 90]),
                [("b1", "b2")], "RIGHT", impl)
    right_exp = [(N, N, N, 0, 2, 60), (1, 4, 7, 10, 4, 70), (2, 5, 8, 20, 5, 80), (N, N, N, 30, 6, 90)]
    assert canon(got) == canon(pa.table([pa.array([r[i] for r in right_exp], type=pa.int32()) for i in range(6)], names=list("abcdef")))
    # join_{left,right,full,existence}_multiple_batches :1033-1238 (both inputs arrive as two batches)
    l2 = [table_i32(a1=[0, 1, 2], b1=[3, 4, 5], c1=[4, 5, 6]), table_i32(a1=[3, 4, 5, 6], b1=[6, 6, 7, 9], c1=[7, 8, 9, 9])]
    r2 = [table_i32(a2=[0, 10, 20], b2=[2, 4, 6], c2=[50, 60, 70]), table_i32(a2=[30, 40], b2=[6, 8], c2=[80, 90])]
    matched = [(1, 4, 5, 10, 4, 60), (3, 6, 7, 20, 6, 70), (3, 6, 7, 30, 6, 80), (4, 6, 8, 20, 6, 70), (4, 6, 8, 30, 6, 80)]
    l_only = [(0, 3, 4, N, N, N), (2, 5, 6, N, N, N), (5, 7, 9, N, N, N), (6, 9, 9, N, N, N)]
    r_only = [(N, N, N, 0, 2, 50), (N, N, N, 40, 8, 90)]
    tab = lambda rows: pa.table([pa.array([r[i] for r in rows], type=pa.int32()) for i in range(6)], names=list("abcdef"))
    assert canon(_join(l2, r2, [("b1", "b2")], "LEFT", impl)) == canon(tab(matched + l_only))
    # join_right_multiple_batches swaps the two tables: the right side is the 7-row one
    swapped = lambda rows: [r[3:] + r[:3] for r in rows]
    assert canon(_join(r2, l2, [("b2", "b1")], "RIGHT", impl)) == canon(tab(swapped(matched + l_only)))
    assert canon(_join(l2, r2, [("b1", "b2")], "FULL", impl)) == canon(tab(matched + l_only + r_only))
    got = _join(l2, r2, [("b1", "b2")], "EXISTENCE", impl)
    assert canon(got) == [(0, 3, 4, False), (1, 4, 5, True), (2, 5, 6, False), (3, 6, 7, True), (4, 6, 8, True), (5, 7, 9, False), (6, 9, 9, False)]


@pytest.mark.parametrize("jt", ["INNER", "LEFT", "RIGHT", "FULL", "SEMI", "ANTI"])
def test_join_fuzz_vs_arrow(jt):
    rng = np.random.default_rng(17)
    nl, nr = 40_000, 3_000
    left = pa.table({"k": pa.array(rng.integers(0, 2500, nl), type=pa.int32(), mask=rng.random(nl) < 0.04), "lv": pa.array(np.arange(nl), type=pa.int64())})
    right = pa.table({"k": pa.array(rng.integers(0, 3500, nr), type=pa.int32(), mask=rng.random(nr) < 0.04), "rv": pa.array(np.arange(nr), type=pa.int64()),
                      "rs": pa.array([f"s{int(i)}" for i in rng.integers(0, 100, nr)])})
    how = {"INNER": "inner", "LEFT": "left outer", "RIGHT": "right outer", "FULL": "full outer", "SEMI": "left semi", "ANTI": "left anti"}[jt]
    exp = left.join(right, keys="k", join_type=how, coalesce_keys=False, right_suffix="_r")
    for impl in ("shjR", "shjL", "smj"):
        got = _join(left, right, [("k", "k")], jt, impl)
        if jt in ("SEMI", "ANTI"):
            e = exp.select(["k", "lv"])
        else:
            e = exp.select(["k", "lv", "k_r", "rv", "rs"])
        assert_same_rows(got, e)


@pytest.mark.parametrize("jt", ["INNER", "LEFT", "RIGHT", "FULL", "SEMI", "ANTI", "EXISTENCE"])
def test_sort_merge_join_streams_sorted_inputs_in_pieces(jt, monkeypatch):
    monkeypatch.setenv("AURON_GPU_CHUNK_ROWS", "2500")     # the FFI reader coalesces exported batches up to this many rows
    # sort_merge_join_exec.rs:294-372: both sides arrive SORTED and in many batches; key groups straddle batch boundaries on both
    # sides, NULL keys lead both streams.  The operator cuts the streams into key-disjoint pieces (metric `pieces`), never holds
    # more than a piece, and its output keeps the order of the driving side (left; right for RIGHT OUTER).
    rng = np.random.default_rng(31)
    nl, nr = 30_000, 12_000
    lk = np.sort(rng.integers(0, 4000, nl)).astype(np.int32)
    rk = np.sort(rng.integers(500, 4500, nr)).astype(np.int32)
    lnull, rnull = np.arange(nl) < 700, np.arange(nr) < 300            # nulls first
    left = pa.table({"k": pa.array(lk, mask=lnull), "d": pa.array((lk % 3).astype(np.int64)), "lv": pa.array(np.arange(nl), type=pa.int64())})
    right = pa.table({"k": pa.array(rk, mask=rnull), "rv": pa.array(np.arange(nr), type=pa.int64())})
    fields = [pa.field(f"l_{f.name}", f.type) for f in left.schema]
    if jt == "EXISTENCE":
        fields.append(pa.field("exists", pa.bool_()))
    elif jt not in ("SEMI", "ANTI"):
        fields += [pa.field(f"r_{f.name}", f.type) for f in right.schema]
    plan = P.sort_merge_join(pa.schema(fields), P.ffi_reader(left.schema, "l"), P.ffi_reader(right.schema, "r"), [(P.col("k"), P.col("k"))], jt)
    td = P.task_definition(plan)
    with runtime.Task(td, {"l": batches(left, 1_000), "r": batches(right, 700)}) as task:
        got = pa.Table.from_batches(list(task), schema=task.schema)
        met = {(op, name): v for _, op, name, v in task.metrics()}
    assert met[("SortMergeJoinExec", "pieces")] >= 4, met
    how = {"INNER": "inner", "LEFT": "left outer", "RIGHT": "right outer", "FULL": "full outer", "SEMI": "left semi", "ANTI": "left anti", "EXISTENCE": "left outer"}[jt]
    exp = left.join(right, keys="k", join_type=how, coalesce_keys=False, right_suffix="_r")
    if jt in ("SEMI", "ANTI"):
        assert_same_rows(got, exp.select(["k", "d", "lv"]))
    elif jt == "EXISTENCE":
        matched = set(exp.filter(pc.is_valid(exp["rv"]))["lv"].to_pylist())
        assert sorted(zip(got["l_lv"].to_pylist(), got["exists"].to_pylist())) == [(i, i in matched) for i in range(nl)]
    else:
        assert_same_rows(got, exp.select(["k", "d", "lv", "k_r", "rv"]))
    # output order = order of the driving side's key (NULL keys first), as Spark's SortMergeJoinExec.outputOrdering promises
    if jt in ("INNER", "LEFT", "SEMI", "ANTI", "EXISTENCE", "RIGHT"):
        col = got.column(3 if jt == "RIGHT" else 0).to_pylist()
        keyed = [(-1 if v is None else v) for v in col]
        assert keyed == sorted(keyed)


def test_join_string_and_multi_key_general_path():
    rng = np.random.default_rng(23)
    n = 20_000
    ks = [f"key-{int(i)}" for i in rng.integers(0, 500, n)]
    left = pa.table({"s": pa.array(ks, mask=rng.random(n) < 0.02), "d": pa.array(rng.integers(0, 4, n), type=pa.int64()), "lv": pa.array(np.arange(n))})
    right = pa.table({"s": pa.array([f"key-{i}" for i in range(0, 600, 2)] * 2), "d": pa.array([0, 1] * 300, type=pa.int64()), "rv": pa.array(np.arange(600))})
    exp = left.join(right, keys=["s", "d"], join_type="inner", coalesce_keys=False, right_suffix="_r").select(["s", "d", "lv", "s_r", "d_r", "rv"])
    got = _join(left, right, [("s", "s"), ("d", "d")], "INNER", "shjR")
    assert_same_rows(got, exp)


# =============================================================================== sort (S1)
def test_sort_golden_limit_and_offset():
    # datafusion-ext-plans/src/sort_exec.rs:1511-1578
    t = table_i32(a=[9, 8, 7, 6, 5, 4, 3, 2, 1, 0], b=[0, 1, 2, 3, 4, 5, 6, 7, 8, 9], c=[5, 6, 7, 8, 9, 0, 1, 2, 3, 4])
    src = lambda: P.ffi_reader(t.schema, "t")
    got = run(P.sort(src(), [P.sort_expr(P.col("a"), True, True)], limit=6), {"t": t})
    assert list(zip(*[c.to_pylist() for c in got.columns])) == [(0, 9, 4), (1, 8, 3), (2, 7, 2), (3, 6, 1), (4, 5, 0), (5, 4, 9)]
    got = run(P.sort(src(), [P.sort_expr(P.col("a"), True, True)], limit=8, offset=3), {"t": t})
    assert list(zip(*[c.to_pylist() for c in got.columns])) == [(3, 6, 1), (4, 5, 0), (5, 4, 9), (6, 3, 8), (7, 2, 7)]


@pytest.mark.parametrize("n", [1, 100, 5000, 1_234_567 // 8])
def test_sort_fuzz_vs_arrow(n):
    # fuzz test of sort_exec.rs:1582-1698 (utf8 + u32 keys, nullable) against an independent sorter
    rng = np.random.default_rng(n)
    words = ["", "a", "ab", "abc", "b", "ba", "天", "天地", "zzzzzzzzzzzz", "zzzzzzzzzzzzz"]
    t = pa.table({"s": pa.array([None if x else words[int(i)] for x, i in zip(rng.random(n) < 0.1, rng.integers(0, len(words), n))]),
                  "u": pa.array(rng.integers(0, 50, n), type=pa.int32(), mask=rng.random(n) < 0.1),
                  "f": pa.array(rng.standard_normal(n), mask=rng.random(n) < 0.05),
                  "l": pa.array(rng.integers(-2**62, 2**62, n), type=pa.int64()),
                  "row": pa.array(np.arange(n), type=pa.int64())})
    cases = [
        ([("s", True, True), ("u", True, True)], [("s", "ascending"), ("u", "ascending")], "at_start"),
        ([("u", False, False), ("f", True, False)], [("u", "descending"), ("f", "ascending")], "at_end"),
        ([("f", False, True)], [("f", "descending")], "at_start"),
        ([("l", True, True)], [("l", "ascending")], "at_start"),
        ([("s", False, False), ("l", False, False)], [("s", "descending"), ("l", "descending")], "at_end"),
    ]
    for keys, akeys, placement in cases:
        plan = P.sort(P.ffi_reader(t.schema, "t"), [P.sort_expr(P.col(k), asc, nf) for k, asc, nf in keys])
        got = run(plan, {"t": t}, chunk=max(1, n // 3))
        assert got.num_rows == n
        key_names = [k for k, _, _ in keys]
        exp_idx = pc.sort_indices(t, sort_keys=akeys, null_placement=placement)
        exp = t.take(exp_idx)
        # ties are unordered in the reference (sort_exec.rs:648-662): compare the key sequence, then the multiset of rows
        assert got.select(key_names).to_pylist() == exp.select(key_names).to_pylist(), keys
        assert sorted(got["row"].to_pylist()) == list(range(n))


@pytest.mark.parametrize("spill", [False, True])
def test_external_sort_runs_spill_and_range_merge(monkeypatch, spill):
    # ExternalSorter (sort_exec.rs:390-447,637-768,913-1061): the input is sorted in runs, runs beyond the memory budget are spilled
    # (pinned host memory here), the output merges them.  Runs of 7,000 rows over 100,000 input rows (15 runs, ~30 key ranges), with
    # a budget that spills every run in the second variant; the fuzz oracle of sort_exec.rs:1617-1697: equality with an independent sorter.
    monkeypatch.setenv("AURON_SORT_RUN_ROWS", "7000")
    if spill:
        monkeypatch.setenv("AURON_SORT_SPILL_BYTES", "1")
    n = 100_000
    rng = np.random.default_rng(11)
    t = pa.table({"u": pa.array(rng.integers(0, 300, n), type=pa.int32(), mask=rng.random(n) < 0.05),
                  "f": pa.array(rng.standard_normal(n), mask=rng.random(n) < 0.05),
                  "d": pa.array(rng.integers(-2**40, 2**40, n), type=pa.int64()).cast(pa.decimal128(20, 0)),
                  "s": pa.array([f"payload-{int(i)}" for i in rng.integers(0, 1000, n)], mask=rng.random(n) < 0.1),
                  "row": pa.array(np.arange(n), type=pa.int64())})
    cases = [([("u", True, True), ("f", False, False)], [("u", "ascending"), ("f", "descending")], None),
             ([("d", False, True)], [("d", "descending")], "at_start"),
             ([("f", True, True), ("row", True, True)], [("f", "ascending"), ("row", "ascending")], "at_start")]
    for keys, akeys, placement in cases:
        td = P.task_definition(P.sort(P.ffi_reader(t.schema, "t"), [P.sort_expr(P.col(k), asc, nf) for k, asc, nf in keys]))
        with runtime.Task(td, {"t": batches(t, 9_000)}) as task:
            got = pa.Table.from_batches(list(task), schema=task.schema)
            met = {(op, name): v for _, op, name, v in task.metrics()}
        assert met[("SortExec", "sorted_runs")] >= 14 and met[("SortExec", "merge_ranges")] >= 10, met
        assert (met.get(("SortExec", "mem_spill_count"), 0) > 0) == spill
        key_names = [k for k, _, _ in keys]
        if placement is None:   # mixed null placement: build the expected order key by key
            tt = t.to_pandas()
            exp = t.take(pa.array(tt.sort_values(by=["u", "f"], ascending=[True, False], na_position="first", kind="stable").index.to_numpy()))
            u_first = exp["u"].to_pylist()
            assert got["u"].to_pylist() == u_first     # u ascending, NULLs first
            # f descending with NULLs last inside every u group
            gu, gf = got["u"].to_pylist(), got["f"].to_pylist()
            for i in range(1, n):
                if gu[i] == gu[i - 1]:
                    a, b = gf[i - 1], gf[i]
                    assert (b is None) or (a is not None and a >= b), (i, a, b)
        else:
            exp = t.take(pc.sort_indices(t, sort_keys=akeys, null_placement=placement))
            assert got.select(key_names).to_pylist() == exp.select(key_names).to_pylist(), keys
        assert sorted(got["row"].to_pylist()) == list(range(n))
        rows = dict(zip(got["row"].to_pylist(), got["s"].to_pylist()))
        assert all(rows[i] == v for i, v in enumerate(t["s"].to_pylist()))     # payload travelled with its row
    # limit / offset across runs
    td = P.task_definition(P.sort(P.ffi_reader(t.schema, "t"), [P.sort_expr(P.col("row"), False, True)], limit=20_000, offset=19_990))
    with runtime.Task(td, {"t": batches(t, 9_000)}) as task:
        got = pa.Table.from_batches(list(task), schema=task.schema)
    assert got["row"].to_pylist() == list(range(n - 1 - 19_990, n - 1 - 20_000, -1))


# =============================================================================== misc operators
def test_limit_union_rename_and_metrics():
    t = table_i32(a=list(range(100)))
    got = run(P.limit(P.ffi_reader(t.schema, "t"), 30, 10), {"t": t}, chunk=7)
    assert got["a"].to_pylist() == list(range(10, 30))
    got = run(P.rename_columns(P.union([P.ffi_reader(t.schema, "t"), P.ffi_reader(t.schema, "u")], t.schema), ["z"]), {"t": t, "u": t})
    assert got.schema.names == ["z"] and got.num_rows == 200
    td = P.task_definition(P.filter_(P.ffi_reader(t.schema, "t"), [P.binary("Lt", P.col("a"), P.lit(5, pa.int32()))]))
    with runtime.Task(td, {"t": batches(t)}) as task:
        rows = sum(b.num_rows for b in task)
        m = task.metrics()
    assert rows == 5 and ("FilterExec", "output_rows", 5) in [(op, name, v) for _, op, name, v in m]


def test_errors_are_reported_not_thrown():
    t = table_i32(a=[1])
    with pytest.raises(runtime.AuronError):
        run(P.filter_(P.ffi_reader(t.schema, "t"), [P.binary("Gt", P.col("nope"), P.lit(1, pa.int32()))]), {"t": t})
    with pytest.raises(runtime.AuronError):
        runtime.run_task(b"\x12\x03\xff\xff\xff", {})


def test_concurrent_tasks_share_the_library():
    # Spark runs several tasks of an executor concurrently in one process: every native task has its own stream and context,
    # the process-wide pools (pinned staging, worker threads, staged uploads, device landing buffers) are shared
    import concurrent.futures
    rng = np.random.default_rng(1)
    n = 200_000
    tables = [pa.table({"k": pa.array(rng.integers(0, 1000, n), type=pa.int32()), "v": pa.array(rng.integers(-50, 50, n), type=pa.int64(),
                                                                                                 mask=rng.random(n) < 0.1),
                        "s": pa.array([f"w{int(x)}" for x in rng.integers(0, 50, n)])}) for _ in range(6)]

    def job(i):
        t = tables[i]
        flt = P.filter_(P.ffi_reader(t.schema, "t"), [P.binary("GtEq", P.col("k"), P.lit(100 * (i % 3), pa.int32())), P.like(P.col("s"), P.lit("w1%", pa.string()))])
        plan = P.agg(flt, [P.col("k")], ["k"], [P.agg_expr("SUM", [P.col("v")], pa.int64()), P.agg_expr("COUNT", [P.col("v")], pa.int64())], ["s", "c"], ["PARTIAL"] * 2)
        out = []
        for _ in range(3):
            out.append(run(plan, {"t": t}, chunk=50_000))
        return out

    with concurrent.futures.ThreadPoolExecutor(max_workers=6) as ex:
        results = list(ex.map(job, range(6)))
    for i, outs in enumerate(results):
        t = tables[i]
        keep = pc.and_kleene(pc.greater_equal(t["k"], 100 * (i % 3)), pc.starts_with(t["s"], "w1"))
        ft = t.filter(keep)
        exp = oracle.agg_sum_count_i64(ft["k"].combine_chunks().cast(pa.int64()), ft["v"].combine_chunks())
        for got in outs:
            got = pa.table({"k": got.column(0).cast(pa.int64()), "s": got.column(1), "c": got.column(2)})
            assert_same_rows(got, exp)


@pytest.mark.parametrize("keys", [["k"], ["k", "g"], []])
def test_agg_min_max_on_strings(keys):
    # AggMaxMin over utf8 (maxmin.rs:100-296): byte-wise order, NULLs ignored, all-NULL groups give NULL; partial -> final merge
    rng = np.random.default_rng(12)
    n = 60_000
    words = ["", "a", "ab", "abc", "b", "zz", "z", "天地", "天", "Ab", "~"]
    t = pa.table({"k": pa.array(rng.integers(0, 300, n), type=pa.int32(), mask=rng.random(n) < 0.02),
                  "g": pa.array([["x", "y", None][int(i)] for i in rng.integers(0, 3, n)]),
                  "s": pa.array([words[int(i)] + str(int(j)) for i, j in zip(rng.integers(0, len(words), n), rng.integers(0, 50, n))], mask=rng.random(n) < 0.3),
                  "f": pa.array(rng.integers(0, 10, n), type=pa.int32())})
    # group 299 only has NULL strings
    s = t["s"].to_pylist()
    k = t["k"].to_pylist()
    s = [None if kk == 299 else v for kk, v in zip(k, s)]
    t = t.set_column(2, "s", pa.array(s, type=pa.string()))
    flt = P.filter_(P.ffi_reader(t.schema, "t"), [P.binary("Lt", P.col("f"), P.lit(7, pa.int32()))])
    part = P.agg(flt, [P.col(c) for c in keys], keys, [P.agg_expr("MIN", [P.col("s")], pa.string()), P.agg_expr("MAX", [P.col("s")], pa.string())],
                 ["mn", "mx"], ["PARTIAL"] * 2)
    final = P.agg(part, [P.col(c) for c in keys], keys, [P.agg_expr("MIN", [P.lit(None, pa.null())], pa.string()), P.agg_expr("MAX", [P.lit(None, pa.null())], pa.string())],
                  ["mn", "mx"], ["FINAL"] * 2)
    got = run(final, {"t": t}, chunk=17_000)
    acc = {}
    cols = [t[c].to_pylist() for c in keys]
    for i, (v, f) in enumerate(zip(s, t["f"].to_pylist())):
        if f >= 7:
            continue
        key = tuple(c[i] for c in cols)
        a = acc.setdefault(key, [None, None])
        if v is not None:
            b = v.encode()
            if a[0] is None or b < a[0]:
                a[0] = b
            if a[1] is None or b > a[1]:
                a[1] = b
    if not keys and not acc:
        acc[()] = [None, None]
    exp_rows = sorted([key + (None if a[0] is None else a[0].decode(), None if a[1] is None else a[1].decode()) for key, a in acc.items()], key=repr)
    got_rows = sorted(zip(*[c.to_pylist() for c in got.columns]), key=repr)
    assert got_rows == exp_rows


@pytest.mark.parametrize("mode", ["PARTIAL", "FINAL"])
def test_agg_spills_to_host_buckets_and_merges_them(mode, monkeypatch):
    # A5 (agg_table.rs:99-135,323-353,474-721): with a budget of one byte every chunk's table is spilled into hash buckets held in
    # host memory; the output merges bucket by bucket and must equal the unspilled result, one row per group
    rng = np.random.default_rng(31)
    n = 120_000
    t = pa.table({"k": pa.array(rng.integers(0, 30_000, n), type=pa.int64(), mask=rng.random(n) < 0.01),
                  "s": pa.array([f"g{int(i)}" for i in rng.integers(0, 7, n)]),
                  "v": pa.array(rng.integers(-10**6, 10**6, n), type=pa.int64(), mask=rng.random(n) < 0.05),
                  "d": pa.array([decimal.Decimal(int(x)) / 100 for x in rng.integers(0, 10**6, n)], type=pa.decimal128(7, 2)),
                  "w": pa.array([f"w{int(i):05d}" for i in rng.integers(0, 50_000, n)])})

    def plan():
        src = P.ffi_reader(t.schema, "t")
        part = P.agg(src, [P.col("k"), P.col("s")], ["k", "s"],
                     [P.agg_expr("SUM", [P.col("v")], pa.int64()), P.agg_expr("COUNT", [P.col("v")], pa.int64()),
                      P.agg_expr("AVG", [P.col("d")], pa.decimal128(11, 6)), P.agg_expr("MIN", [P.col("w")], pa.string()),
                      P.agg_expr("MAX", [P.col("v")], pa.int64())], ["sv", "c", "ad", "mw", "xv"], ["PARTIAL"] * 5)
        if mode == "PARTIAL":
            return part
        N = pa.null()
        return P.agg(part, [P.col("k"), P.col("s")], ["k", "s"],
                     [P.agg_expr("SUM", [P.lit(None, N)], pa.int64()), P.agg_expr("COUNT", [P.lit(None, N)], pa.int64()),
                      P.agg_expr("AVG", [P.lit(None, N)], pa.decimal128(11, 6)), P.agg_expr("MIN", [P.lit(None, N)], pa.string()),
                      P.agg_expr("MAX", [P.lit(None, N)], pa.int64())], ["sv", "c", "ad", "mw", "xv"], ["FINAL"] * 5)

    exp = run(plan(), {"t": t}, chunk=25_000)
    monkeypatch.setenv("AURON_AGG_SPILL_BYTES", "1")
    monkeypatch.setenv("AURON_GPU_CHUNK_ROWS", "30000")               # several device chunks, so several spills
    td = P.task_definition(plan())
    with runtime.Task(td, {"t": batches(t, 25_000)}) as task:
        out = list(task)
        m = {(op, name): v for _, op, name, v in task.metrics()}
    got = pa.Table.from_batches(out, schema=exp.schema)
    assert m[("AggExec", "mem_spill_count")] >= 3 and m[("AggExec", "mem_spill_size")] > 0
    assert len(out) > 8                                                # one output batch per non-empty bucket
    assert_same_rows(got, exp)
    keys = list(zip(got["k"].to_pylist(), got["s"].to_pylist()))
    assert len(keys) == len(set(keys))                                 # every group exactly once: buckets partition the key space


@pytest.mark.parametrize("key_type,span", [(pa.int32(), 5_000), (pa.int64(), 10**15), (pa.date32(), 30_000)])
@pytest.mark.parametrize("side", ["LEFT", "RIGHT"])
def test_inner_join_on_unique_build_keys_passes_the_probe_side_through_a_mask(key_type, span, side, monkeypatch):
    # A dimension table joined on its primary key: the probe is one pass into a partner index + match mask (direct-address table for
    # small-range integer keys, the hashed table otherwise) and the probe columns reach the consumers unmaterialised -- a projection,
    # an aggregate, and the plain materialised output must all equal the pair-list path (AURON_JOIN_NO_MASK=1) and pandas.
    rng = np.random.default_rng(31)
    nb, n = 4_000, 120_000
    keys = rng.choice(span, nb, replace=False).astype(np.int64)
    def as_key(a, mask=None):
        base = pa.int64() if key_type == pa.int64() else pa.int32()
        return pa.array(a.astype(np.int64 if base == pa.int64() else np.int32), type=base, mask=mask).cast(key_type)

    dim = pa.table({"id": as_key(keys), "name": pa.array([f"n{int(k) % 101}" for k in keys]),
                    "w": pa.array((keys % 1000).astype(np.int64), mask=keys % 13 == 0)})
    probe_keys = np.where(rng.random(n) < 0.7, rng.choice(keys, n), rng.integers(0, span, n)).astype(np.int64)
    fact = pa.table({"k": as_key(probe_keys, rng.random(n) < 0.03),
                     "v": pa.array(rng.integers(-50, 50, n), type=pa.int64(), mask=rng.random(n) < 0.05),
                     "s": pa.array([f"s{int(x)}" for x in rng.integers(0, 9, n)])})
    fs, ds = P.ffi_reader(fact.schema, "f"), P.ffi_reader(dim.schema, "d")
    if side == "RIGHT":   # build side on the right: output [fact..., dim...]
        js = pa.schema(list(fact.schema) + list(dim.schema))
        join = P.hash_join(js, fs, ds, [(P.col("k"), P.col("id"))], "INNER", "RIGHT")
    else:
        js = pa.schema(list(dim.schema) + list(fact.schema))
        join = P.hash_join(js, ds, fs, [(P.col("id"), P.col("k"))], "INNER", "LEFT")
    plans = {
        "rows": join,
        "project": P.projection(join, [P.col("s"), P.binary("Plus", P.col("v"), P.col("w")), P.col("name")], ["s", "vw", "name"], [pa.string(), pa.int64(), pa.string()]),
        "agg": P.agg(join, [P.col("name")], ["name"], [P.agg_expr("SUM", [P.col("v")], pa.int64()), P.agg_expr("COUNT", [P.col("w")], pa.int64())], ["sv", "cw"], ["PARTIAL"] * 2),
    }
    inputs = lambda: {"f": batches(fact, 25_000), "d": batches(dim)}
    got = {name: runtime.run_task(P.task_definition(pl), inputs()) for name, pl in plans.items()}
    monkeypatch.setenv("AURON_JOIN_NO_MASK", "1")
    ref = {name: runtime.run_task(P.task_definition(pl), inputs()) for name, pl in plans.items()}
    for name in plans:
        assert got[name].schema == ref[name].schema
        assert canon(got[name]) == canon(ref[name]), name
    m = fact.to_pandas().merge(dim.to_pandas(), left_on="k", right_on="id")
    assert got["rows"].num_rows == len(m) > 0.5 * n
    exp = m.groupby("name").agg(sv=("v", "sum"), cw=("w", "count"))
    agg = {r[0]: (r[1], r[2]) for r in zip(*[got["agg"].column(i).to_pylist() for i in range(3)])}
    for name_, row in exp.iterrows():
        sv = None if m[m.name == name_].v.notna().sum() == 0 else int(row.sv)
        assert agg[name_] == (sv, int(row.cw)), name_


def test_device_wide_memory_budget_makes_aggregate_and_sort_spill():
    # auron-memmgr/src/lib.rs:201-423: one budget shared by every spillable consumer.  With a tiny HBM budget (no per-operator
    # override) the aggregate spills its partial tables into host buckets and the sorter spills runs, and both still return the
    # results of the unconstrained run.
    rng = np.random.default_rng(44)
    n = 150_000
    t = pa.table({"k": pa.array(rng.integers(0, 40_000, n), type=pa.int64(), mask=rng.random(n) < 0.01),
                  "v": pa.array(rng.integers(-10**6, 10**6, n), type=pa.int64(), mask=rng.random(n) < 0.05),
                  "w": pa.array([f"w{int(i):05d}" for i in rng.integers(0, 50_000, n)])})
    src = lambda: P.ffi_reader(t.schema, "t")
    agg = lambda: P.agg(src(), [P.col("k")], ["k"], [P.agg_expr("SUM", [P.col("v")], pa.int64()), P.agg_expr("MAX", [P.col("w")], pa.string())], ["sv", "mw"], ["PARTIAL"] * 2)
    srt = lambda: P.sort(src(), [P.sort_expr(P.col("v"), True, True), P.sort_expr(P.col("k"), False, False)])

    def go(plan):
        os.environ["AURON_GPU_CHUNK_ROWS"] = "30000"
        os.environ["AURON_SORT_RUN_ROWS"] = "30000"
        try:
            with runtime.Task(P.task_definition(plan), {"t": batches(t, 25_000)}) as task:
                out = pa.Table.from_batches(list(task), schema=task.schema)
                return out, {name: v for _, _, name, v in task.metrics()}
        finally:
            os.environ.pop("AURON_GPU_CHUNK_ROWS", None)
            os.environ.pop("AURON_SORT_RUN_ROWS", None)

    exp_agg, m0 = go(agg())
    exp_sort, m1 = go(srt())
    assert m0.get("mem_spill_count", 0) == 0 and m1.get("mem_spill_count", 0) == 0
    default = runtime.set_hbm_budget(0)
    assert default > (1 << 30)
    try:
        assert runtime.set_hbm_budget(200_000) == 200_000
        got_agg, ma = go(agg())
        got_sort, ms = go(srt())
    finally:
        assert runtime.set_hbm_budget(0) == default
    assert ma["mem_spill_count"] >= 1 and ms["mem_spill_count"] >= 1
    assert_same_rows(got_agg, exp_agg)
    assert got_sort.column("v").to_pylist() == exp_sort.column("v").to_pylist() and got_sort.column("k").to_pylist() == exp_sort.column("k").to_pylist()
