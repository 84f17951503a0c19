"""CPU oracle for the auron_b200 hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (auron_b200/, libauron_b200.so) never does.

Two layers:
  * libauron_oracle.so (auron_oracle.c): byte/integer arithmetic restated from the
    reference (murmur3/xxhash64, varint, compacted batch serde, Spark casts, partitioner,
    hash aggregate baseline).  Pinned by the reference's own golden vectors
    (tests/test_oracle_golden.py).
  * numpy / pyarrow (Arrow C++ 24) restatements of operator semantics (filter, project,
    aggregate, joins, sort, shuffle partitioning) following the cited reference files.
    Where the reference delegates to arrow-rs kernels that are not under /root/reference
    (arith/compare/like/row-order/parquet decode) parity is "unpinned" at the Rust level
    (SURVEY.md section 8c); Arrow C++ plus the documented Spark deviations is the oracle.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libauron_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "auron_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_murmur3_bytes.restype = C.c_int32
        _lib.orc_murmur3_bytes.argtypes = [C.c_char_p, C.c_int64, C.c_int32]
        _lib.orc_xxhash64_bytes.restype = C.c_int64
        _lib.orc_xxhash64_bytes.argtypes = [C.c_char_p, C.c_int64, C.c_int64]
        _lib.orc_write_len.restype = C.c_int64
        _lib.orc_read_len.restype = C.c_int64
        for f in ("orc_serde_write_fixed", "orc_serde_write_bool", "orc_serde_write_bytes", "orc_serde_read_fixed",
                  "orc_serde_read_bool", "orc_serde_read_bytes", "orc_agg_sum_count_i64"):
            getattr(_lib, f).restype = C.c_int64
        _lib.orc_f64_to_int.restype = C.c_int64
        _lib.orc_f64_to_int.argtypes = [C.c_double, C.c_int]
        _lib.orc_str_to_int.restype = C.c_int
        _lib.orc_str_to_date.restype = C.c_int
    return _lib


def _p(a):
    """numpy array -> void* (None stays NULL)"""
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------- hashing
def murmur3_bytes(b: bytes, seed: int = 42) -> int:
    return lib().orc_murmur3_bytes(b, len(b), seed)


def xxhash64_bytes(b: bytes, seed: int = 42) -> int:
    return lib().orc_xxhash64_bytes(b, len(b), seed)


def _validity_np(arr: pa.Array):
    """validity bitmap re-based to bit 0, or None when the array has no nulls"""
    if arr.null_count == 0:
        return None
    v = pc.is_valid(arr).to_numpy(zero_copy_only=False)
    return np.packbits(v, bitorder="little")


_FIXED = {
    pa.int8(): (1, 4, 1), pa.int16(): (2, 4, 1), pa.int32(): (4, 4, 1), pa.int64(): (8, 8, 1),
    pa.float32(): (4, 4, 0), pa.float64(): (8, 8, 0), pa.date32(): (4, 4, 1), pa.date64(): (8, 8, 1),
}


def _values_np(arr: pa.Array, width: int) -> np.ndarray:
    buf = arr.buffers()[1]
    raw = np.frombuffer(buf, dtype=np.uint8)
    return np.ascontiguousarray(raw[arr.offset * width:(arr.offset + len(arr)) * width])


def hash_columns(cols, kind: str = "murmur3", seed: int = 42) -> np.ndarray:
    """create_murmur3_hashes / create_xxhash64_hashes (spark_hash.rs:28-57): chained over columns."""
    n = len(cols[0]) if cols else 0
    k = 0 if kind == "murmur3" else 1
    h32 = np.full(n, seed, dtype=np.int32)
    h64 = np.full(n, seed, dtype=np.int64)
    L = lib()
    for col in cols:
        if isinstance(col, pa.ChunkedArray):
            col = col.combine_chunks()
        t = col.type
        valid = _validity_np(col)
        if t in _FIXED or pa.types.is_timestamp(t):
            win, wh, sg = _FIXED.get(t, (8, 8, 1))
            vals = _values_np(col, win)
            L.orc_hash_fixed(k, _p(vals), win, wh, sg, _p(valid), C.c_int64(n), _p(h32), _p(h64))
        elif pa.types.is_decimal128(t):
            vals = _values_np(col, 16)
            L.orc_hash_fixed(k, _p(vals), 16, 16, 1, _p(valid), C.c_int64(n), _p(h32), _p(h64))
        elif pa.types.is_boolean(t):
            bits = np.packbits(col.fill_null(False).to_numpy(zero_copy_only=False), bitorder="little")
            L.orc_hash_bool(k, _p(bits), _p(valid), C.c_int64(n), _p(h32), _p(h64))
        elif pa.types.is_string(t) or pa.types.is_binary(t):
            offs = np.frombuffer(col.buffers()[1], dtype=np.int32)[col.offset:col.offset + n + 1].copy()
            data = np.frombuffer(col.buffers()[2], dtype=np.uint8) if col.buffers()[2] is not None else np.zeros(1, np.uint8)
            L.orc_hash_bytes(k, _p(offs), _p(data), _p(valid), C.c_int64(n), _p(h32), _p(h64))
        else:
            raise NotImplementedError(str(t))
    return h32 if k == 0 else h64


def partition_ids(cols, num_parts: int, seed: int = 42) -> np.ndarray:
    """evaluate_partition_ids (shuffle/mod.rs:163-188)"""
    h = hash_columns(cols, "murmur3", seed)
    out = np.empty(len(h), dtype=np.int32)
    lib().orc_pmod(_p(h), C.c_int64(len(h)), C.c_int32(num_parts), _p(out))
    return out


# ---------------------------------------------------------------- serde
def write_len(n: int) -> bytes:
    buf = (C.c_uint8 * 16)()
    k = lib().orc_write_len(C.c_uint64(n), buf)
    return bytes(buf[:k])


def read_len(b: bytes, pos: int = 0):
    out = C.c_uint64()
    k = lib().orc_read_len(C.c_char_p(b[pos:pos + 16]), C.byref(out))
    return out.value, pos + k


def serde_write_batch(batch: pa.RecordBatch) -> bytes:
    """write_batch (batch_serde.rs:68-79): varint num_rows | column*"""
    n = batch.num_rows
    out = bytearray(write_len(n))
    L = lib()
    for col in batch.columns:
        t = col.type
        valid = _validity_np(col)
        if pa.types.is_boolean(t):
            bits = np.packbits(col.fill_null(False).to_numpy(zero_copy_only=False), bitorder="little")
            buf = np.empty(n // 4 + 32, dtype=np.uint8)
            k = L.orc_serde_write_bool(_p(bits), C.c_int64(0), _p(valid), C.c_int64(0), C.c_int64(n), _p(buf))
        elif pa.types.is_string(t) or pa.types.is_binary(t):
            offs = np.frombuffer(col.buffers()[1], dtype=np.int32)[col.offset:col.offset + n + 1].copy()
            data = np.frombuffer(col.buffers()[2], dtype=np.uint8) if col.buffers()[2] is not None else np.zeros(1, np.uint8)
            buf = np.empty(n * 5 + int(offs[-1] - offs[0]) + 32, dtype=np.uint8)
            k = L.orc_serde_write_bytes(_p(offs), _p(data), _p(valid), C.c_int64(0), C.c_int64(n), _p(buf))
        else:
            w = 16 if pa.types.is_decimal128(t) else t.bit_width // 8
            vals = _values_np(col, w)
            buf = np.empty(n * w + n // 8 + 32, dtype=np.uint8)
            k = L.orc_serde_write_fixed(_p(vals), w, _p(valid), C.c_int64(0), C.c_int64(n), _p(buf))
        out += buf[:k].tobytes()
    return bytes(out)


def serde_read_batch(b: bytes, schema: pa.Schema, pos: int = 0):
    """read_batch (batch_serde.rs:81-101). Returns (RecordBatch | None, new_pos)."""
    if pos >= len(b):
        return None, pos
    n, pos = read_len(b, pos)
    L = lib()
    src = np.frombuffer(b, dtype=np.uint8)
    cols = []
    for f in schema:
        t = f.type
        validity = np.zeros((n + 7) // 8 + 1, dtype=np.uint8)
        has_nulls = C.c_int()
        inp = src[pos:]
        if pa.types.is_boolean(t):
            bits = np.zeros((n + 7) // 8 + 1, dtype=np.uint8)
            k = L.orc_serde_read_bool(_p(inp), C.c_int64(n), _p(bits), _p(validity), C.byref(has_nulls))
            bufs = [pa.py_buffer(validity.tobytes()) if has_nulls.value else None, pa.py_buffer(bits.tobytes())]
        elif pa.types.is_string(t) or pa.types.is_binary(t):
            offs = np.zeros(n + 1, dtype=np.int32)
            dlen = C.c_int64()
            L.orc_serde_read_bytes(_p(inp), C.c_int64(n), _p(offs), None, C.byref(dlen), _p(validity), C.byref(has_nulls))
            data = np.zeros(max(dlen.value, 1), dtype=np.uint8)
            k = L.orc_serde_read_bytes(_p(inp), C.c_int64(n), _p(offs), _p(data), C.byref(dlen), _p(validity), C.byref(has_nulls))
            bufs = [pa.py_buffer(validity.tobytes()) if has_nulls.value else None, pa.py_buffer(offs.tobytes()),
                    pa.py_buffer(data[:dlen.value].tobytes())]
        else:
            w = 16 if pa.types.is_decimal128(t) else t.bit_width // 8
            vals = np.zeros(max(n * w, 1), dtype=np.uint8)
            k = L.orc_serde_read_fixed(_p(inp), w, C.c_int64(n), _p(vals), _p(validity), C.byref(has_nulls))
            bufs = [pa.py_buffer(validity.tobytes()) if has_nulls.value else None, pa.py_buffer(vals[:n * w].tobytes())]
        cols.append(pa.Array.from_buffers(t, n, bufs))
        pos += k
    return pa.RecordBatch.from_arrays(cols, schema=schema), pos


# ---------------------------------------------------------------- casts
def str_to_int(s: bytes, bits: int):
    out = C.c_int64()
    ok = lib().orc_str_to_int(C.c_char_p(s), C.c_int64(len(s)), C.c_int(bits), C.byref(out))
    return out.value if ok else None


def str_to_date(s: bytes):
    out = C.c_int32()
    ok = lib().orc_str_to_date(C.c_char_p(s), C.c_int64(len(s)), C.byref(out))
    return out.value if ok else None


def f64_to_int(v: float, bits: int) -> int:
    return lib().orc_f64_to_int(v, bits)


# ---------------------------------------------------------------- operators (semantic oracles)
def agg_sum_count_i64(keys: pa.Array, vals: pa.Array, pred: np.ndarray | None = None):
    """GROUP BY int64 key, SUM(int64)/COUNT -- C port timed as the CPU baseline.
    Returns a pyarrow Table(k, sum, cnt) (group order unspecified, agg_table.rs:177-205)."""
    n = len(keys)
    k = np.ascontiguousarray(keys.fill_null(0).to_numpy(zero_copy_only=False).astype(np.int64))
    v = np.ascontiguousarray(vals.fill_null(0).to_numpy(zero_copy_only=False).astype(np.int64))
    kv, vv = _validity_np(keys), _validity_np(vals)
    cap = n + 1
    ok, okv = np.empty(cap, np.int64), np.empty(cap, np.uint8)
    osum, osv, ocnt = np.empty(cap, np.int64), np.empty(cap, np.uint8), np.empty(cap, np.int64)
    pr = None if pred is None else np.ascontiguousarray(pred.astype(np.uint8))
    g = lib().orc_agg_sum_count_i64(_p(k), _p(kv), _p(v), _p(vv), _p(pr), C.c_int64(n), _p(ok), _p(okv), _p(osum), _p(osv),
                                    _p(ocnt), C.c_int64(cap))
    return pa.table({
        "k": pa.array(ok[:g], mask=okv[:g] == 0),
        "sum": pa.array(osum[:g], mask=osv[:g] == 0),
        "cnt": pa.array(ocnt[:g]),
    })


def partition_rows(part_ids: np.ndarray, num_parts: int):
    n = len(part_ids)
    rows = np.empty(n, np.int32)
    offs = np.empty(num_parts + 1, np.int64)
    lib().orc_partition_rows(_p(np.ascontiguousarray(part_ids.astype(np.int32))), C.c_int64(n), C.c_int32(num_parts), _p(rows), _p(offs))
    return rows, offs


def sort_table_canonical(t: pa.Table) -> pa.Table:
    """Order-insensitive comparison helper: sort by all columns (nulls first)."""
    if t.num_rows == 0:
        return t
    idx = pc.sort_indices(t, sort_keys=[(n, "ascending") for n in t.column_names], null_placement="at_start")
    return t.take(idx)


def window_functions(rows, partition_of, order_of, specs):
    """WindowExec's processors restated row by row (datafusion-ext-plans/src/window/processors/*.rs) over `rows` that are already in
    window order.  partition_of(row) / order_of(row) give the partition / order key; specs = [(function, argument getter or None, extra)]:
      ROW_NUMBER (row_number_processor.rs:37-66), RANK / DENSE_RANK (rank_processor.rs:47-80),
      SUM / COUNT / MIN / MAX / AVG: the accumulator after every row (agg_processor.rs:49-93),
      PERCENT_RANK (percent_rank_processor.rs:40-86), CUME_DIST (cume_dist_processor.rs:40-81),
      LEAD with extra = (offset, default getter) (lead_processor.rs:38-109), NTH_VALUE / NTH_VALUE_IGNORE_NULLS with extra = n
      (nth_value_processor.rs:82-120).
    Returns one tuple per row."""
    n = len(rows)
    out = [[None] * len(specs) for _ in range(n)]
    i = 0
    while i < n:
        j = i
        while j < n and partition_of(rows[j]) == partition_of(rows[i]):
            j += 1
        part = rows[i:j]
        size = len(part)
        rank = dense = 0
        equals = 1
        acc = [dict(s=None, c=0, mn=None, mx=None, nth=None, seen=0) for _ in specs]
        k = 0
        peer_end = 0
        for q, row in enumerate(part):
            if q == 0 or order_of(row) != order_of(part[q - 1]):
                rank += equals if q else 1
                dense += 1
                equals = 1
                peer_end = q
                while peer_end < size and order_of(part[peer_end]) == order_of(row):
                    peer_end += 1
            else:
                equals += 1
            for c, (fn, arg, extra) in enumerate(specs):
                a = acc[c]
                v = arg(row) if arg else None
                if fn == "ROW_NUMBER":
                    r = q + 1
                elif fn == "RANK":
                    r = rank
                elif fn == "DENSE_RANK":
                    r = dense
                elif fn == "PERCENT_RANK":
                    r = 0.0 if size <= 1 else (rank - 1) / (size - 1)
                elif fn == "CUME_DIST":
                    r = peer_end / size
                elif fn == "LEAD":
                    off, dflt = extra
                    t = q + off
                    r = arg(part[t]) if 0 <= t < size else dflt(row)
                elif fn in ("NTH_VALUE", "NTH_VALUE_IGNORE_NULLS"):
                    if a["nth"] is None and (fn == "NTH_VALUE" or v is not None):
                        a["seen"] += 1
                        if a["seen"] == extra:
                            a["nth"] = (v,)
                    r = a["nth"][0] if a["nth"] is not None else None
                else:
                    if fn == "COUNT" and arg is None:
                        a["c"] += 1
                    elif v is not None:
                        a["s"] = v if a["s"] is None else a["s"] + v
                        a["c"] += 1
                        a["mn"] = v if a["mn"] is None else min(a["mn"], v)
                        a["mx"] = v if a["mx"] is None else max(a["mx"], v)
                    r = {"SUM": a["s"], "COUNT": a["c"], "MIN": a["mn"], "MAX": a["mx"], "AVG": None if a["c"] == 0 or a["s"] is None else a["s"] / a["c"]}[fn]
                out[i + q][c] = r
        i = j
    return [tuple(r) for r in out]


# ---------------------------------------------------------------- expression functions (E4 / rounding)
def spark_round(value, scale: int, kind: str, in_scale: int = 0, half_even: bool = False):
    """spark_round / spark_bround on ONE column value (array branch: datafusion-ext-functions/src/spark_round.rs:62-132,
    spark_bround.rs:59-131).  kind: 'decimal' (value = unscaled int, result = unscaled int at the same scale), 'int16/32/64',
    'f32', 'f64'.  None stays None."""
    if value is None:
        return None

    def int_round(v: int, digits: int) -> int:          # round_i128_half_up / round_i128_half_even with scale = -digits
        if digits <= 0:
            return v
        factor = 10 ** digits
        rem = abs(v) % factor * (1 if v >= 0 else -1)    # Rust `%` keeps the sign of the dividend
        base = v - rem
        twice = abs(rem) * 2
        if half_even:
            if twice > factor:
                return base + factor if v >= 0 else base - factor
            if twice < factor:
                return base
            q = abs(base) // factor
            return base if q % 2 == 0 else (base + factor if v >= 0 else base - factor)
        if twice >= factor:
            return base + factor if v >= 0 else base - factor
        return base

    if kind == "decimal":
        diff = in_scale - scale
        return int_round(value, diff) if diff >= 0 else value * 10 ** (-diff)
    if kind.startswith("int"):
        bits = int(kind[3:])
        r = int_round(int(value), -scale) & ((1 << bits) - 1)               # `as i16/i32/i64` wraps
        return r - (1 << bits) if r >> (bits - 1) else r
    ft = np.float32 if kind == "f32" else np.float64
    x = ft(value)
    if np.isnan(x) or np.isinf(x):
        return float(x)
    f = ft(1)
    for _ in range(abs(scale)):                                             # powi: repeated multiplication, reciprocal for n < 0
        f = ft(f * ft(10))
    if scale < 0:
        f = ft(ft(1) / f)
    y = ft(x * f)
    if half_even:
        ax = abs(y)
        fl = np.floor(ax)
        d = ft(ax - fl)
        r = fl + 1 if d > 0.5 else (fl if d < 0.5 else (fl if int(fl) % 2 == 0 else fl + 1))
        r = np.copysign(ft(r), y)
    else:
        r = np.floor(ft(y + ft(0.5))) if y >= 0 else np.ceil(ft(y - ft(0.5)))
    return float(ft(ft(r) / f))


def spark_time_part(value, unit: str, which: str, zone: str | None = None):
    """spark_hour / minute / second and the date parts with an optional session time zone (spark_dates.rs:200-345): `value` in
    `unit` ('s', 'ms', 'us', 'ns', 'date32') is cast to Timestamp(ms) the way arrow does (division toward zero), shifted by the
    zone's UTC offset at that instant (Python's zoneinfo reads the same tz database chrono-tz compiles in) and decomposed.
    which: hour | minute | second | year | month | day | dayofweek | quarter."""
    import datetime as dt
    import zoneinfo
    if value is None:
        return None
    v = int(value)
    trunc = lambda a, b: abs(a) // b * (1 if a >= 0 else -1)
    ms = {"s": v * 1000, "ms": v, "us": trunc(v, 1000), "ns": trunc(v, 1_000_000), "date32": v * 86_400_000}[unit]
    if zone is not None:
        try:
            tz = zoneinfo.ZoneInfo(zone)
            off = dt.datetime.fromtimestamp(ms // 1000, tz=dt.timezone.utc).astimezone(tz).utcoffset()
            ms += int(off.total_seconds()) * 1000
        except (zoneinfo.ZoneInfoNotFoundError, ValueError):
            pass                                                            # a name chrono-tz cannot parse counts as no zone (:97-102)
    if which in ("hour", "minute", "second"):
        day_ms = ms % 86_400_000
        return {"hour": day_ms // 3_600_000, "minute": day_ms % 3_600_000 // 60_000, "second": day_ms % 60_000 // 1000}[which]
    if zone is None and unit != "date32":
        days = trunc(ms, 86_400_000)                                        # arrow's cast to Date32 divides toward zero (unpinned)
    else:
        days = ms // 86_400_000                                             # ts_ms_to_local_date32 floors (:213-227)
    d = dt.date(1970, 1, 1) + dt.timedelta(days=days)
    return {"year": d.year, "month": d.month, "day": d.day, "dayofweek": d.isoweekday() % 7 + 1, "quarter": (d.month - 1) // 3 + 1}[which]
