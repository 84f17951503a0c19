"""Shared helpers for the parity tests."""
from __future__ import annotations

import math

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc

from auron_b200 import proto as P
from auron_b200 import runtime


def batches(table: pa.Table, chunk: int | None = None) -> list[pa.RecordBatch]:
    t = table.combine_chunks()
    if t.num_rows == 0:
        return [pa.RecordBatch.from_pydict({n: pa.array([], type=t.schema.field(n).type) for n in t.column_names})]
    return t.to_batches(max_chunksize=chunk or t.num_rows)


def source(table: pa.Table, rid: str):
    """ffi_reader plan node + its input registration"""
    return P.ffi_reader(table.schema, rid)


def run(plan: bytes, inputs: dict[str, pa.Table], chunk: int | None = None) -> pa.Table:
    td = P.task_definition(plan, stage_id=1, partition_id=0, task_id=7)
    return runtime.run_task(td, {k: batches(v, chunk) for k, v in inputs.items()})


def canon(t: pa.Table) -> list[tuple]:
    """rows as tuples sorted with None first (order-insensitive comparison)"""
    cols = [c.to_pylist() for c in t.columns]
    rows = list(zip(*cols)) if cols else []

    def key(r):
        return tuple((0, 0) if v is None else (1, (v if not isinstance(v, float) or not math.isnan(v) else float("inf"))) for v in r)

    return sorted(rows, key=key)


def assert_same_rows(got: pa.Table, exp: pa.Table, float_tol: float | None = None):
    assert got.num_columns == exp.num_columns, (got.schema, exp.schema)
    g, e = canon(got), canon(exp)
    assert len(g) == len(e), f"row count {len(g)} != {len(e)}"
    for rg, re_ in zip(g, e):
        for a, b in zip(rg, re_):
            if float_tol is not None and isinstance(a, float) and isinstance(b, float):
                if math.isnan(a) and math.isnan(b):
                    continue
                assert abs(a - b) <= float_tol * max(1.0, abs(b)), (rg, re_)
            else:
                assert a == b, (rg, re_)


def i32(vals):
    return pa.array(vals, type=pa.int32())


def table_i32(**cols):
    return pa.table({k: i32(v) for k, v in cols.items()})
