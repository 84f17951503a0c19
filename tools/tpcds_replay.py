#!/usr/bin/env python
"""Plan replay for BASELINE configs[4] (TPC-DS end to end): the physical plans Spark + AuronSparkSessionExtension would hand to the
native engine for the TPC-DS queries the operator set covers, run as ONE native task each over synthetic Parquet tables, and checked
with the rule of the reference's integration harness (dev/auron-it .../QueryResultComparator.scala:64-140): equal row counts, then
row by row in result order -- doubles within 1e-6, everything else by its string form.

No Spark / JVM exists in this image, so the plans are written here by hand in the shape Spark plans them (broadcast hash joins on
the dimension tables, partial + final hash aggregate, TakeOrderedAndProject as sort-with-limit) and the expected results come from
pandas over the very same tables.  Queries: q3, q42, q52, q55 (date_dim x store_sales x item star joins), q7 (four dimensions, AVG of
an integer and of three decimals), q43 (CASE WHEN sums per weekday), q67 (the ranking core: RANK() OVER (PARTITION BY ... ORDER BY sum DESC)
with a rank filter), q96 (three dimensions, global COUNT).

    python tools/tpcds_replay.py [--rows 2000000] [--dir /tmp/auron_tpcds] [--queries q3,q7]

prints one line per query: rows, native milliseconds, pass / fail."""
import argparse
import decimal
import os
import sys
import time

import numpy as np
import pandas as pd
import pyarrow as pa
import pyarrow.parquet as pq

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from auron_b200 import proto as P  # noqa: E402
from auron_b200 import runtime  # noqa: E402

D = decimal.Decimal
I, L, S, F = pa.int32(), pa.int64(), pa.string(), pa.float64()
DEC = pa.decimal128(7, 2)
DAY_NAMES = ["Sunday", "Monday", "Tuesday", "Wednesday", "Thursday", "Friday", "Saturday"]


def _dec(cents: np.ndarray, mask=None) -> pa.Array:
    """decimal(7,2) from integer cents (stored INT32-backed, as Spark writes them)"""
    return pa.array([None if (mask is not None and m) else D(int(c)).scaleb(-2) for c, m in zip(cents, mask if mask is not None else [False] * len(cents))], type=DEC)


def gen_tables(d: str, n_sales: int, seed: int = 11) -> dict[str, pa.Table]:
    rng = np.random.default_rng(seed)
    n_dates, n_items, n_stores, n_cd, n_promo, n_hd, n_time = 2_200, 3_000, 12, 1_920, 300, 720, 86_400 // 60
    dd = np.arange(n_dates)
    date_dim = pa.table({"d_date_sk": pa.array((2450815 + dd).astype(np.int32)), "d_year": pa.array((1998 + dd // 365).astype(np.int32)),
                         "d_moy": pa.array((1 + (dd % 365) // 31).astype(np.int32)), "d_day_name": pa.array([DAY_NAMES[int(x) % 7] for x in dd])})
    item = pa.table({"i_item_sk": pa.array(np.arange(1, n_items + 1, dtype=np.int32)),
                     "i_item_id": pa.array([f"AAAAAAAA{int(i) // 2:08d}" for i in range(n_items)]),          # two item versions share an id (as in TPC-DS)
                     "i_brand_id": pa.array(rng.integers(1001001, 1001040, n_items).astype(np.int32)),
                     "i_brand": pa.array([f"brand #{int(b):02d}" for b in rng.integers(1, 40, n_items)], mask=rng.random(n_items) < 0.01),
                     "i_category_id": pa.array(rng.integers(1, 11, n_items).astype(np.int32)),
                     "i_category": pa.array([["Books", "Music", "Home", "Sports", "Shoes"][int(c)] for c in rng.integers(0, 5, n_items)]),
                     "i_manufact_id": pa.array(rng.integers(120, 136, n_items).astype(np.int32)),
                     "i_manager_id": pa.array(rng.integers(1, 40, n_items).astype(np.int32))})
    store = pa.table({"s_store_sk": pa.array(np.arange(1, n_stores + 1, dtype=np.int32)), "s_store_id": pa.array([f"S{int(i) // 2:015d}" for i in range(n_stores)]),
                      "s_store_name": pa.array([["ese", "able", "ought", "bar"][int(i) % 4] for i in range(n_stores)]),
                      "s_gmt_offset": _dec(np.array([-500, -600] * (n_stores // 2)))})
    cd = pa.table({"cd_demo_sk": pa.array(np.arange(1, n_cd + 1, dtype=np.int32)), "cd_gender": pa.array([["M", "F"][int(i) % 2] for i in range(n_cd)]),
                   "cd_marital_status": pa.array([["S", "M", "D", "W", "U"][int(i) % 5] for i in range(n_cd)]),
                   "cd_education_status": pa.array([["College", "Primary", "Secondary", "Unknown"][int(i) % 4] for i in range(n_cd)])})
    promo = pa.table({"p_promo_sk": pa.array(np.arange(1, n_promo + 1, dtype=np.int32)), "p_channel_email": pa.array([["N", "Y"][int(i) % 7 == 0] for i in range(n_promo)]),
                      "p_channel_event": pa.array([["N", "Y"][int(i) % 5 == 0] for i in range(n_promo)])})
    hd = pa.table({"hd_demo_sk": pa.array(np.arange(1, n_hd + 1, dtype=np.int32)), "hd_dep_count": pa.array((np.arange(n_hd) % 10).astype(np.int32))})
    tt = np.arange(n_time)
    time_dim = pa.table({"t_time_sk": pa.array((tt * 60).astype(np.int32)), "t_hour": pa.array((tt // 60).astype(np.int32)), "t_minute": pa.array((tt % 60).astype(np.int32))})

    def fk(lo, hi, null):
        return pa.array(rng.integers(lo, hi, n_sales).astype(np.int32), mask=rng.random(n_sales) < null)

    store_sales = pa.table({
        "ss_sold_date_sk": fk(2450815 - 20, 2450815 + n_dates + 20, 0.04), "ss_sold_time_sk": pa.array((rng.integers(0, n_time, n_sales) * 60).astype(np.int32), mask=rng.random(n_sales) < 0.02),
        "ss_item_sk": fk(1, n_items + 1, 0.0), "ss_cdemo_sk": fk(1, n_cd + 1, 0.03), "ss_hdemo_sk": fk(1, n_hd + 1, 0.03), "ss_promo_sk": fk(1, n_promo + 1, 0.03),
        "ss_store_sk": fk(1, n_stores + 1, 0.03), "ss_quantity": fk(1, 101, 0.03),
        "ss_list_price": _dec(rng.integers(100, 20_000, n_sales), rng.random(n_sales) < 0.03), "ss_sales_price": _dec(rng.integers(0, 20_000, n_sales), rng.random(n_sales) < 0.03),
        "ss_coupon_amt": _dec(rng.integers(0, 50_000, n_sales), rng.random(n_sales) < 0.03), "ss_ext_sales_price": _dec(rng.integers(0, 999_999, n_sales), rng.random(n_sales) < 0.03)})
    tables = {"date_dim": date_dim, "item": item, "store": store, "customer_demographics": cd, "promotion": promo, "household_demographics": hd,
              "time_dim": time_dim, "store_sales": store_sales}
    os.makedirs(d, exist_ok=True)
    for name, t in tables.items():
        pq.write_table(t, os.path.join(d, name + ".parquet"), compression="snappy", row_group_size=max(1000, t.num_rows // 4), store_decimal_as_integer=True)
    return tables


class Q:
    """plan-building helpers bound to one table directory"""

    def __init__(self, d, tables):
        self.d, self.t = d, tables

    def scan(self, name, cols):
        t, path = self.t[name], os.path.join(self.d, name + ".parquet")
        return P.parquet_scan(t.schema, [(path, os.path.getsize(path))], [t.schema.get_field_index(c) for c in cols]), [t.schema.field(c) for c in cols]

    def dim(self, name, cols, preds, keep):
        """filtered + projected dimension table (the broadcast side)"""
        scan, fields = self.scan(name, cols)
        src = P.filter_(scan, preds) if preds else scan
        types = {f.name: f.type for f in fields}
        return P.projection(src, [P.col(c) for c in keep], keep, [types[c] for c in keep]), [pa.field(c, types[c]) for c in keep]

    @staticmethod
    def bjoin(left, lf, right, rf, lk, rk):
        """fact (left, streamed) x dimension (right, broadcast)"""
        sch = pa.schema(list(lf) + list(rf))
        return P.broadcast_join(sch, left, right, [(P.col(lk), P.col(rk))], "INNER", "RIGHT"), list(sch)

    @staticmethod
    def two_phase(src, keys, aggs):
        """partial + final HashAggregate; aggs: (fn, column, return type, name)"""
        kexpr = [P.col(k) for k in keys]
        partial = P.agg(src, kexpr, keys, [P.agg_expr(fn, [P.col(c)] if c else [], t) for fn, c, t, _ in aggs], [n for *_, n in aggs], ["PARTIAL"] * len(aggs))
        return P.agg(partial, kexpr, keys, [P.agg_expr(fn, [P.lit(None, pa.null())], t) for fn, _, t, _ in aggs], [n for *_, n in aggs], ["FINAL"] * len(aggs))


def _cents(s: pd.Series) -> pd.Series:
    return s.map(lambda v: None if v is None else int(v.scaleb(2)))


def _sum_dec(s: pd.Series):
    """SUM of a decimal column: NULL when no value"""
    v = [x for x in s if x is not None]
    return sum(v, D(0)) if v else None


def _avg_dec(s: pd.Series):
    """AVG(decimal(7,2)) -> decimal(11,6): the sum rescaled by 10^4, floor-divided by the count (avg.rs:165-170)"""
    v = [x for x in s if x is not None]
    return D(int(sum(v, D(0)).scaleb(2)) * 10**4 // len(v)).scaleb(-6) if v else None


def _frames(tables):
    return {k: t.to_pandas() for k, t in tables.items()}


def _order(rows, keys):
    """stable multi-key sort; keys = [(index, asc, nulls_first)]"""
    for idx, asc, nf in reversed(keys):
        nul = [r for r in rows if r[idx] is None]
        val = sorted([r for r in rows if r[idx] is not None], key=lambda r: r[idx], reverse=not asc)
        rows = (nul + val) if nf else (val + nul)
    return rows


# ---------------------------------------------------------------------------------------------------------------- the queries
def star_brand(q: Q, fr, year_pred, item_pred_col, item_pred_val, group, order, with_year_key=True):
    """q3 / q42 / q52 / q55 skeleton: select <group>, sum(ss_ext_sales_price) from date_dim, store_sales, item where ... group by ... order by ... limit 100"""
    dpred = [P.binary("Eq", P.col("d_moy"), P.lit(11, I))] + ([P.binary("Eq", P.col("d_year"), P.lit(year_pred, I))] if year_pred else [])
    dd, ddf = q.dim("date_dim", ["d_date_sk", "d_year", "d_moy"], dpred, ["d_date_sk", "d_year"])
    icols = sorted({"i_item_sk", item_pred_col, *[g for g in group if g.startswith("i_")]})
    it, itf = q.dim("item", icols, [P.binary("Eq", P.col(item_pred_col), P.lit(item_pred_val, I))], ["i_item_sk"] + [g for g in group if g.startswith("i_")])
    ss, ssf = q.scan("store_sales", ["ss_sold_date_sk", "ss_item_sk", "ss_ext_sales_price"])
    ss = P.filter_(ss, [P.is_not_null(P.col("ss_sold_date_sk"))])
    j1, f1 = q.bjoin(ss, ssf, dd, ddf, "ss_sold_date_sk", "d_date_sk")
    j2, f2 = q.bjoin(j1, f1, it, itf, "ss_item_sk", "i_item_sk")
    types = {f.name: f.type for f in f2}
    proj = P.projection(j2, [P.col(c) for c in group + ["ss_ext_sales_price"]], group + ["ss_ext_sales_price"], [types[c] for c in group + ["ss_ext_sales_price"]])
    final = q.two_phase(proj, group, [("SUM", "ss_ext_sales_price", pa.decimal128(17, 2), "sum_agg")])
    out_cols = group + ["sum_agg"]
    plan = P.sort(final, [P.sort_expr(P.col(c), asc, nf) for c, asc, nf in order], limit=100)
    # pandas
    d = fr["date_dim"]
    d = d[(d.d_moy == 11) & ((d.d_year == year_pred) if year_pred else True)]
    i = fr["item"]
    i = i[i[item_pred_col] == item_pred_val]
    m = fr["store_sales"].merge(d, left_on="ss_sold_date_sk", right_on="d_date_sk").merge(i, left_on="ss_item_sk", right_on="i_item_sk")
    rows = [tuple(None if (isinstance(k, float) and np.isnan(k)) else (int(k) if isinstance(k, (np.integer, float)) else k) for k in (key if isinstance(key, tuple) else (key,)))
            + (_sum_dec(g.ss_ext_sales_price),) for key, g in m.groupby(group, dropna=False)]
    exp = _order(rows, [(out_cols.index(c), asc, nf) for c, asc, nf in order])[:100]
    return plan, out_cols, exp


def q3(q, fr):
    return star_brand(q, fr, None, "i_manufact_id", 128, ["d_year", "i_brand", "i_brand_id"], [("d_year", True, True), ("sum_agg", False, False), ("i_brand_id", True, True), ("i_brand", True, True)])


def q42(q, fr):
    return star_brand(q, fr, 2000, "i_manager_id", 1, ["d_year", "i_category_id", "i_category"],
                      [("sum_agg", False, False), ("d_year", True, True), ("i_category_id", True, True), ("i_category", True, True)])


def q52(q, fr):
    return star_brand(q, fr, 2000, "i_manager_id", 1, ["d_year", "i_brand", "i_brand_id"], [("d_year", True, True), ("sum_agg", False, False), ("i_brand_id", True, True), ("i_brand", True, True)])


def q55(q, fr):
    return star_brand(q, fr, 1999, "i_manager_id", 28, ["i_brand_id", "i_brand"], [("sum_agg", False, False), ("i_brand_id", True, True), ("i_brand", True, True)])


def q7(q, fr):
    # select i_item_id, avg(ss_quantity), avg(ss_list_price), avg(ss_coupon_amt), avg(ss_sales_price) from store_sales, customer_demographics, date_dim, item, promotion
    # where ... cd_gender = 'M' and cd_marital_status = 'S' and cd_education_status = 'College' and (p_channel_email = 'N' or p_channel_event = 'N') and d_year = 2000
    # group by i_item_id order by i_item_id limit 100
    cd, cdf = q.dim("customer_demographics", ["cd_demo_sk", "cd_gender", "cd_marital_status", "cd_education_status"],
                    [P.binary("Eq", P.col("cd_gender"), P.lit("M", S)), P.binary("Eq", P.col("cd_marital_status"), P.lit("S", S)),
                     P.binary("Eq", P.col("cd_education_status"), P.lit("College", S))], ["cd_demo_sk"])
    dd, ddf = q.dim("date_dim", ["d_date_sk", "d_year"], [P.binary("Eq", P.col("d_year"), P.lit(2000, I))], ["d_date_sk"])
    it, itf = q.dim("item", ["i_item_sk", "i_item_id"], [], ["i_item_sk", "i_item_id"])
    pr, prf = q.dim("promotion", ["p_promo_sk", "p_channel_email", "p_channel_event"],
                    [P.binary("Or", P.binary("Eq", P.col("p_channel_email"), P.lit("N", S)), P.binary("Eq", P.col("p_channel_event"), P.lit("N", S)))], ["p_promo_sk"])
    ss, ssf = q.scan("store_sales", ["ss_sold_date_sk", "ss_item_sk", "ss_cdemo_sk", "ss_promo_sk", "ss_quantity", "ss_list_price", "ss_sales_price", "ss_coupon_amt"])
    ss = P.filter_(ss, [P.is_not_null(P.col("ss_cdemo_sk")), P.is_not_null(P.col("ss_sold_date_sk")), P.is_not_null(P.col("ss_promo_sk"))])
    j, f = q.bjoin(ss, ssf, cd, cdf, "ss_cdemo_sk", "cd_demo_sk")
    j, f = q.bjoin(j, f, dd, ddf, "ss_sold_date_sk", "d_date_sk")
    j, f = q.bjoin(j, f, it, itf, "ss_item_sk", "i_item_sk")
    j, f = q.bjoin(j, f, pr, prf, "ss_promo_sk", "p_promo_sk")
    cols = ["i_item_id", "ss_quantity", "ss_list_price", "ss_coupon_amt", "ss_sales_price"]
    types = {x.name: x.type for x in f}
    proj = P.projection(j, [P.col(c) for c in cols], cols, [types[c] for c in cols])
    A = pa.decimal128(11, 6)
    final = q.two_phase(proj, ["i_item_id"], [("AVG", "ss_quantity", F, "agg1"), ("AVG", "ss_list_price", A, "agg2"), ("AVG", "ss_coupon_amt", A, "agg3"), ("AVG", "ss_sales_price", A, "agg4")])
    plan = P.sort(final, [P.sort_expr(P.col("i_item_id"), True, True)], limit=100)
    c, d, p = fr["customer_demographics"], fr["date_dim"], fr["promotion"]
    m = (fr["store_sales"].merge(c[(c.cd_gender == "M") & (c.cd_marital_status == "S") & (c.cd_education_status == "College")], left_on="ss_cdemo_sk", right_on="cd_demo_sk")
         .merge(d[d.d_year == 2000], left_on="ss_sold_date_sk", right_on="d_date_sk").merge(fr["item"], left_on="ss_item_sk", right_on="i_item_sk")
         .merge(p[(p.p_channel_email == "N") | (p.p_channel_event == "N")], left_on="ss_promo_sk", right_on="p_promo_sk"))
    rows = []
    for key, g in m.groupby("i_item_id"):
        qv = g.ss_quantity.dropna()
        rows.append((key, float(qv.sum()) / len(qv) if len(qv) else None, _avg_dec(g.ss_list_price), _avg_dec(g.ss_coupon_amt), _avg_dec(g.ss_sales_price)))
    return plan, ["i_item_id", "agg1", "agg2", "agg3", "agg4"], _order(rows, [(0, True, True)])[:100]


def q43(q, fr):
    # select s_store_name, s_store_id, sum(case when d_day_name = 'Sunday' then ss_sales_price else null end) sun_sales, ... (seven weekdays)
    # from date_dim, store_sales, store where d_date_sk = ss_sold_date_sk and s_store_sk = ss_store_sk and s_gmt_offset = -5 and d_year = 2000
    # group by s_store_name, s_store_id order by s_store_name, s_store_id, sun_sales, ... limit 100
    dd, ddf = q.dim("date_dim", ["d_date_sk", "d_year", "d_day_name"], [P.binary("Eq", P.col("d_year"), P.lit(2000, I))], ["d_date_sk", "d_day_name"])
    st, stf = q.dim("store", ["s_store_sk", "s_store_id", "s_store_name", "s_gmt_offset"], [P.binary("Eq", P.col("s_gmt_offset"), P.lit(D("-5.00"), DEC))],
                    ["s_store_sk", "s_store_id", "s_store_name"])
    ss, ssf = q.scan("store_sales", ["ss_sold_date_sk", "ss_store_sk", "ss_sales_price"])
    ss = P.filter_(ss, [P.is_not_null(P.col("ss_sold_date_sk")), P.is_not_null(P.col("ss_store_sk"))])
    j, f = q.bjoin(ss, ssf, dd, ddf, "ss_sold_date_sk", "d_date_sk")
    j, f = q.bjoin(j, f, st, stf, "ss_store_sk", "s_store_sk")
    names = [n[:3].lower() + "_sales" for n in DAY_NAMES]
    cases = [P.case([(P.binary("Eq", P.col("d_day_name"), P.lit(n, S)), P.col("ss_sales_price"))], P.lit(None, DEC)) for n in DAY_NAMES]
    proj = P.projection(j, [P.col("s_store_name"), P.col("s_store_id")] + cases, ["s_store_name", "s_store_id"] + names, [S, S] + [DEC] * 7)
    final = q.two_phase(proj, ["s_store_name", "s_store_id"], [("SUM", n, pa.decimal128(17, 2), n) for n in names])
    cols = ["s_store_name", "s_store_id"] + names
    plan = P.sort(final, [P.sort_expr(P.col(c), True, True) for c in cols], limit=100)
    d, s = fr["date_dim"], fr["store"]
    m = fr["store_sales"].merge(d[d.d_year == 2000], left_on="ss_sold_date_sk", right_on="d_date_sk").merge(s[s.s_gmt_offset == D("-5.00")], left_on="ss_store_sk", right_on="s_store_sk")
    rows = [(k[0], k[1]) + tuple(_sum_dec(g.ss_sales_price[g.d_day_name == n]) for n in DAY_NAMES) for k, g in m.groupby(["s_store_name", "s_store_id"])]
    return plan, cols, _order(rows, [(i, True, True) for i in range(len(cols))])[:100]


def q96(q, fr):
    # select count(*) from store_sales, household_demographics, time_dim, store
    # where ... t_hour = 20 and t_minute >= 30 and hd_dep_count = 7 and s_store_name = 'ese'
    hd, hdf = q.dim("household_demographics", ["hd_demo_sk", "hd_dep_count"], [P.binary("Eq", P.col("hd_dep_count"), P.lit(7, I))], ["hd_demo_sk"])
    td, tdf = q.dim("time_dim", ["t_time_sk", "t_hour", "t_minute"], [P.binary("Eq", P.col("t_hour"), P.lit(20, I)), P.binary("GtEq", P.col("t_minute"), P.lit(30, I))], ["t_time_sk"])
    st, stf = q.dim("store", ["s_store_sk", "s_store_name"], [P.binary("Eq", P.col("s_store_name"), P.lit("ese", S))], ["s_store_sk"])
    ss, ssf = q.scan("store_sales", ["ss_sold_time_sk", "ss_hdemo_sk", "ss_store_sk"])
    ss = P.filter_(ss, [P.is_not_null(P.col("ss_sold_time_sk")), P.is_not_null(P.col("ss_hdemo_sk")), P.is_not_null(P.col("ss_store_sk"))])
    j, f = q.bjoin(ss, ssf, hd, hdf, "ss_hdemo_sk", "hd_demo_sk")
    j, f = q.bjoin(j, f, td, tdf, "ss_sold_time_sk", "t_time_sk")
    j, f = q.bjoin(j, f, st, stf, "ss_store_sk", "s_store_sk")
    partial = P.agg(j, [], [], [P.agg_expr("COUNT", [], L)], ["cnt"], ["PARTIAL"])
    plan = P.agg(partial, [], [], [P.agg_expr("COUNT", [P.lit(None, pa.null())], L)], ["cnt"], ["FINAL"])
    h, t, s = fr["household_demographics"], fr["time_dim"], fr["store"]
    m = (fr["store_sales"].merge(h[h.hd_dep_count == 7], left_on="ss_hdemo_sk", right_on="hd_demo_sk")
         .merge(t[(t.t_hour == 20) & (t.t_minute >= 30)], left_on="ss_sold_time_sk", right_on="t_time_sk").merge(s[s.s_store_name == "ese"], left_on="ss_store_sk", right_on="s_store_sk"))
    return plan, ["cnt"], [(len(m),)]


def q67(q, fr):
    # the ranking core of q67 / q70 (ROLLUP left out): the five best (brand, year) cells of every category by revenue
    # select * from (select i_category, i_brand, d_year, sumsales, rank() over (partition by i_category order by sumsales desc) rk
    #                from (select i_category, i_brand, d_year, sum(ss_ext_sales_price) sumsales from store_sales, date_dim, item
    #                      where ss_sold_date_sk = d_date_sk and ss_item_sk = i_item_sk group by i_category, i_brand, d_year) where sumsales is not null)
    # where rk <= 5 order by i_category, rk, i_brand, d_year limit 100
    dd, ddf = q.dim("date_dim", ["d_date_sk", "d_year"], [], ["d_date_sk", "d_year"])
    it, itf = q.dim("item", ["i_item_sk", "i_category", "i_brand"], [], ["i_item_sk", "i_category", "i_brand"])
    ss, ssf = q.scan("store_sales", ["ss_sold_date_sk", "ss_item_sk", "ss_ext_sales_price"])
    ss = P.filter_(ss, [P.is_not_null(P.col("ss_sold_date_sk"))])
    j, f = q.bjoin(ss, ssf, dd, ddf, "ss_sold_date_sk", "d_date_sk")
    j, f = q.bjoin(j, f, it, itf, "ss_item_sk", "i_item_sk")
    keys = ["i_category", "i_brand", "d_year"]
    types = {x.name: x.type for x in f}
    proj = P.projection(j, [P.col(c) for c in keys + ["ss_ext_sales_price"]], keys + ["ss_ext_sales_price"], [types[c] for c in keys + ["ss_ext_sales_price"]])
    agg = P.filter_(q.two_phase(proj, keys, [("SUM", "ss_ext_sales_price", pa.decimal128(17, 2), "sumsales")]), [P.is_not_null(P.col("sumsales"))])
    order = [P.sort_expr(P.col("i_category"), True, True), P.sort_expr(P.col("sumsales"), False, False)]
    win = P.window(P.sort(agg, order), [P.window_expr("rk", I, "RANK")], [P.col("i_category")], [P.sort_expr(P.col("sumsales"), False, False)])
    top = P.filter_(win, [P.binary("LtEq", P.col("rk"), P.lit(5, I))])
    cols = keys + ["sumsales", "rk"]
    plan = P.sort(top, [P.sort_expr(P.col(c), True, True) for c in ["i_category", "rk", "i_brand", "d_year"]], limit=100)
    m = fr["store_sales"].merge(fr["date_dim"], left_on="ss_sold_date_sk", right_on="d_date_sk").merge(fr["item"], left_on="ss_item_sk", right_on="i_item_sk")
    cells = []
    for key, g in m.groupby(keys, dropna=False):
        sm = _sum_dec(g.ss_ext_sales_price)
        if sm is not None:
            cells.append(tuple(None if (isinstance(k, float) and np.isnan(k)) else (int(k) if isinstance(k, (np.integer, float)) else k) for k in key) + (sm,))
    rows = []
    for cat in {c[0] for c in cells}:
        part = [c for c in cells if c[0] == cat]
        for c in part:
            rk = 1 + sum(1 for o in part if o[3] > c[3])
            if rk <= 5:
                rows.append(c + (rk,))
    return plan, cols, _order(rows, [(0, True, True), (4, True, True), (1, True, True), (2, True, True)])[:100]


QUERIES = {"q3": q3, "q7": q7, "q42": q42, "q43": q43, "q52": q52, "q55": q55, "q67": q67, "q96": q96}


# ---------------------------------------------------------------------------------------------------------------- the comparator
def compare(query_id: str, expected: list[tuple], got: list[tuple], types: list[pa.DataType], tol: float = 1e-6) -> list[str]:
    """QueryResultComparator.scala:64-140: row counts, then position by position; returns the mismatches (empty = pass)"""
    if len(expected) != len(got):
        return [f"{query_id}: {len(got)} rows, expected {len(expected)}"]
    out = []
    has_double = any(pa.types.is_floating(t) for t in types)
    for r, (e, g) in enumerate(zip(expected, got)):
        for c, (ev, gv) in enumerate(zip(e, g)):
            if ev is None or gv is None:
                if (ev is None) != (gv is None):
                    out.append(f"{query_id} row {r} col {c}: expected {ev}, got {gv}")
            elif has_double and pa.types.is_floating(types[c]):
                if abs(float(ev) - float(gv)) > tol:
                    out.append(f"{query_id} row {r} col {c}: expected {ev}, got {gv}")
            elif str(ev) != str(gv):
                out.append(f"{query_id} row {r} col {c}: expected {ev}, got {gv}")
    return out


def run_query(name: str, q: Q, fr) -> tuple[float, int, list[str]]:
    plan, cols, exp = QUERIES[name](q, fr)
    t0 = time.perf_counter()
    with runtime.Task(P.task_definition(plan)) as task:
        out = pa.Table.from_batches(list(task), schema=task.schema)
    ms = 1000 * (time.perf_counter() - t0)
    got = list(zip(*[out.column(i).to_pylist() for i in range(out.num_columns)])) if out.num_rows else []
    return ms, out.num_rows, compare(name, exp, got, [f.type for f in out.schema])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=2_000_000)
    ap.add_argument("--dir", default="/tmp/auron_tpcds")
    ap.add_argument("--queries", default=",".join(QUERIES))
    a = ap.parse_args()
    tables = gen_tables(a.dir, a.rows)
    q, fr = Q(a.dir, tables), _frames(tables)
    bad = 0
    for name in a.queries.split(","):
        run_query(name, q, fr)                       # warm-up (first touch of the files, allocator pools)
        ms, rows, errs = run_query(name, q, fr)
        bad += bool(errs)
        print(f"{name:5s} rows={rows:4d} native={ms:8.1f} ms  {'PASS' if not errs else 'FAIL: ' + '; '.join(errs[:3])}")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
