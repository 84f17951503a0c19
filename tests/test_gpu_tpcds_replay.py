"""BASELINE configs[4] (TPC-DS end to end) as a plan replay: tools/tpcds_replay.py builds the physical plans of q3, q7, q42, q43, q52,
q55 and q96 the way Spark + AuronSparkSessionExtension plan them, the engine runs each as one native task over SNAPPY Parquet tables,
and the rows are checked in result order with the rule of the reference's integration harness (QueryResultComparator.scala:64-140)
against pandas over the same tables."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import tpcds_replay as R  # noqa: E402


@pytest.fixture(scope="module")
def world(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("tpcds"))
    tables = R.gen_tables(d, 300_000)
    return R.Q(d, tables), R._frames(tables)


@pytest.mark.parametrize("name", sorted(R.QUERIES))
def test_plans_are_accepted_by_the_planner(world, name):
    # CPU: the planner decodes every plan (operators, schemas, expressions) on no device
    from auron_b200 import proto as P
    from auron_b200 import runtime
    plan, cols, exp = R.QUERIES[name](*world)
    tree = runtime.explain(P.task_definition(plan))
    assert tree and len(exp) > 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(R.QUERIES))
def test_query_matches_pandas_under_the_reference_comparator(world, name):
    ms, rows, errs = R.run_query(name, *world)
    assert not errs, errs[:5]
    assert rows > 0


def test_comparator_rules():
    import pyarrow as pa
    t = [pa.string(), pa.float64(), pa.int64()]
    assert R.compare("q", [("a", 1.0, 1)], [("a", 1.0 + 5e-7, 1)], t) == []                    # doubles within 1e-6
    assert R.compare("q", [("a", 1.0, 1)], [("a", 1.0 + 5e-6, 1)], t)                          # ... not beyond
    assert R.compare("q", [("a", None, 1)], [("a", 0.0, 1)], t)                                # NULL only equals NULL
    assert R.compare("q", [("a", 1.0, 1)], [("a", 1.0, 1), ("b", 2.0, 2)], t)                  # row counts first
    assert R.compare("q", [("a", 1.0, 1)], [("a", 1.0, 2)], t)                                 # everything else by its string form
