"""pytest plugin for the build container (no GPU): lets the gpu-marked tests START, records every TaskDefinition they hand to the engine
and returns mock results, so that each test runs on until its first assertion about data.  Afterwards every recorded plan is decoded by
the engine's planner on no device (auron_b200_explain) and walked against the reference's protobuf contract.  Use after touching the
planner or the plan encoder:

    PYTHONPATH=tools python -m pytest tests -m gpu -p record_test_plans -q --deselect tests/test_gpu_fullsize.py ; python tools/record_test_plans.py

(the pytest run reports failures -- no data comes back -- which is expected; the second command prints what matters)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.environ.get("AURON_RECORDED_PLANS", "/tmp/auron_recorded_plans.bin")


def _rec(td):
    with open(OUT, "ab") as f:
        f.write(len(td).to_bytes(4, "little") + bytes(td))


def _install():
    from unittest import mock

    import pyarrow as pa
    import torch

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from auron_b200 import proto as P
    from auron_b200 import runtime
    open(OUT, "wb").close()

    class FakeTask:
        def __init__(self, task_definition, *a, **k):
            _rec(task_definition)
            self.schema = pa.schema([])

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

        def __iter__(self):
            return iter([])

        def next_batch(self):
            return None

        def metrics(self):
            return []

        def close(self):
            pass

    def fake_run_task(td, *a, **k):
        _rec(td)
        return mock.MagicMock()

    runtime.Task = FakeTask
    runtime.run_task = fake_run_task
    import helpers
    import jni_helpers

    def fake_run(plan, inputs, chunk=None):
        return fake_run_task(P.task_definition(plan, stage_id=1, partition_id=0, task_id=7))

    helpers.run = fake_run
    orig = jni_helpers.MockJvm.__init__

    def jinit(self, task_definition, *a, **k):
        _rec(task_definition)
        return orig(self, task_definition, *a, **k)

    jni_helpers.MockJvm.__init__ = jinit
    torch.cuda.is_available = lambda: True


def _report():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_proto_contract as tc
    from auron_b200 import runtime
    raw = open(OUT, "rb").read()
    seen, i = set(), 0
    n = decoded = conform = 0
    problems = {}
    while i < len(raw):
        ln = int.from_bytes(raw[i:i + 4], "little")
        td = raw[i + 4:i + 4 + ln]
        i += 4 + ln
        if td in seen:
            continue
        seen.add(td)
        n += 1
        try:
            runtime.explain(td)
            decoded += 1
        except Exception as e:   # noqa: BLE001
            problems["planner: " + str(e)[:100]] = problems.get("planner: " + str(e)[:100], 0) + 1
        try:
            tc.walk("TaskDefinition", td, set())
            conform += 1
        except (AssertionError, IndexError, KeyError) as e:
            problems["contract: " + str(e)[:100]] = problems.get("contract: " + str(e)[:100], 0) + 1
    print(f"{n} distinct plans: {decoded} decoded by the planner, {conform} conform to the reference protobuf")
    for k, v in problems.items():
        print(f"  {v} x {k}")


if __name__ == "__main__":
    _report()
else:
    _install()
