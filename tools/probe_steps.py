"""Per-step wall times of the bench plan on HBM-resident files (diagnostic: looks for drift across steps) + raw pinned H2D bandwidth."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyarrow as pa, torch
import bench
from auron_b200 import proto as P, runtime
os.environ.setdefault("AURON_GPU_CHUNK_ROWS", "320000000")
files = bench.gen_dataset(os.path.join("/tmp", "auron_b200_bench"), bench.SF100_ROWS)
paths = [f for f, _ in files]; sizes = [os.path.getsize(p) for p in paths]
hp = [f"hbm://{os.path.basename(p)}" for p in paths]
for p, h in zip(paths, hp):
    runtime.put_device_file(h, open(p, "rb").read())
plan = bench.build_plan(P, hp, sizes)
prof = os.environ.get("AURON_PROFILE")
ts = []
for i in range(16):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with runtime.Task(plan) as task:
        out = pa.Table.from_batches(list(task), schema=task.schema)
        if prof: task.metrics()
    torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print("profile" if prof else "noprofile", "step ms:", " ".join(f"{x:.1f}" for x in ts))
x = torch.empty(1_360_000_000, dtype=torch.uint8).pin_memory(); y = torch.empty_like(x, device="cuda")
for _ in range(2): y.copy_(x, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter(); y.copy_(x, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"pinned H2D 1.36 GB: {dt*1e3:.1f} ms = {1.36/dt:.1f} GB/s")
