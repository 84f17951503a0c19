"""N>1 coverage.

CPU (gloo, world_size 2): the host-side logic of the multi-rank path -- id distribution, partition ownership, the
reference arm's rank-0-only rule -- runs without a GPU.
GPU: tests/multi_gpu_exchange.py under torch.distributed.run on 2 GPUs (NCCL all-to-all-v of the repartition step)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import torch.distributed as dist
from auron_b200 import runtime
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
ids = [bytes(range(128)) if rank == 0 else None]          # stands for runtime.nccl_unique_id() (needs a GPU)
dist.broadcast_object_list(ids, src=0)
assert ids[0] == bytes(range(128))
nparts = 200
mine = [p for p in range(nparts) if runtime.owner_of_partition(p, world, nparts) == rank]
counts = [None] * world
dist.all_gather_object(counts, mine)
allp = sorted(p for c in counts for p in c)
assert allp == list(range(nparts)), "ownership must partition [0, N)"
assert all(c == sorted(c) and c == list(range(c[0], c[-1] + 1)) for c in counts), "owners hold contiguous blocks"
assert abs(len(counts[0]) - len(counts[1])) <= 1
dist.barrier()
dist.destroy_process_group()
print("RANK_OK", rank)
"""


def _torchrun(args, env=None, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29533"] + args
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run(cmd, cwd=ROOT, env=e, capture_output=True, text=True, timeout=timeout)


def test_gloo_world2_ownership_and_id_distribution(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % ROOT)
    r = _torchrun([str(script)])
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("RANK_OK") == 2


def test_reference_arm_runs_on_rank0_only(tmp_path):
    # bench.py --impl reference under torchrun: rank 0 prints the line, the other ranks exit 0 without work
    r = _torchrun(["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--rows", "200000", "--data-dir", str(tmp_path)],
                  timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1 and '"impl": "reference"' in lines[0]


@pytest.mark.gpu
def test_nccl_exchange_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    r = _torchrun(["tests/multi_gpu_exchange.py"])
    assert r.returncode == 0 and "EXCHANGE_OK" in r.stdout and "BROADCAST_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
