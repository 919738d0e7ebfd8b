"""not gpu: bench.py host logic -- batch sharding, the reference (CPU port) arm, and the N>1
max-over-ranks aggregation over gloo with world_size 2."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_shard_sizes():
    assert bench.shard_sizes(256, 8) == [32] * 8
    assert bench.shard_sizes(10, 4) == [3, 3, 2, 2]
    assert sum(bench.shard_sizes(4097, 8)) == 4097
    assert bench.shard_sizes(3, 4) == [1, 1, 1, 0]


def test_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-images", "32"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["higher_is_better"] is True
    for key in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config"):
        assert key in line


WORKER = r'''
import os, sys, json
sys.path.insert(0, %r)
import torch, torch.distributed as dist
import bench
dist.init_process_group(backend="gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# each rank "measures" a different time; the job time is the max over ranks (bench.py contract)
ms = torch.tensor([10.0 + 5.0 * rank], dtype=torch.float64)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
per_rank = bench.shard_sizes(4096, world)[rank]
tot = torch.tensor([per_rank], dtype=torch.int64)
dist.all_reduce(tot)
if rank == 0:
    print(json.dumps({"ms": float(ms.item()), "total": int(tot.item()), "world": world}))
dist.destroy_process_group()
'''


def test_gloo_world2_max_over_ranks(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line == {"ms": 15.0, "total": 4096, "world": 2}


def test_reference_arm_under_torchrun_only_rank0_prints():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29732", os.path.join(ROOT, "bench.py"),
                          "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--cpu-images", "16"],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1 and json.loads(lines[0])["n_gpus"] == 2
