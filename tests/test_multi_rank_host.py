"""Host-side logic of the N>1 launch, on CPU with the gloo backend (world_size 2).

The path shards with zero exchange: every rank compresses its own tensors and the only
collectives of bench.py are the barrier and the max over ranks of the device time.  These tests cover that
plumbing: environment parsing, the max-over-ranks reduction, the whole-job aggregate, and the reference arm's
"rank 0 alone works" rule."""
import json
import os
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import bench

    assert bench.dist_env() == (rank, world, rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    local_ms = 10.0 + 5.0 * rank          # rank 1 is the slow one
    slowest = bench.max_over_ranks(local_ms, world, torch.device("cpu"))
    dist.barrier()
    value = bench.whole_job_gbs(world, 13.58 * (1 << 20), slowest)
    with open(os.path.join(out_dir, f"r{rank}.json"), "w") as f:
        json.dump({"slowest": slowest, "value": value}, f)
    dist.destroy_process_group()


def test_max_over_ranks_and_weak_scaling_aggregate(tmp_path):
    world, port = 2, 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = [json.load(open(tmp_path / f"r{r}.json")) for r in range(world)]
    assert got[0] == got[1], "every rank must report the same (max) time"
    assert got[0]["slowest"] == 15.0
    one_rank = 13.58 * (1 << 20) / 15e-3 / 1e9
    assert got[0]["value"] == pytest.approx(2 * one_rank)


def test_single_rank_needs_no_process_group():
    import bench

    assert bench.max_over_ranks(3.5, 1, torch.device("cpu")) == 3.5
    assert bench.whole_job_gbs(1, 1e9, 1000.0) == pytest.approx(1.0)


def test_reference_arm_runs_on_rank_zero_only():
    """Under torchrun the reference arm is timed by rank 0 alone; the other ranks exit 0 without work or output."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
