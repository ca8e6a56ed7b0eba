"""Host logic of the compressed all-reduce (smart_compress/util/pytorch/allreduce.py): the shard
geometry on CPU; the full pipeline on >= 2 GPUs against the CPU oracle (tools/allreduce_check.py under torchrun)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("tiles,world", [(1, 2), (7, 8), (8, 8), (9, 8), (1000, 3), (0, 4), (131072, 8)])
def test_shards_partition_the_tiles(tiles, world):
    from smart_compress.util.pytorch.allreduce import shard_tiles

    sh = shard_tiles(tiles, world)
    assert len(sh) == world and sh[0][0] == 0
    assert sum(c for _, c in sh) == tiles
    for (f0, c0), (f1, _) in zip(sh, sh[1:]):
        assert f1 == f0 + c0
    counts = [c for _, c in sh]
    assert max(counts) - min(counts) <= 1


def test_a_tile_range_is_a_contiguous_slice_of_every_section():
    """What lets a rank read only ITS shard of a peer's stream: planes and extras are both indexed by warp tile at a
    fixed stride (stream SQB3)."""
    from smart_compress.compress.packed import packed_layout

    lay = packed_layout(10 * 8192 + 5, 6, 8)
    assert lay.planes_bytes == lay.n_warp_tiles * 6 * 128
    assert lay.extras_stride_bytes == 1024 * 2 // 8
    assert lay.extras_off == lay.planes_off + lay.planes_bytes
    assert lay.total_capacity_bytes >= lay.extras_off + lay.n_warp_tiles * lay.extras_stride_bytes


@pytest.mark.gpu
def test_compressed_allreduce_against_the_oracle_on_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "allreduce_check.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
