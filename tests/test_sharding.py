"""CPU: multi-rank host logic (world_size 2, gloo): contiguous shards by global env index, seeds
invariant to the number of ranks, the optional per-step statistics all-reduce, max-over-ranks timing."""
import json
import os
import subprocess
import sys

import numpy as np

from gymcity_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition_the_batch():
    for n in (1, 7, 37, 262144):
        for world in (1, 2, 3, 8):
            edges = [sharding.shard_range(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_run_matches_single_process():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29613",
                          os.path.join(ROOT, "tests", "_dist_worker.py")], capture_output=True, text=True, env=env,
                         timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")][0]
    res = json.loads(line[7:])
    assert res["world"] == 2
    got = []
    for lo, hi, seeds in res["shards"]:
        got += seeds
    assert got == sharding.episode_seeds(5, range(37), 3, 1).tolist()      # same seeds as one process
    rewards = np.arange(37, dtype=np.float64)
    n_done = sum(len(range(lo, hi)[::3]) for lo, hi, _ in res["shards"])
    assert res["stats"] == [rewards.sum(), float(n_done), 37.0]
    assert res["tmax"] == [11.0, 3.0]
