"""Worker for tests/test_sharding.py: run under torch.distributed.run with the gloo backend."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gymcity_b200 import sharding  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n_total = 37
    lo, hi = sharding.shard_range(n_total, world, rank)
    seeds = sharding.episode_seeds(5, range(lo, hi), 3, 1)
    # gather every rank's (range, seeds) on rank 0
    objs = [None] * world
    dist.all_gather_object(objs, (lo, hi, seeds.tolist()))
    # the optional per-step statistics collective
    reward = torch.arange(lo, hi, dtype=torch.float32).reshape(-1, 1)
    done = torch.zeros(hi - lo, dtype=torch.bool)
    done[::3] = True
    stats = sharding.all_reduce_stats(sharding.step_stats(reward, done))
    t = sharding.max_over_ranks([10.0 + rank, 3.0 - rank])
    if rank == 0:
        print("RESULT " + json.dumps({"shards": objs, "stats": stats.tolist(), "tmax": t.tolist(), "world": world}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
