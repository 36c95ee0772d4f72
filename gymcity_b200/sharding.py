"""Host-side sharding of the env batch across the GPUs of one node.

Every env is independent (own map, own RNG stream, own shadow map): ranks own contiguous ranges of
GLOBAL env indices and there is no collective on the tick path.  Per-env seeds are derived from the
global index, so a rollout does not depend on how many GPUs it is sharded over.  The only optional
collective is one small reduction of per-step reward / population statistics."""
import numpy as np


def shard_range(n_total, world, rank):
    """Contiguous [lo, hi) of global env indices owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def mix(*vals):
    """splitmix-style hash of integers -> 31-bit seed (deterministic, host side)."""
    h = 0x9e3779b97f4a7c15
    for v in vals:
        h ^= (int(v) + 0x9e3779b97f4a7c15 + ((h << 6) & 0xffffffffffffffff) + (h >> 2)) & 0xffffffffffffffff
        h = (h * 0xbf58476d1ce4e5b9) & 0xffffffffffffffff
        h ^= h >> 31
    return int(h & 0x7fffffff)


def episode_seeds(seed, global_indices, episode, which):
    """which = 0: terrain seed (generateSomeCity), 1: simRandom seed, per global env index."""
    return np.array([mix(seed, int(i), episode, which) for i in global_indices], np.int32)


def step_stats(reward, done, metrics=None):
    """Local per-step statistics vector [sum reward, n done, n envs, (sum of each metric)]."""
    import torch
    parts = [reward.double().sum().reshape(1), done.double().sum().reshape(1),
             torch.tensor([float(reward.numel())], dtype=torch.float64, device=reward.device)]
    if metrics is not None:
        parts.append(metrics.double().sum(0))
    return torch.cat(parts)


def all_reduce_stats(stats):
    """The one optional collective per step: sum of the local statistics over ranks (NCCL on GPUs,
    gloo in the CPU tests).  A no-op without an initialised process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def max_over_ranks(values):
    """Timing rule of the bench: the slowest rank defines the step time."""
    import torch
    import torch.distributed as dist
    t = values if torch.is_tensor(values) else torch.tensor(values, dtype=torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t
