"""Soak of the device-side auto-reset + terrain prefetch against reference workers: N envs x STEPS steps with short
episodes (max_step varies the reset rhythm), every reward / done, the observation every 10 steps and the engine state at
the end, one reference env per worker reset with sharding.mix seeds exactly when its episode ends.
  python tools/soak_episodes.py [envs] [steps] [max_step]"""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, '.')
from gymcity_b200.micropolis_env import MicropolisVecEnv
from gymcity_b200.sharding import mix
from oracle.ref_engine import RefEngine, diff_snapshots
from oracle.ref_gym import RefEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
max_step = int(sys.argv[3]) if len(sys.argv) > 3 else 23
seed, off, size = 77, 1000, 16
env = MicropolisVecEnv(n, size, seed=seed, max_step=max_step, empty_start=False, env_offset=off)
refs = [RefEnv(RefEngine(), size, max_step=max_step, empty_start=False) for _ in range(n)]
episode = np.zeros(n, int)


def ref_reset(i):
    refs[i].reset_begin(mix(seed, off + i, episode[i], 0), mix(seed, off + i, episode[i], 1))
    episode[i] += 1
    return refs[i].reset_finish().astype(np.float32)


obs = env.reset().cpu().numpy()
for i in range(n):
    assert np.array_equal(ref_reset(i), obs[i]), i
env.set_num_steps((np.arange(n) * max_step // n).astype(np.int32))           # staggered episodes
for i, r in enumerate(refs):
    r.num_step = int(i * max_step // n)
rng = np.random.default_rng(1)
t0 = time.time()
resets = 0
for t in range(steps):
    acts = rng.integers(0, 20 * size * size, n).astype(np.int64)
    obs, rew, done, _ = env.step(torch.from_numpy(acts))
    rew, done = rew.cpu().numpy().reshape(-1), done.cpu().numpy()
    full = t % 10 == 9
    if full:
        obs = obs.cpu().numpy()
    for i, r in enumerate(refs):
        o, rr, d, _ = r.step(int(acts[i]), False)
        assert rr == rew[i] and d == bool(done[i]), (t, i, rr, rew[i], d, done[i])
        if d:
            o = ref_reset(i)
            resets += 1
        if full:
            assert np.array_equal(np.asarray(o, np.float32), obs[i]), (t, i)
for i, r in enumerate(refs):
    assert not diff_snapshots(r.micro.engine.snapshot(), env.micro.engine.snapshot(i)), i
on, served, inline = env.micro.prefetch_stats()
print("soak ok: %d envs x %d steps, max_step %d: %d resets (%d from prefetched terrain, %d generated in line), %.0f s"
      % (n, steps, max_step, resets, served, inline, time.time() - t0))
