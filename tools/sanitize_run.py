"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): the frame kernel at 32 and 8 cities per
CTA (MPB_FEPC), the one-warp kernels (pre / post / reset with the 32-lane terrain generator / auto-reset lists),
and the Game-of-Life kernels.  Run as
    compute-sanitizer --tool racecheck python tools/sanitize_run.py
"""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
fepc = os.environ.get("MPB_FEPC", "32")
from gymcity_b200.micropolis_env import MicropolisVecEnv
from gymcity_b200 import GoLMultiEnv, GameOfLifeEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 70
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
env = MicropolisVecEnv(n, 16, seed=3, max_step=3, empty_start=False)
env.reset()
rng = np.random.default_rng(0)
for t in range(steps):
    obs, rew, done, _ = env.step(torch.from_numpy(rng.integers(0, 20 * 256, n)))
print("micropolis ok: fepc", fepc, "reward sum", float(rew.sum()), "resets", int(env.micro.counters()[:, 16].sum()) - n)
env.close()
g = GoLMultiEnv(); g.configure(map_width=16, num_proc=300, max_step=50); g.reset()
for t in range(3):
    g.step(torch.from_numpy(rng.integers(0, 256, 300)))
g.step_n(torch.from_numpy(rng.integers(0, 256, (5, 300))).cuda())
g2 = GoLMultiEnv(); g2.configure(map_width=12, num_proc=50, max_step=50); g2.reset()
g2.step(torch.from_numpy(rng.integers(0, 144, 50)))
c = GameOfLifeEnv(); c.configure(map_width=16, num_proc=40, max_step=50); c.reset()
c.step(torch.from_numpy(rng.integers(0, 256, 40)))
torch.cuda.synchronize()
print("gol ok")
