import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from gymcity_b200.micropolis_env import MicropolisEnv
n = 32768
env = MicropolisEnv(num_envs=n); env.seed(0); env.setMapSize(16, max_step=10**9, empty_start=False); env.reset()
g = torch.Generator(device='cuda'); g.manual_seed(0)
for t in range(46):
    env.step(torch.randint(0, 20 * 256, (n,), device='cuda', generator=g))
    if t >= 42:
        print("step", t, "group end times (ms):", np.round(env.micro.engine.group_times(), 1))
c = env.micro.engine.read_cycles()
