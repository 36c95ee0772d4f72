"""Per-env simulation cost distribution on the GPU (load imbalance of one step)."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from gymcity_b200.micropolis_env import MicropolisEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
age = int(sys.argv[2]) if len(sys.argv) > 2 else 40
env = MicropolisEnv(num_envs=n); env.seed(0); env.setMapSize(16, max_step=10**9, empty_start=False); env.reset()
g = torch.Generator(device='cuda'); g.manual_seed(0)
for t in range(age + 3):
    a = torch.randint(0, 20 * 256, (n,), device='cuda', generator=g)
    if t == age: env.micro.engine.read_cycles()
    env.step(a)
c = env.micro.engine.read_cycles().astype(np.float64) / 3
print("per-step cycles per env: mean %.3g p50 %.3g p90 %.3g p99 %.3g max %.3g  max/mean %.1f" % (
    c.mean(), np.percentile(c, 50), np.percentile(c, 90), np.percentile(c, 99), c.max(), c.max() / c.mean()))
ctr = env.micro.counters()
m = np.stack([env.micro.engine.get(i, "MPF_MAP") & 1023 for i in np.argsort(c)[-5:]])
for row, i in zip(m, np.argsort(c)[-5:]):
    print("env", i, "cycles %.3g" % c[i], "active", int((row >= 48).sum()), "flood", int(((row >= 48) & (row < 52)).sum()),
          "rad", int((row == 52).sum()), "fire", int(((row >= 56) & (row < 64)).sum()), "road", int(((row >= 64) & (row < 208)).sum()),
          "bldg", int(((row >= 240) & (row < 860)).sum()), "sprites", len(env.micro.engine.sprites(int(i))))
act = np.array([int(((env.micro.engine.get(i, "MPF_MAP") & 1023) >= 48).sum()) for i in range(0, n, max(1, n // 256))])
print("active tiles (sampled envs): mean %.1f p90 %.1f max %d" % (act.mean(), np.percentile(act, 90), act.max()))
