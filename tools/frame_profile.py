"""Per-env per-frame cycles of one aged env-step (mpb_frame_profile) -> gpurun_out/frame_prof.npz,
plus the barrier-wait analysis for the schedule that step used."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from gymcity_b200.micropolis_env import MicropolisEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
age = int(sys.argv[2]) if len(sys.argv) > 2 else 40
env = MicropolisEnv(num_envs=n); env.seed(0); env.setMapSize(16, max_step=10**9, empty_start=False); env.reset()
g = torch.Generator(device='cuda'); g.manual_seed(0)
for t in range(age):
    env.step(torch.randint(0, 20 * 256, (n,), device='cuda', generator=g))
eng = env.micro.engine
perm, G = eng.read_schedule()
eng.frame_profile(True)
env.step(torch.randint(0, 20 * 256, (n,), device='cuda', generator=g))
prof = eng.frame_profile(True, read=True)[:, :200]
env.step(torch.randint(0, 20 * 256, (n,), device='cuda', generator=g))
prof2 = eng.frame_profile(True, read=True)[:, :200]
nspr = np.array([len(eng.sprites(int(i))) for i in range(0, n, 8)])
np.savez_compressed('gpurun_out/frame_prof.npz', prof=prof, prof2=prof2, perm=perm, groups=G, nspr=nspr)
c = prof.astype(np.float64)
print("mean cycles/frame %.0f; per env-step mean %.3g max %.3g" % (c.mean(), c.sum(1).mean(), c.sum(1).max()))
