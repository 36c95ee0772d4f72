"""What do the slow frames of the off-city map-scan strips contain?  (per-frame profile + tile census)"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from gymcity_b200.micropolis_env import MicropolisEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
age = 40
env = MicropolisEnv(num_envs=n); env.seed(0); env.setMapSize(16, max_step=10**9, empty_start=False); env.reset()
g = torch.Generator(device='cuda'); g.manual_seed(0)
for t in range(age):
    env.step(torch.randint(0, 20 * 256, (n,), device='cuda', generator=g))
eng = env.micro.engine
maps0 = np.stack([eng.get(i, "MPF_MAP") for i in range(n)])
eng.frame_profile(True)
env.step(torch.randint(0, 20 * 256, (n,), device='cuda', generator=g))
prof = eng.frame_profile(True, read=True)[:, :200].astype(np.float64)
# step 41 starts at phase 8 (8200 frames since reset incl. the reset tick)
ph = (np.arange(200) + 8) % 16
rows = []
for p in (4, 5, 6, 7, 8, 1, 2, 3):
    fr = np.where(ph == p)[0]
    c = prof[:, fr].mean(1)                      # mean cycles of this strip's frames, per env
    x0, x1 = (p - 1) * 15, p * 15
    t = (maps0.reshape(n, 120, 100)[:, x0:x1, :] & 1023).reshape(n, -1)
    kinds = {"flood": (t >= 48) & (t < 52), "rad": t == 52, "fire": (t >= 56) & (t < 64), "road": (t >= 64) & (t < 208),
             "wire": (t >= 208) & (t < 224), "rail": (t >= 224) & (t < 240), "bldg": (t >= 240) & (t < 860), "expl": (t >= 860) & (t < 868)}
    cnt = {k: v.sum(1) for k, v in kinds.items()}
    print("phase %d (cols %d-%d): frame cycles p50 %.0f p90 %.0f p99 %.0f" % (p, x0, x1 - 1, *np.percentile(c, [50, 90, 99])))
    for lo, hi in ((0, 50), (50, 90), (90, 97), (97, 100)):
        a, b = np.percentile(c, lo), np.percentile(c, hi)
        sel = (c >= a) & (c <= b)
        print("   envs p%d-p%d (%.0f-%.0f cyc): mean tiles " % (lo, hi, a, b) + " ".join("%s %.1f" % (k, v[sel].mean()) for k, v in cnt.items()))
