"""Time of MicropolisEnv.reset() for a batch (terrain path): reset_begin (clearMap + generateSomeCity + doSimInit +
updateMap) and reset_finish (MicropolisControl.simTick + observation), CUDA-synchronised wall clock."""
import sys
import time
import torch
sys.path.insert(0, '.')
from gymcity_b200.micropolis_env import MicropolisEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
env = MicropolisEnv(num_envs=n); env.seed(0); env.setMapSize(16, max_step=10**9, empty_start=False)
m = env.micro
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    m.newMapDerived()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    m._ck(m._L.mpb_env_reset_finish(m._h, None, m.engine._stream()))
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("envs %d: reset_begin %.1f ms, reset_finish %.1f ms" % (n, (t1 - t0) * 1e3, (t2 - t1) * 1e3), flush=True)
m2 = MicropolisEnv(num_envs=4); m2.seed(0); m2.setMapSize(16, max_step=10**9, empty_start=False)
torch.cuda.synchronize(); t0 = time.perf_counter(); m2.micro.newMapDerived(); torch.cuda.synchronize()
print("4 envs: reset_begin latency %.2f ms" % ((time.perf_counter() - t0) * 1e3))
