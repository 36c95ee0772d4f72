"""Isolate what makes the episodic workload's main part slow: (V1) auto-reset on, no episode ever ends; (V2) episodic
steady state; (V3) the V2 state with max_step raised so that no reset happens during the timed steps."""
import ctypes as C
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from gymcity_b200.micropolis_env import MicropolisVecEnv
n = 32768
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
acts = torch.randint(0, 20 * 256, (64, n), device=dev, dtype=torch.int64, generator=g)
stream = torch.cuda.current_stream(dev)

def run(m, k0, k):
    ev = []
    for i in range(k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        m._ck(m._L.mpb_env_step(m._h, acts[(k0 + i) % 64].data_ptr(), 0, m.engine._stream()))
        e1.record(stream)
        ev.append((e0, e1))
    torch.cuda.synchronize()
    return np.array([a.elapsed_time(b) for a, b in ev])

env = MicropolisVecEnv(n, 16, seed=0, max_step=10**9, empty_start=False); env.reset()
ms = run(env.micro, 0, 70)
print("V1 auto-reset on, nothing ends: %.2f ms" % ms[45:].mean(), flush=True)
env.close()

env = MicropolisVecEnv(n, 16, seed=0, max_step=200, empty_start=False); env.reset()
env.set_num_steps((np.arange(n, dtype=np.int64) * 200 // n).astype(np.int32))
m = env.micro
run(m, 0, 210)
ms = run(m, 210, 20)
print("V2 episodic steady state: %.2f ms" % ms.mean(), flush=True)
c = m.counters()
print("   busy-cost proxy: envs by num_step quartile; phase classes:", np.bincount((c[:, 4] > 100).astype(int)))
m._ck(m._L.mpb_env_set_max_step(m._h, 10**9))
ms = run(m, 230, 20)
print("V3 same state, no reset in the timed steps: %.2f ms (first 5: %s)" % (ms.mean(), np.round(ms[:5], 1)), flush=True)
m._L.mpb_env_set_auto_reset.argtypes = [C.c_void_p, C.c_int, C.c_int]
m._ck(m._L.mpb_env_set_auto_reset(m._h, 0, 1))
ms = run(m, 250, 20)
print("V4 same, auto-reset launches off: %.2f ms" % ms.mean(), flush=True)
