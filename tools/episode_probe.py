"""Where does the episodic workload (max_step 200, staggered, device auto-reset) spend its step?
(1) ageing curve without resets: ms per step at ages 0..260; (2) the episodic steady state: step time and the
auto-reset part of it (mpb_auto_reset_time)."""
import ctypes as C
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from gymcity_b200.micropolis_env import MicropolisEnv, MicropolisVecEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
acts = torch.randint(0, 20 * 256, (64, n), device=dev, dtype=torch.int64, generator=g)
stream = torch.cuda.current_stream(dev)

def timed(m, i):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    m._ck(m._L.mpb_env_step(m._h, acts[i % 64].data_ptr(), 0, m.engine._stream()))
    e1.record(stream)
    return e0, e1

env = MicropolisEnv(num_envs=n); env.seed(0); env.setMapSize(16, max_step=10**9, empty_start=False); env.reset()
ev = [timed(env.micro, i) for i in range(260)]
torch.cuda.synchronize()
ms = np.array([a.elapsed_time(b) for a, b in ev])
print("no-reset ageing curve, ms/step by age:", {a: round(float(ms[a:a + 10].mean()), 2) for a in range(0, 260, 20)})
env.close()

env = MicropolisVecEnv(n, 16, seed=0, max_step=200, empty_start=False)
env.reset()
env.set_num_steps((np.arange(n, dtype=np.int64) * 200 // n).astype(np.int32))
m = env.micro
m._L.mpb_auto_reset_time.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
for i in range(210):
    m._ck(m._L.mpb_env_step(m._h, acts[i % 64].data_ptr(), 0, m.engine._stream()))
tot, ar, cnt = [], [], []
for i in range(20):
    e0, e1 = timed(m, i)
    torch.cuda.synchronize()
    t = C.c_float(); k = C.c_int()
    m._L.mpb_auto_reset_time(m._h, C.byref(t), C.byref(k))
    tot.append(e0.elapsed_time(e1)); ar.append(t.value); cnt.append(k.value)
print("episodic steady state: step %.2f ms, of which auto-reset %.2f ms (resets per step %.0f)" % (np.mean(tot), np.mean(ar), np.mean(cnt)))
c = m.counters()
print("num_step distribution: min %d max %d" % (c[:, 4].min(), c[:, 4].max()))
