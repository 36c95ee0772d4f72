"""Device time of the phases of aged headline steps (MPB_TIMING events inside the library) next to the whole step."""
import ctypes as C
import os
import sys
import numpy as np
import torch
os.environ["MPB_TIMING"] = "1"
sys.path.insert(0, '.')
from gymcity_b200.micropolis_env import MicropolisEnv
n = 32768
dev = torch.device("cuda", 0)
env = MicropolisEnv(num_envs=n); env.seed(0); env.setMapSize(16, max_step=10**9, empty_start=False); env.reset()
m = env.micro
m._L.mpb_step_phase_times.argtypes = [C.c_void_p, C.c_void_p]
g = torch.Generator(device=dev); g.manual_seed(1)
acts = torch.randint(0, 20 * 256, (64, n), device=dev, dtype=torch.int64, generator=g)
stream = torch.cuda.current_stream(dev)
for i in range(45):
    m._ck(m._L.mpb_env_step(m._h, acts[i % 64].data_ptr(), 0, m.engine._stream()))
rows, tot = [], []
for i in range(20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    m._ck(m._L.mpb_env_step(m._h, acts[i % 64].data_ptr(), 0, m.engine._stream()))
    e1.record(stream)
    torch.cuda.synchronize()
    t = (C.c_float * 5)()
    m._ck(m._L.mpb_step_phase_times(m._h, t))
    rows.append(list(t)); tot.append(e0.elapsed_time(e1))
r = np.array(rows).mean(0)
print("schedule %.3f  action %.3f  frames %.3f  post %.3f  auto-reset %.3f  | sum %.3f  step %.3f ms" % (*r, r.sum(), np.mean(tot)))
