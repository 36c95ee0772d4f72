"""Does a batch whose envs sit in two phase classes (half of them one env-step = 200 frames = 8 phases ahead) step
slower than a single-class batch, without any reset in the picture?"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from gymcity_b200.micropolis_env import MicropolisEnv
n = 32768
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
acts = torch.randint(0, 20 * 256, (64, n), device=dev, dtype=torch.int64, generator=g)
stream = torch.cuda.current_stream(dev)
for name, mask in (("single class", None), ("even envs one step ahead", (np.arange(n) % 2 == 0)),
                   ("random half one step ahead", np.random.default_rng(0).random(n) < 0.5),
                   ("5 % one step ahead", np.random.default_rng(0).random(n) < 0.05)):
    env = MicropolisEnv(num_envs=n); env.seed(0); env.setMapSize(16, max_step=10**9, empty_start=False); env.reset()
    m = env.micro
    if mask is not None:
        a = np.ascontiguousarray(acts[63].cpu().numpy()); mk = np.ascontiguousarray(mask.astype(np.uint8))
        m._ck(m._L.mpb_env_step_masked(m._h, a.ctypes.data, 0, mk.ctypes.data, m.engine._stream()))
    ev = []
    for i in range(70):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        m._ck(m._L.mpb_env_step(m._h, acts[i % 64].data_ptr(), 0, m.engine._stream()))
        e1.record(stream)
        ev.append((e0, e1))
    torch.cuda.synchronize()
    ms = np.array([a.elapsed_time(b) for a, b in ev])
    print("%-30s ms/step at steps 45-70: mean %.2f (even steps %.2f, odd steps %.2f)" % (name, ms[45:].mean(), ms[46::2].mean(), ms[45::2].mean()), flush=True)
    env.close()
