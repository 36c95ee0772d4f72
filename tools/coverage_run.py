"""Function-hit coverage of the device engine (run with GYMCITY_B200_LIB pointing at the -DMP_COVERAGE build):
drives the workloads of the GPU parity suite in one process -- a 48-city random-action soak, the 24 dense cities with
whole-map tools, the rare functions called directly, disasters forced, every tool on every terrain, paint env,
Extinguisher, power puzzle -- and prints {"count": N, "hit": [...], "missed": [...]} as one JSON line.  Parity of these
workloads is asserted by the tests they are taken from; this run only records which functions they entered."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gymcity_b200 import lib  # noqa: E402
from gymcity_b200.engine import EngineBatch  # noqa: E402
from gymcity_b200.micropolis_env import MicropolisPaintEnv, MicropolisVecEnv  # noqa: E402
from gymcity_b200.wrappers import Extinguisher  # noqa: E402

quick = "--quick" in sys.argv
rng = np.random.default_rng(0)

# 1. random-action soak with episodes (auto-reset: terrain generator, reset tick)
env = MicropolisVecEnv(48, 16, seed=1, max_step=120, empty_start=False)
env.reset()
for t in range(60 if quick else 250):
    env.step(torch.from_numpy(rng.integers(0, 20 * 256, 48)))
env.close()

# 2. dense cities, whole-map tools, disasters on
z = np.load(os.path.join(ROOT, "tests", "golden", "cty_maps.npz"))
maps = z["maps"]
n = maps.shape[0]
b = EngineBatch(n)
b.generate(np.arange(n) + 1005, np.arange(n) + 5)
for i in range(n):
    b.set(i, "MPF_MAP", maps[i])
for t in range(40 if quick else 160):
    tools = rng.integers(0, 20, n)
    tools[tools == 5] = 7
    b.toolDown(tools, rng.integers(0, 120, n), rng.integers(0, 100, n))
    b.setFunds(2000000)
    b.controlSimTick()
    if t % 20 == 10:                                       # force every disaster and sprite generator now and then
        for fn, a, c in [("MAKE_FLOOD", 0, 0), ("SET_FIRE", 0, 0), ("MAKE_TORNADO", 0, 0), ("MAKE_MONSTER", 0, 0),
                         ("MAKE_EARTHQUAKE", 0, 0), ("MAKE_MELTDOWN", 0, 0), ("GENERATE_SHIP", 0, 0),
                         ("GENERATE_PLANE", 60, 50), ("GENERATE_COPTER", 60, 50), ("GENERATE_TRAIN", 60, 50),
                         ("MAKE_EXPLOSION", 60, 50)]:
            b.call(fn, a, c)
for fn in ["DEC_TRAFFIC", "DEC_ROG", "FIRE_ANALYSIS", "CRIME_SCAN", "POP_DENSITY_SCAN", "PTLV_SCAN", "SET_VALVES", "TAKE10",
           "TAKE120", "COLLECT_TAX", "CITY_EVAL", "SEND_MESSAGES", "ANIMATE", "MOVE_OBJECTS", "DO_DISASTERS", "POWER_SCAN",
           "MAP_SCAN", "SIMULATE", "DO_SIM_INIT"]:
    b.call(fn)
# the speed divider (simSpeed 1 / 2: sim_loop_head skips frames; the gym path always runs at 3)
for i, sp in ((0, 1), (1, 2)):
    sc = b.scalars(i)
    sc.simSpeed = sp
    b.set_scalars(i, sc)
b.simFrames(40)
# every tool on every terrain, world edges included
for k in range(1000 if quick else 4000):
    tools = rng.integers(0, 20, n)
    b.toolDown(tools, rng.integers(-1, 121, n), rng.integers(-1, 101, n))
b.close()

# 3. paint env, Extinguisher (all four kinds), power puzzle, random static start
p = MicropolisPaintEnv(6, 12, seed=2, max_step=5, empty_start=False)
p.reset()
for t in range(7):
    p.step(torch.from_numpy(rng.integers(0, 20, (6, 12, 12))))
p.close()
for kind in ("age", "spatial", "random", "Monster"):
    x = Extinguisher(MicropolisVecEnv(4, 16, seed=3, max_step=50, empty_start=False), kind, 0.5, seed=1)
    x.reset()
    for t in range(6):
        x.step(torch.from_numpy(rng.integers(0, 20 * 256, 4)))
    x.close()
pz = MicropolisVecEnv(8, 32, seed=4, max_step=20, empty_start=False, power_puzzle=True)
pz.reset()
for t in range(10):
    pz.step(torch.from_numpy(rng.integers(0, 32 * 32, 8)))
pz.close()
rs = MicropolisVecEnv(4, 16, seed=5, max_step=20, empty_start=False, random_builds=True)
rs.reset()
rs.step(torch.from_numpy(rng.integers(0, 20 * 256, 4)))
rs.close()
pe = MicropolisVecEnv(4, 16, seed=6, max_step=20, empty_start=False, poet=True)
pe.reset()
pe.step(torch.from_numpy(rng.integers(0, 20 * 256, 4)))
pe.close()

L = lib()
L.mpb_coverage.restype, L.mpb_coverage.argtypes = C.c_int, [C.c_void_p, C.c_int]
L.mpb_coverage_name.restype, L.mpb_coverage_name.argtypes = C.c_char_p, [C.c_int]
bits = (C.c_ulonglong * 8)()
cnt = L.mpb_coverage(bits, 8)
if cnt < 0:
    print(json.dumps({"count": -1, "error": "not a coverage build"}))
    sys.exit(0)
hit, missed = [], []
for i in range(cnt):
    (hit if (bits[i >> 6] >> (i & 63)) & 1 else missed).append(L.mpb_coverage_name(i).decode())
print(json.dumps({"count": cnt, "hit": hit, "missed": missed}))
