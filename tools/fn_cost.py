"""Instruction cost of single engine functions on aged cities: run under
ncu --metrics smsp__inst_executed.sum -k regex:k_mp_call and read the launches in order."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from gymcity_b200.micropolis_env import MicropolisEnv
FNS = ["PTLV_SCAN", "CRIME_SCAN", "POP_DENSITY_SCAN", "FIRE_ANALYSIS", "POWER_SCAN", "DEC_TRAFFIC", "DEC_ROG",
       "ANIMATE", "TAKE10", "SET_VALVES", "CITY_EVAL", "SEND_MESSAGES", "DO_DISASTERS", "MOVE_OBJECTS"]
if __name__ == "__main__":
    n = 8192
    env = MicropolisEnv(num_envs=n); env.seed(0); env.setMapSize(16, max_step=10**9, empty_start=False); env.reset()
    g = torch.Generator(device='cuda'); g.manual_seed(0)
    for t in range(30):
        env.step(torch.randint(0, 20 * 256, (n,), device='cuda', generator=g))
    for fn in FNS:
        env.micro.engine.call(fn)
    env.micro.engine.sync()
    print("order:", " ".join(FNS))
