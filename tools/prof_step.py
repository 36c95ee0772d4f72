"""One aged env-step inside cudaProfilerStart/Stop, for
  ncu --profile-from-start off --section SourceCounters --import-source on -k regex:k_mp_ ...
usage: prof_step.py [n_envs] [age_steps] [profiled_steps]"""
import sys
import torch
sys.path.insert(0, '.')
from gymcity_b200.micropolis_env import MicropolisEnv

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    age = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    env = MicropolisEnv(num_envs=n); env.seed(0); env.setMapSize(16, max_step=10**9, empty_start=False); env.reset()
    g = torch.Generator(device='cuda'); g.manual_seed(0)
    for t in range(age):
        env.step(torch.randint(0, 20 * 256, (n,), device='cuda', generator=g))
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for t in range(steps):
        env.step(torch.randint(0, 20 * 256, (n,), device='cuda', generator=g))
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("done")
