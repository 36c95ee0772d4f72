"""CPU timing of the reference engine -- TEST / BENCH INFRASTRUCTURE ONLY (bench.py's cpu_baseline and
--impl reference legs).  Each worker process owns one reference `Micropolis` instance
(oracle/_ref/libmicropolis_ref.so) and replays env-ticks entirely in C++ (ref_shim.cpp
mpref_bench_rollout: Water/Land/Clear + the tool on a uniformly random window tile, setFunds,
MicropolisControl.simTick = 2 x 100 simFrames + animateTiles), like one SubprocVecEnv worker of the
reference (subproc_vec_env.py:9-56) minus its Python and pipe overhead.

Workers are plain subprocesses (`python -m oracle.cpu_bench --worker ...`) so nothing here depends on
how the caller was launched (torchrun, pytest, stdin)."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(idx, seconds, win, terrain, win_h=None, xs=16, ys=8, dense=0):
    try:
        os.sched_setaffinity(0, {idx % (os.cpu_count() or 1)})
    except Exception:
        pass
    sys.path.insert(0, ROOT)
    from oracle.ref_engine import RefEngine
    r = RefEngine()
    if terrain:
        r.generate(1000 + idx)
    else:
        r.clearMap()
    if dense:                                             # BASELINE config 4, dense variant: one of the reference's 24 cities
        import numpy as np
        maps = np.load(os.path.join(ROOT, "tests", "golden", "cty_maps.npz"))["maps"]
        r.set("MPF_MAP", maps[idx % maps.shape[0]])
    r.seed(idx)
    win_h = win_h or win
    r.bench_rollout(50, win, win_h, xs, ys, act_seed=idx)       # warm-up
    steps, busy = 0, 0.0
    t_end = time.perf_counter() + seconds
    while time.perf_counter() < t_end:
        secs, _ = r.bench_rollout(250, win, win_h, xs, ys, act_seed=idx * 7919 + steps)
        steps += 250
        busy += secs
    print(json.dumps({"steps": steps, "busy": busy}), flush=True)


def run(seconds=8.0, procs=None, win=16, terrain=True, win_h=None, xs=16, ys=8, dense=False):
    """Returns (env_ticks_per_s_total, procs, total_steps, wall_seconds)."""
    procs = procs or (os.cpu_count() or 1)
    t0 = time.perf_counter()
    ps = [subprocess.Popen([sys.executable, "-m", "oracle.cpu_bench", "--worker", str(i), str(seconds), str(win),
                            str(int(terrain)), str(win_h or win), str(xs), str(ys), str(int(dense))], cwd=ROOT, stdout=subprocess.PIPE, text=True)
          for i in range(procs)]
    res = []
    for p in ps:
        out, _ = p.communicate(timeout=seconds * 4 + 120)
        for line in out.splitlines():
            if line.startswith("{"):
                res.append(json.loads(line))
    wall = time.perf_counter() - t0
    rate = sum(r["steps"] / r["busy"] for r in res if r["busy"] > 0)
    return rate, procs, sum(r["steps"] for r in res), wall


if __name__ == "__main__":
    if len(sys.argv) >= 6 and sys.argv[1] == "--worker":
        extra = [int(a) for a in sys.argv[6:10]]
        _worker(int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4]), bool(int(sys.argv[5])), *extra)
    else:
        print(run(seconds=float(sys.argv[1]) if len(sys.argv) > 1 else 3.0))
