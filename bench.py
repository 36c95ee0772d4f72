#!/usr/bin/env python
"""bench.py -- Micropolis env-ticks/s (BASELINE.json metric) on N B200s of one node.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
  python bench.py --impl reference ...                      (the reference C++ engine on the host cores)

Workload (config.workload): MicropolisEnv 16x16 build window at (16,8) in the 120x100 world, terrain
path (simulation live), uniform random actions, E envs per GPU (weak scaling; E = 32768 = BASELINE
config 5's per-GPU share at 8 GPUs).  One step = one env-tick for every env of the batch = one fused
kernel launch: action through the TileMap clear-then-build logic, setFunds, 2 x 100 simFrames +
animateTiles, observation / reward / done.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Micropolis env-ticks/sec"
UNIT = "env-ticks/s"
B_STATE = 57722             # S1, S3-S8 as the reference stores them


def b_tick(W, H):
    """Algorithmic bytes per env-tick: state + shadow read and written once, observation
    written once, 9 bytes of action / reward / done."""
    return 2 * (B_STATE + 26 * W * H) + 32 * W * H * 4 + 9


# BASELINE.json configs as bench workloads.  "mp16" is the headline (config 5's per-GPU share, and config
# 1's setup batched); the others are reported on request (--workload) and are otherwise parity-test cases.
WORKLOADS = {
    "mp16": dict(size=16, envs=32768, xs=16, ys=8, puzzle=False,
                 text="MicropolisEnv 16x16 window @ (16,8), terrain path, uniform random actions"),
    "mp32puzzle": dict(size=32, envs=8192, xs=16, ys=8, puzzle=True,
                       text="MicropolisEnv 32x32 power puzzle (5 static Residential + 1 static Nuclear, Wire-only "
                            "actions), terrain path"),
    "mp120": dict(size=(120, 100), envs=2048, xs=0, ys=0, puzzle=False,
                  text="Full 120x100 world window @ (0,0), terrain path, uniform random actions over all 20 tools"),
}


def workload(args):
    w = WORKLOADS[args.workload]
    W, H = (w["size"], w["size"]) if isinstance(w["size"], int) else w["size"]
    return {"workload": "%s, %d envs/GPU, env-tick = action + setFunds + 200 simFrames + animateTiles + obs/reward"
                        % (w["text"], args.envs),
            "name": args.workload, "envs_per_gpu": args.envs, "map": "%dx%d window of the 120x100 world" % (W, H),
            "sim_frames_per_tick": 200,
            "rollout_buffer": ("obs[%d, E, 32, W, H] written in place (mpb_env_set_obs_buffer)" % (args.rollout_steps + 1))
            if getattr(args, "rollout_steps", 0) > 0 else None,
            "l2": "state working set %.1f GB per GPU >> 126 MB L2 (no flush needed)" % (args.envs * 95e3 / 1e9)}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the GPU is under load (one long-running
    `nvidia-smi -lms 100` child, stopped by PID)."""

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.samples, self.reasons, self.max_mhz, self.proc = gpu, [], set(), None, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                f = [t.strip() for t in line.strip().split(",")]
                if len(f) < 6:
                    continue
                try:
                    self.samples.append(float(f[0]))
                    self.max_mhz = float(f[1])
                except ValueError:
                    continue
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "window": "warm-up + timed region + e2e region"}


def cpu_reference(seconds, name="mp16"):
    from oracle import cpu_bench
    w = WORKLOADS[name]
    W, H = (w["size"], w["size"]) if isinstance(w["size"], int) else w["size"]
    rate, procs, steps, wall = cpu_bench.run(seconds=seconds, win=W, win_h=H, xs=w["xs"], ys=w["ys"])
    return {"value": rate, "unit": UNIT, "cores": procs, "kind": "reference",
            "sample": "%d env-ticks over %d processes (one reference engine each, pinned one per core), %.1f s "
                      "busy each, %dx%d window, terrain path, uniform random tool + Water/Land/Clear per tick, "
                      "C++ loop (oracle/ref_shim.cpp mpref_bench_rollout)" % (steps, procs, seconds, W, H)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per = max(2.0, min(20.0, 4.0 * max(1, args.steps) / 10))
    t0 = time.perf_counter()
    cb = cpu_reference(per, args.workload)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int16/uint16 (+fp32 valves)", "data": "synthetic",
            "config": workload(args), "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from gymcity_b200.micropolis_env import MicropolisEnv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    E, K, W = args.envs, args.steps, max(3, args.warmup)
    dev = torch.device("cuda", local)

    # the reference-engine baseline runs first on rank 0 (host cores only), before the GPU is busy
    cb = None
    if rank == 0 and not args.no_cpu_baseline:
        cb = cpu_reference(args.cpu_seconds, args.workload)

    wl = WORKLOADS[args.workload]
    env = MicropolisEnv(num_envs=E, device=local, env_offset=rank * E)
    env.seed(args.seed)
    env.setMapSize(wl["size"], max_step=10 ** 9, empty_start=False, power_puzzle=wl["puzzle"], MAP_XS=wl["xs"],
                   MAP_YS=wl["ys"])
    env.reset()
    m = env.micro
    nA = env.action_space.n
    g = torch.Generator(device=dev)
    g.manual_seed(args.seed * 1000 + rank)
    acts = torch.randint(0, nA, (W + 2 * K + 2, E), device=dev, dtype=torch.int64, generator=g)
    stream = torch.cuda.current_stream(dev)
    launches_per_step = int(m._L.mpb_launches_per_env_step(m._h))
    sp = m.engine._stream()

    # BASELINE config 5's consumer: an A2C rollout buffer obs[T + 1, E, 32, W, H] (storage.py:13-15, --num-steps 5)
    # filled in place -- the observation kernel writes step t's batch straight into slot (t mod T) + 1
    rollout = None
    if args.rollout_steps > 0:
        rollout = torch.zeros(args.rollout_steps + 1, E, m.obs_channels, env.MAP_X, env.MAP_Y, device=dev)
    nstep = [0]

    def step_dev(i):
        if rollout is not None:
            slot = rollout[nstep[0] % args.rollout_steps + 1]
            m._ck(m._L.mpb_env_set_obs_buffer(m._h, slot.data_ptr()))
            nstep[0] += 1
        m._ck(m._L.mpb_env_step(m._h, acts[i].data_ptr(), 0, sp))

    for i in range(args.age_steps):                 # untimed build-up: cities reach their random-action steady state
        step_dev(i % (W + 2 * K + 2))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(W):
        step_dev(i)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    ev[0].record(stream)
    for i in range(K):
        step_dev(W + i)
        ev[i + 1].record(stream)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    total_ms = ev[0].elapsed_time(ev[K])
    kern_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(K)]
    # end to end through the host-buffer entry point: H2D actions, step, D2H reward + done, every step
    acts_host = acts[W + K: W + 2 * K].cpu().numpy()
    rew_h, done_h = np.zeros(E, np.float32), np.zeros(E, np.uint8)
    env.step_host(acts_host[0].copy(), rew_h, done_h)            # warm
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(K):
        env.step_host(acts_host[i], rew_h, done_h)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    if rank == 0:
        sampler.stop()
        sampler.join(timeout=2)
    t = torch.tensor([total_ms, e2e_s * 1000.0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(t[0].item()), float(t[1].item())
    checksum = float(m.reward.sum().item())

    if rank == 0:
        peaks, peak_src = None, "fallback 6650 GB/s (B200_PROFILING.md)"
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            peak, peak_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
        except Exception:
            peak = 6650.0
        avg_kernel_s = (sum(kern_ms) / K) / 1000.0
        B = b_tick(env.MAP_X, env.MAP_Y)
        achieved = B * E / avg_kernel_s / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tp) and args.workload == "mp16":
            try:
                tj = json.load(open(tp))
                traffic = tj.get("dram_bytes_per_env_tick") * E if tj.get("dram_bytes_per_env_tick") else None
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": world * E * K / (total_ms / 1000.0), "unit": UNIT, "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int16/uint16 (+fp32 valves)", "data": "synthetic",
            "config": workload(args), "gpu_launches": K * launches_per_step,
            "kernels": "per env group: k_mp_env_pre + k_mp_frames<32> launches of %d frames (32 envs per CTA, frame "
                       "lockstep) + k_mp_env_post; + memset + 3 scheduling kernels" % int(os.environ.get("MPB_FRAMES_PER_LAUNCH", 50)),
            "sim_frames_per_s": world * E * K * 200 / (total_ms / 1000.0),
            "e2e": {"value": world * E * K / (e2e_ms / 1000.0), "unit": UNIT, "h2d_bytes_per_step": E * 8,
                    "d2h_bytes_per_step": E * 5,
                    "note": "mpb_env_step_host: int64 actions H2D from pinned memory, fused step, float32 reward + "
                            "uint8 done D2H; the (N,32,W,H) fp32 observation stays in HBM for the learner (DLPack)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_env_tick": B,
                         "note": "the tick is not HBM bound (SURVEY 8d: a few percent expected): ~300 KB of branchy SASS "
                                 "run by one lane per warp is bound by instruction supply -- ncu: GPC-level L1.5 "
                                 "instruction cache (gcc) requests at 55-65 % of peak, issue slots 30 % busy, DRAM 2 % "
                                 "(profiles/r01_k_mp_frames_ncu_full.txt, profiles/r01_notes.md)"},
            "cpu_baseline": cb, "clocks": sampler.result(), "reward_checksum": checksum,
        }
        print(json.dumps(line), flush=True)
    env.close()
    if world > 1:
        dist.destroy_process_group()


def run_gol(args):
    """BASELINE config 2: 1-player Game of Life 16x16 torus, N batched envs, uniform random toggle per tick
    (multi_env.py:170-232).  Reported on request (--workload gol16); a different metric from the headline."""
    import numpy as np
    import torch
    from gymcity_b200 import GoLMultiEnv
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    N, K, W = args.envs, args.steps, max(3, args.warmup)
    w = 16
    env = GoLMultiEnv()
    env.configure(map_width=w, num_proc=N, max_step=10 ** 9, device=local)
    rng = np.random.default_rng(args.seed)
    env.set_state((rng.random((N, w, w)) < 0.2).astype(np.uint8))
    g = torch.Generator(device=dev)
    g.manual_seed(args.seed)
    acts = torch.randint(0, w * w, (64, N), device=dev, dtype=torch.int64, generator=g)
    stream = torch.cuda.current_stream(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    inner = 1
    for i in range(W):
        env.step(acts[i % 64])
    torch.cuda.synchronize(dev)
    ms = []
    for i in range(K):
        flush.zero_()                                                 # L2 flush between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        env.step(acts[i % 64])
        e1.record(stream)
        torch.cuda.synchronize(dev)
        ms.append(e0.elapsed_time(e1))
    total_ms = sum(ms)
    acts_h = acts.cpu().numpy()
    rew_h = np.zeros(N, np.float32)
    env.step_host(acts_h[0], rew_h)
    t0 = time.perf_counter()
    for i in range(K):
        env.step_host(acts_h[i % 64], rew_h)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    try:
        peak, src = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        peak, src = 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"
    B = 32 + 32 + 8 + 4 + 4 + 1 + w * w * 4           # packed rows r/w, int64 action, reward, pop, done, fp32 state plane
    achieved = B * N / (total_ms / K / 1000.0) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_gol_traffic.json")))["dram_bytes_per_env_tick"] * N
    except Exception:
        pass
    # CPU baseline: the numpy restatement of world_pytorch.py (oracle/gol_oracle.py), one core, bounded sample
    from oracle import gol_oracle as go
    n_cpu = min(N, 4096)
    st = (rng.random((n_cpu, w, w)) < 0.2).astype(np.uint8)
    t1 = time.perf_counter()
    reps = 0
    while time.perf_counter() - t1 < 3.0:
        st, _, _ = go.multi_step(st, rng.integers(0, w * w, n_cpu), w * w * 2 / 3, w * w * 2 / 3, 10 ** 9, n_cpu)
        reps += 1
    cpu_rate = reps * n_cpu / (time.perf_counter() - t1)
    line = {"metric": "Game of Life env-ticks/sec", "value": N * K / (total_ms / 1000.0), "unit": UNIT, "n_gpus": 1,
            "steps": K, "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 bit-packed rows", "data": "synthetic",
            "config": {"workload": "1-player Game of Life 16x16 torus, %d batched envs, uniform random toggle per tick" % N,
                       "name": "gol16", "envs_per_gpu": N, "l2": "256 MB buffer written between timed steps (L2 flush)"},
            "gpu_launches": K, "kernels": "k_gol_step (one launch per env-tick of the whole batch)",
            "e2e": {"value": N * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": N * 8, "d2h_bytes_per_step": N * 4,
                    "note": "golb_step_host: int64 actions H2D, step, float32 rewards D2H; the (N,2,16,16) fp32 "
                            "observation stays in HBM (DLPack)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": src, "algorithmic_bytes_per_env_tick": B},
            "cpu_baseline": {"value": cpu_rate, "unit": UNIT, "cores": 1, "kind": "port",
                             "sample": "%d numpy steps of %d boards (oracle/gol_oracle.py multi_step)" % (reps, n_cpu)}}
    print(json.dumps(line), flush=True)
    env.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mp16", choices=sorted(WORKLOADS) + ["gol16"])
    ap.add_argument("--envs", type=int, default=None, help="envs per GPU (default: the workload's)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=8.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--rollout-steps", type=int, default=0,
                    help="> 0: observations go straight into an A2C rollout buffer obs[T+1, E, 32, W, H] (config 5)")
    ap.add_argument("--age-steps", type=int, default=40,
                    help="untimed env-ticks before the warm-up so the random-action cities are in steady state")
    args = ap.parse_args()
    if args.workload == "gol16":
        args.envs = args.envs or 4096
        return run_gol(args)
    if args.envs is None:
        args.envs = WORKLOADS[args.workload]["envs"]
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
