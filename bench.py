#!/usr/bin/env python
"""bench.py -- Micropolis env-ticks/s (BASELINE.json metric) on N B200s of one node.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
  python bench.py --impl reference ...                      (the reference C++ engine on the host cores)

Workload (config.workload): MicropolisEnv 16x16 build window at (16,8) in the 120x100 world, terrain
path (simulation live), uniform random actions, E envs per GPU (weak scaling; E = 32768 = BASELINE
config 5's per-GPU share at 8 GPUs), observations written in place into an A2C rollout buffer
obs[T+1, E, 32, 16, 16].  One step = one env-tick for every env of the batch: action through the TileMap
clear-then-build logic, setFunds, 2 x 100 simFrames + animateTiles, observation / reward / done.

At N = 1 the line also carries `extra_workloads`: the headline with episodes (max_step 200, staggered, device-side
auto-reset inside the timed region), the same through MicropolisVecEnv.step, BASELINE config 4 (120x100, 2048 envs),
config 3 (32x32 power puzzle, 8192 envs) and config 2 (Game of Life 16x16, 4096 and 1 M boards), each with its
roofline fraction and CPU arm.  `step_ms` gives min / median / max of the per-step times: the random-action cities
keep ageing (no reset in the headline), so the workload drifts slowly over the timed steps.
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
    "mp120dense": dict(size=(120, 100), envs=2048, xs=0, ys=0, puzzle=False, dense=True,
                       text="Full 120x100 world window @ (0,0), every env starts from one of the reference's 24 shipped cities "
                            "(round robin, tests/golden/cty_maps.npz), uniform random actions over all 20 tools"),
}


def load_dense_cities(env):
    """BASELINE config 4, dense variant: the maps of the reference's 24 .cty cities into the envs, round robin, and the
    TileMap shadow re-read from them (MicropolisControl.updateMap) -- big power grids, traffic, zone growth from tick 1."""
    import numpy as np
    maps = np.load(os.path.join(ROOT, "tests", "golden", "cty_maps.npz"))["maps"]
    eng = env.micro.engine
    for i in range(env.num_envs):
        eng.set(i, "MPF_MAP", maps[i % maps.shape[0]])
    env.micro.updateMap()


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
    rate, procs, steps, wall = cpu_bench.run(seconds=seconds, win=W, win_h=H, xs=w["xs"], ys=w["ys"], dense=bool(w.get("dense")))
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


def hbm_peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    except Exception:
        return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def ms_stats(ms):
    s = sorted(ms)
    return {"min": s[0], "median": s[len(s) // 2], "max": s[-1], "first": ms[0], "last": ms[-1], "n": len(ms)}


def time_steps(step, K, stream, dev):
    """K calls of step(i) timed with CUDA events on `stream`: (total ms, per-step ms list)."""
    import torch
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    ev[0].record(stream)
    for i in range(K):
        step(i)
        ev[i + 1].record(stream)
    torch.cuda.synchronize(dev)
    return ev[0].elapsed_time(ev[K]), [ev[i].elapsed_time(ev[i + 1]) for i in range(K)]


def divergence_flags(m):
    """Sticky per-env flags that must be understood before a rollout counts as bit-exact: sprite_overflow (an env
    needed more sprite slots than the table has; the reference's list is unbounded) must be 0; tilemap_error
    counts envs in which the reference's Python TileMap would have died with a KeyError (tilemap.py:488-491)."""
    c = m.counters()
    return {"sprite_overflow_envs": int((c[:, 15] != 0).sum()), "tilemap_error_envs": int((c[:, 5] != 0).sum())}


def make_env(name, E, local, rank, seed, max_step=10 ** 9, vec=False):
    from gymcity_b200.micropolis_env import MicropolisEnv, MicropolisVecEnv
    wl = WORKLOADS[name]
    kw = dict(max_step=max_step, empty_start=False, power_puzzle=wl["puzzle"], MAP_XS=wl["xs"], MAP_YS=wl["ys"])
    if vec:
        return MicropolisVecEnv(E, wl["size"], device=local, env_offset=rank * E, seed=seed, **kw)
    env = MicropolisEnv(num_envs=E, device=local, env_offset=rank * E)
    env.seed(seed)
    env.setMapSize(wl["size"], **kw)
    return env


def run_extra_mp(name, args, local, dev, K, age, cpu_seconds):
    """A BASELINE config other than the headline, through the same code: value, per-step times, roofline, CPU arm."""
    import torch
    wl = WORKLOADS[name]
    E = wl["envs"]
    cb = None if args.no_cpu_baseline else cpu_reference(cpu_seconds, name)
    env = make_env(name, E, local, 0, args.seed)
    env.reset()
    if wl.get("dense"):
        load_dense_cities(env)
    m = env.micro
    g = torch.Generator(device=dev)
    g.manual_seed(args.seed * 1000 + 17)
    acts = torch.randint(0, env.action_space.n, (age + 3 + K, E), device=dev, dtype=torch.int64, generator=g)
    sp, stream = m.engine._stream(), torch.cuda.current_stream(dev)
    for i in range(age + 3):
        m._ck(m._L.mpb_env_step(m._h, acts[i].data_ptr(), 0, sp))
    torch.cuda.synchronize(dev)
    total_ms, ms = time_steps(lambda i: m._ck(m._L.mpb_env_step(m._h, acts[age + 3 + i].data_ptr(), 0, sp)), K, stream, dev)
    peak, _ = hbm_peak()
    B = b_tick(env.MAP_X, env.MAP_Y)
    achieved = B * E / (total_ms / K / 1000.0) / 1e9
    out = {"workload": "%s, %d envs" % (wl["text"], E), "value": E * K / (total_ms / 1000.0), "unit": UNIT, "steps": K,
           "age_steps": age, "ms_per_step": total_ms / K, "step_ms": ms_stats(ms),
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "algorithmic_bytes_per_env_tick": B},
           "cpu_baseline": cb, "flags": divergence_flags(m)}
    env.close()
    return out


def run_episode_and_venv(args, local, dev, K):
    """The headline workload the way the reference's training loop runs it: episodes of max_step = 200 env-ticks
    (arguments.py:134), every env at its own point of its episode (num_step staggered uniformly, as the workers of a
    SubprocVecEnv drift apart), finished envs reset inside the step on the device -- terrain generation, the
    reset's simTick and its observation are inside the timed region.  Then the same env through the Python vec-env
    API (MicropolisVecEnv.step: what envs.py:363-375 calls), wall clock with a device sync at both ends."""
    import numpy as np
    import torch
    E, max_step = WORKLOADS["mp16"]["envs"], 200
    env = make_env("mp16", E, local, 0, args.seed, max_step=max_step, vec=True)
    assert env._device_auto_reset
    env.reset()
    env.set_num_steps((np.arange(E, dtype=np.int64) * max_step // E).astype(np.int32))
    m = env.micro
    g = torch.Generator(device=dev)
    g.manual_seed(args.seed * 1000 + 29)
    pool = 64
    acts = torch.randint(0, env.action_space.n, (pool, E), device=dev, dtype=torch.int64, generator=g)
    sp, stream = m.engine._stream(), torch.cuda.current_stream(dev)
    ep0 = m.counters()[:, 16].sum()
    for i in range(max_step + 5):                    # one full episode length: every env has been reset once, ages uniform
        m._ck(m._L.mpb_env_step(m._h, acts[i % pool].data_ptr(), 0, sp))
    torch.cuda.synchronize(dev)
    ep1 = m.counters()[:, 16].sum()
    total_ms, ms = time_steps(lambda i: m._ck(m._L.mpb_env_step(m._h, acts[i % pool].data_ptr(), 0, sp)), K, stream, dev)
    ep2 = m.counters()[:, 16].sum()
    import ctypes as C
    ar_ms, ar_n = C.c_float(), C.c_int()
    m._L.mpb_auto_reset_time.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    m._ck(m._L.mpb_auto_reset_time(m._h, C.byref(ar_ms), C.byref(ar_n)))
    episode = {"workload": "mp16 with episodes: max_step 200, per-env staggered num_step, device-side auto-reset "
                           "(terrain generation + reset tick + reset observation inside the timed region)",
               "value": E * K / (total_ms / 1000.0), "unit": UNIT, "steps": K, "ms_per_step": total_ms / K,
               "step_ms": ms_stats(ms), "resets_in_timed_region": int(ep2 - ep1), "resets_during_ageing": int(ep1 - ep0),
               "auto_reset_part_of_last_step_ms": ar_ms.value, "resets_in_last_step": ar_n.value,
               "flags": divergence_flags(m)}
    a2 = acts.reshape(pool, E, 1)
    for i in range(3):
        env.step(a2[i])
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for i in range(K):
        obs, rew, done, infos = env.step(a2[i % pool])
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    venv = {"workload": "the same env through MicropolisVecEnv.step(actions (N,1) CUDA tensor) -> (obs, reward, done, "
                        "infos), device auto-reset on; wall clock, device sync at both ends",
            "value": E * K / dt, "unit": UNIT, "steps": K, "ms_per_step": dt * 1000.0 / K}
    # the reference's own rhythm: every worker starts together and runs max_step ticks, so all envs are the same age
    # and reset in the same step.  One whole episode, from the reset's first step to the step that resets everyone.
    env.reset()
    for i in range(2):                               # two steps of warm-up belong to the episode (ages 0, 1)
        m._ck(m._L.mpb_env_step(m._h, acts[i % pool].data_ptr(), 0, sp))
    torch.cuda.synchronize(dev)
    ep3 = m.counters()[:, 16].sum()
    n_sync = max_step + 1                            # one episode length (+1): contains the step that resets every env
    total_ms, ms = time_steps(lambda i: m._ck(m._L.mpb_env_step(m._h, acts[(i + 2) % pool].data_ptr(), 0, sp)), n_sync, stream, dev)
    ep4 = m.counters()[:, 16].sum()
    sync = {"workload": "mp16 with episodes the way the reference's workers run them: all envs start together, max_step 200, "
                        "the whole batch is reset on the device inside one step of the timed region (max_step + 1 consecutive steps timed)",
            "value": E * n_sync / (total_ms / 1000.0), "unit": UNIT, "steps": n_sync, "ms_per_step": total_ms / n_sync,
            "step_ms": ms_stats(ms), "resets_in_timed_region": int(ep4 - ep3), "flags": divergence_flags(m)}
    env.close()
    return episode, venv, sync


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

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

    env = make_env(args.workload, E, local, rank, args.seed)
    env.reset()
    if WORKLOADS[args.workload].get("dense"):
        load_dense_cities(env)
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
        if args.allreduce_stats:
            from gymcity_b200 import sharding
            stats[0] = sharding.all_reduce_stats(sharding.step_stats(m.reward, m.done))
    stats = [None]

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
    total_ms, kern_ms = time_steps(lambda i: step_dev(W + i), K, stream, dev)
    if world > 1:
        dist.barrier()
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
    flags = divergence_flags(m)
    flags["bit_exact"] = flags["sprite_overflow_envs"] == 0      # an env that needed more than 48 sprites diverged from the reference
    W_, H_ = env.MAP_X, env.MAP_Y
    env.close()
    del rollout
    torch.cuda.empty_cache()

    line = None
    if rank == 0:
        peak, peak_src = hbm_peak()
        avg_kernel_s = (sum(kern_ms) / K) / 1000.0
        B = b_tick(W_, H_)
        achieved = B * E / avg_kernel_s / 1e9
        traffic = None
        for tp in ("r02_traffic.json", "r01_traffic.json"):
            tp = os.path.join(ROOT, "profiles", tp)
            if os.path.exists(tp) and args.workload == "mp16":
                try:
                    tj = json.load(open(tp))
                    traffic = tj.get("dram_bytes_per_env_tick") * E if tj.get("dram_bytes_per_env_tick") else None
                    break
                except Exception:
                    traffic = None
        line = {
            "metric": METRIC, "value": world * E * K / (total_ms / 1000.0), "unit": UNIT, "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True,
            "scaling": "strong" if args.total_envs > 0 else "weak",
            "vs_baseline": None, "dtype": "int16/uint16 (+fp32 valves)", "data": "synthetic",
            "config": workload(args), "gpu_launches": K * launches_per_step,
            "collective": ("all-reduce of 3 doubles per step (sum reward, n done, n envs)" if args.allreduce_stats
                           else "none on the tick path (NCCL: barrier + max of the timings only)"),
            "step_ms": ms_stats(kern_ms),
            "kernels": "one-class batches of >= 18 944 envs (the headline): k_mp_sched_class + k_mp_sched_keys (+ CUB radix "
                       "sort), k_mp_pre_keys, k_mp_env_pre_t (thread per env, tool order), k_mp_frames<32> x 2 (a simTick of "
                       "100 frames per launch, 32 envs per CTA in frame lockstep; the schedule is rebuilt between them), "
                       "k_mp_env_post; other batches: three env groups on streams with 50-frame launches",
            "sim_frames_per_s": world * E * K * 200 / (total_ms / 1000.0),
            "e2e": {"value": world * E * K / (e2e_ms / 1000.0), "unit": UNIT, "h2d_bytes_per_step": E * 8,
                    "d2h_bytes_per_step": E * 5,
                    "note": "mpb_env_step_host: int64 actions H2D from pinned memory, fused step, float32 reward + "
                            "uint8 done D2H; the (N,32,W,H) fp32 observation stays in HBM for the learner (DLPack)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_env_tick": B,
                         "note": "the tick is not HBM bound (SURVEY 8d: a few percent expected): ~430 KB of branchy SASS "
                                 "run by one lane per warp is bound by instruction supply and by the wait for the slowest "
                                 "of the 32 cities of a CTA (profiles/r02_notes.md)"},
            "cpu_baseline": cb, "clocks": sampler.result(), "reward_checksum": checksum, "flags": flags,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    # BASELINE's other configs and the episode / vec-env variants of the headline, on one GPU only (the scaling runs
    # stay what they were: one line per N, the headline workload)
    if world == 1 and args.workload == "mp16" and not args.no_extra:
        Kx = max(5, min(K, 20))
        extra = {}
        try:
            extra["mp16_episode"], extra["mp16_venv_step"], extra["mp16_episode_sync"] = run_episode_and_venv(args, local, dev, Kx)
            extra["mp120"] = run_extra_mp("mp120", args, local, dev, Kx, 40, min(4.0, args.cpu_seconds))
            extra["mp120dense"] = run_extra_mp("mp120dense", args, local, dev, Kx, 10, min(4.0, args.cpu_seconds))
            extra["mp32puzzle"] = run_extra_mp("mp32puzzle", args, local, dev, Kx, 40, min(4.0, args.cpu_seconds))
            extra["gol16_4096"] = gol_measure(args, 4096, Kx, local)
            extra["gol16_1m"] = gol_measure(args, 1 << 20, Kx, local, cpu=False)
        except Exception as ex:                      # the headline line must not be lost to an extra
            extra["error"] = repr(ex)
        line["extra_workloads"] = extra
    print(json.dumps(line), flush=True)


def gol_measure(args, N, K, local, cpu=True, T=32):
    """BASELINE config 2: 1-player Game of Life 16x16 torus, N batched envs, uniform random toggle per tick
    (multi_env.py:170-232).  Two figures: one launch per tick (golb_step: what an agent that picks every action from
    the last observation gets) and T ticks per launch (golb_step_n with the per-tick observations and rewards
    written out: scripted / random-action rollouts).  L2 is flushed between timed launches."""
    import numpy as np
    import torch
    from gymcity_b200 import GoLMultiEnv
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    w = 16
    env = GoLMultiEnv()
    env.configure(map_width=w, num_proc=N, max_step=10 ** 9, device=local)
    rng = np.random.default_rng(args.seed)
    env.set_state((rng.integers(0, 5, (N, w, w), dtype=np.uint8) == 0).astype(np.uint8))     # 20 % alive
    g = torch.Generator(device=dev)
    g.manual_seed(args.seed)
    acts = torch.randint(0, w * w, (T, N), device=dev, dtype=torch.int64, generator=g)
    stream = torch.cuda.current_stream(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    Tn = T if N <= 65536 else 4                                        # obs_seq is T x N x 2 KB
    obs_seq = torch.empty(Tn, N, 2, w, w, device=dev)
    rew_seq = torch.empty(Tn, N, device=dev)
    for i in range(3):
        env.step(acts[i])
        env.step_n(acts[:Tn], obs_seq, rew_seq)
    torch.cuda.synchronize(dev)
    ms1, msn = [], []
    for i in range(K):
        for which in (0, 1):
            flush.zero_()                                             # L2 flush between timed iterations
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            if which == 0:
                env.step(acts[i % T])
            else:
                env.step_n(acts[:Tn], obs_seq, rew_seq)
            e1.record(stream)
            torch.cuda.synchronize(dev)
            (ms1 if which == 0 else msn).append(e0.elapsed_time(e1))
    acts_h = acts.cpu().numpy()
    rew_h = np.zeros(N, np.float32)
    env.step_host(acts_h[0], rew_h)
    t0 = time.perf_counter()
    for i in range(K):
        env.step_host(acts_h[i % T], rew_h)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    peak, src = hbm_peak()
    B1 = 32 + 32 + 8 + 4 + 4 + 1 + w * w * 4          # packed rows r/w, int64 action, reward, pop, done, fp32 state plane
    Bn = 8 + 4 + 2 * w * w * 4                        # per tick of a T-tick launch: action, reward, both fp32 planes
    a1 = B1 * N / (sum(ms1) / K / 1000.0) / 1e9
    an = Bn * N * Tn / (sum(msn) / K / 1000.0) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_gol_traffic.json")))["dram_bytes_per_env_tick"] * N
    except Exception:
        pass
    cbl = None
    if cpu and not args.no_cpu_baseline:
        # CPU baseline: the reference's own torch pipeline (world_pytorch.World with cuda=False) restated in
        # oracle/gol_oracle.py TorchWorldPort, all of torch's CPU threads, bounded sample
        from oracle import gol_oracle as go
        n_cpu = min(N, 4096)
        tw = go.TorchWorldPort((rng.random((n_cpu, w, w)) < 0.2).astype(np.uint8))
        t1 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t1 < 3.0:
            tw.step(rng.integers(0, w * w, n_cpu), w * w * 2 / 3, w * w * 2 / 3, 10 ** 9)
            reps += 1
        cbl = {"value": reps * n_cpu / (time.perf_counter() - t1), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": "%d steps of %d boards through oracle/gol_oracle.py TorchWorldPort (the reference's torch ops: circular "
                         "pad, Conv2d 3x3, round, GoLU; torch CPU threads = cores)" % (reps, n_cpu)}
    out = {"workload": "1-player Game of Life 16x16 torus, %d batched envs, uniform random toggle per tick" % N,
           "unit": UNIT, "steps": K, "l2": "256 MB buffer written between timed launches (L2 flush)",
           "per_launch_1_tick": {"value": N * K / (sum(ms1) / 1000.0), "ms_per_launch": sum(ms1) / K, "step_ms": ms_stats(ms1),
                                 "roofline": {"bound": "hbm", "achieved": a1, "peak": peak, "unit": "GB/s", "frac": a1 / peak,
                                              "traffic": traffic, "algorithmic_bytes_per_env_tick": B1}},
           "per_launch_T_ticks": {"T": Tn, "value": N * Tn * K / (sum(msn) / 1000.0), "ms_per_launch": sum(msn) / K,
                                  "step_ms": ms_stats(msn),
                                  "roofline": {"bound": "hbm", "achieved": an, "peak": peak, "unit": "GB/s", "frac": an / peak,
                                               "algorithmic_bytes_per_env_tick": Bn},
                                  "note": "golb_step_n: T ticks per launch, every tick's observation (both planes) and reward written out"},
           "e2e": {"value": N * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": N * 8, "d2h_bytes_per_step": N * 4},
           "cpu_baseline": cbl}
    out["value"] = out["per_launch_T_ticks"]["value"]
    env.close()
    del obs_seq, rew_seq, flush
    torch.cuda.empty_cache()
    return out


def run_gol(args):
    """--workload gol16: the Game-of-Life figures as a line of their own (a different metric from the headline)."""
    import torch
    local = int(os.environ.get("LOCAL_RANK", "0"))
    K = args.steps
    r = gol_measure(args, args.envs, K, local)
    one = r["per_launch_1_tick"]
    line = {"metric": "Game of Life env-ticks/sec", "value": r["per_launch_T_ticks"]["value"], "unit": UNIT, "n_gpus": 1,
            "steps": K, "warmup": max(3, args.warmup), "ms_per_step": r["per_launch_T_ticks"]["ms_per_launch"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 bit-packed rows",
            "data": "synthetic", "config": {"workload": r["workload"], "name": "gol16", "envs_per_gpu": args.envs, "l2": r["l2"]},
            "gpu_launches": K, "kernels": "k_gol_step (T ticks per launch)", "e2e": r["e2e"],
            "roofline": r["per_launch_T_ticks"]["roofline"], "per_launch_1_tick": one, "per_launch_T_ticks": r["per_launch_T_ticks"],
            "cpu_baseline": r["cpu_baseline"]}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mp16", choices=sorted(WORKLOADS) + ["gol16"])
    ap.add_argument("--envs", type=int, default=None, help="envs per GPU (default: the workload's)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=8.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--rollout-steps", type=int, default=5,
                    help="> 0: observations go straight into an A2C rollout buffer obs[T+1, E, 32, W, H] (config 5: "
                         "--num-steps 5); 0: into the library's own buffer")
    ap.add_argument("--total-envs", type=int, default=0,
                    help="> 0: STRONG scaling -- this many envs in total, split evenly over the N GPUs (config 5: 262144)")
    ap.add_argument("--allreduce-stats", action="store_true",
                    help="add the one optional collective of the tick path to every timed step: an all-reduce of "
                         "[sum reward, n done, n envs] over the ranks (sharding.all_reduce_stats)")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip extra_workloads (BASELINE configs 2-4 and the episode / vec-env variants; N = 1 only)")
    ap.add_argument("--age-steps", type=int, default=40,
                    help="untimed env-ticks before the warm-up so the random-action cities are in steady state")
    args = ap.parse_args()
    if args.workload == "gol16":
        args.envs = args.envs or 4096
        return run_gol(args)
    if args.envs is None:
        args.envs = WORKLOADS[args.workload]["envs"]
    if args.total_envs > 0:
        args.envs = args.total_envs // max(1, int(os.environ.get("WORLD_SIZE", "1")))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
