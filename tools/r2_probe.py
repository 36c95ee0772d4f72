"""Round-2 probe: the mp16 bench loop under several scheduler configurations in ONE process (the library reads
its MPB_* knobs at mpb_create), with a checksum of rewards / final state so that every configuration is
verified to produce the same rollout.  Prints one JSON line per configuration.

  python tools/r2_probe.py [--envs 32768] [--age 40] [--steps 20] [--configs name=ENV:VAL,ENV:VAL ...]
"""
import argparse
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DEFAULT = [
    "default=",
    "groups1=MPB_GROUPS:1",
    "f25=MPB_FRAMES_PER_LAUNCH:25",
    "f100=MPB_FRAMES_PER_LAUNCH:100",
]


def run(name, envs, kv, args):
    import torch
    from gymcity_b200.micropolis_env import MicropolisEnv
    keys = []
    for item in kv.split(","):
        if not item:
            continue
        k, v = item.split(":")
        os.environ[k] = v
        keys.append(k)
    try:
        dev = torch.device("cuda", 0)
        env = MicropolisEnv(num_envs=envs, device=0)
        env.seed(0)
        env.setMapSize(args.size, max_step=10 ** 9, empty_start=False, MAP_XS=16, MAP_YS=8)
        env.reset()
        m = env.micro
        g = torch.Generator(device=dev)
        g.manual_seed(1234)
        T = args.age + args.warm + args.steps
        acts = torch.randint(0, env.action_space.n, (T, envs), device=dev, dtype=torch.int64, generator=g)
        sp = m.engine._stream()
        stream = torch.cuda.current_stream(dev)
        rsum = 0.0
        ms = []
        for i in range(T):
            if i >= args.age + args.warm or args.curve:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                m._ck(m._L.mpb_env_step(m._h, acts[i].data_ptr(), 0, sp))
                e1.record(stream)
                ms.append((e0, e1))
            else:
                m._ck(m._L.mpb_env_step(m._h, acts[i].data_ptr(), 0, sp))
            if i % 8 == 7 or i == T - 1:
                rsum += float(m.reward.double().sum().item())
        torch.cuda.synchronize(dev)
        t = np.array([a.elapsed_time(b) for a, b in ms])
        tt = t[-args.steps:]
        crc = 0
        for i in (0, 1, envs // 2, envs - 1):
            crc = zlib.crc32(m.engine.get(i, "MPF_MAP").tobytes(), crc)
        c = m.counters()
        out = {"config": name, "env": kv, "envs": envs, "env_ticks_per_s": envs * len(tt) / (tt.sum() / 1e3),
               "ms_min": float(tt.min()), "ms_med": float(np.median(tt)), "ms_max": float(tt.max()),
               "reward_sum": rsum, "map_crc": crc, "tilemap_error": int(c[:, 5].sum()), "sprite_overflow": int(c[:, 15].sum())}
        if args.curve:
            out["curve_ms"] = [round(float(x), 2) for x in t]
        print(json.dumps(out), flush=True)
        env.close()
    finally:
        for k in keys:
            os.environ.pop(k, None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=32768)
    ap.add_argument("--size", type=int, default=16)
    ap.add_argument("--age", type=int, default=40)
    ap.add_argument("--warm", type=int, default=5)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--curve", action="store_true", help="time every step from the first and print the per-step ms")
    ap.add_argument("--configs", nargs="*", default=DEFAULT)
    args = ap.parse_args()
    for c in args.configs:
        name, _, kv = c.partition("=")
        try:
            run(name, args.envs, kv, args)
        except Exception as ex:                       # keep going: one bad configuration must not hide the others
            print(json.dumps({"config": name, "error": repr(ex)}), flush=True)


if __name__ == "__main__":
    main()
