"""Generates tests/golden/mp_golden.npz from the REFERENCE engine (oracle/_ref/libmicropolis_ref.so,
compiled from /root/reference by oracle/build_ref.sh) driven through the oracle gym layer
(oracle/ref_gym.py, pinned against the reference's TileMap class by tests/test_gym_oracle.py).
Run once here: python tests/golden/make_mp_golden.py"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_engine import RefEngine, PARITY_FIELDS  # noqa: E402
from oracle.ref_gym import RefEnv  # noqa: E402

CASES = [("t16", 16, True, 0, 200), ("e16", 16, False, 1, 120), ("t32", 32, True, 2, 80)]


def state_crc(engine):
    c = 0
    for f in PARITY_FIELDS:
        c = zlib.crc32(engine.get(f).tobytes(), c)
    s = engine.scalars().as_dict()
    c = zlib.crc32(repr(sorted((k, v) for k, v in s.items() if not isinstance(v, float))).encode(), c)
    return c


def main():
    out = {}
    for name, size, terrain, seed, steps in CASES:
        r = RefEngine()
        env = RefEnv(r, size, max_step=10 ** 6, empty_start=not terrain)
        rng = np.random.default_rng(seed)
        env.reset_begin(1000 + seed, seed)
        if not terrain:
            r.seed(seed)
        obs0 = env.reset_finish()
        acts = rng.integers(0, env.num_tools * size * size, steps)
        rewards, crcs, obs_crc = [], [], []
        for t in range(steps):
            o, rew, d, _ = env.step(int(acts[t]))
            rewards.append(rew)
            crcs.append(state_crc(r))
            obs_crc.append(zlib.crc32(o.astype(np.float32).tobytes()))
        out[name + "_acts"] = acts
        out[name + "_rewards"] = np.array(rewards, np.float64)
        out[name + "_state_crc"] = np.array(crcs, np.uint32)
        out[name + "_obs_crc"] = np.array(obs_crc, np.uint32)
        out[name + "_obs0_crc"] = np.array([zlib.crc32(obs0.astype(np.float32).tobytes())], np.uint32)
        out[name + "_final_map"] = r.get("MPF_MAP")
        out[name + "_final_obs"] = o.astype(np.float32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mp_golden.npz"), **out)
    print("wrote mp_golden.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
