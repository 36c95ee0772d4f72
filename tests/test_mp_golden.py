"""Committed golden vectors (tests/golden/mp_golden.npz, made by make_mp_golden.py from the reference
engine): the CPU harness must reproduce them here, the CUDA path on the GPU box."""
import os
import zlib

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "mp_golden.npz"))
CASES = [("t16", 16, True, 0), ("e16", 16, False, 1), ("t32", 32, True, 2)]
PARITY = None


def _crc_of_snapshot(get, scalars):
    from oracle.ref_engine import PARITY_FIELDS
    c = 0
    for f in PARITY_FIELDS:
        c = zlib.crc32(get(f).tobytes(), c)
    c = zlib.crc32(repr(sorted((k, v) for k, v in scalars.items() if not isinstance(v, float))).encode(), c)
    return c


@pytest.mark.parametrize("name,size,terrain,seed", CASES)
def test_host_harness_reproduces_golden(name, size, terrain, seed):
    from tests.host_engine import HostGym
    h = HostGym()
    h.configure(size, size, 10 ** 6)
    h.reset_begin(terrain, 1000 + seed, seed)
    if not terrain:
        h.seed(seed)
    o = h.reset_finish()
    assert zlib.crc32(o.tobytes()) == int(G[name + "_obs0_crc"][0])
    acts = G[name + "_acts"]
    for t in range(len(acts)):
        o, rew, d = h.step(int(acts[t]))
        assert rew == G[name + "_rewards"][t], t
        assert zlib.crc32(o.tobytes()) == int(G[name + "_obs_crc"][t]), t
        assert _crc_of_snapshot(h.get, h.scalars().as_dict()) == int(G[name + "_state_crc"][t]), t
    assert np.array_equal(h.get("MPF_MAP"), G[name + "_final_map"])


@pytest.mark.gpu
@pytest.mark.parametrize("name,size,terrain,seed", CASES)
def test_cuda_reproduces_golden(name, size, terrain, seed):
    import torch
    from gymcity_b200.micropolis_env import MicropolisEnv
    env = MicropolisEnv(num_envs=1)
    env.setMapSize(size, max_step=10 ** 6, empty_start=not terrain)
    m = env.micro
    if terrain:
        m.newMap([1000 + seed], [seed])
    else:
        m.clearMap()
        m.engine.seed([seed])
    m._ck(m._L.mpb_env_reset_finish(m._h, None, m.engine._stream()))
    o = m.obs[0].cpu().numpy()
    assert zlib.crc32(o.tobytes()) == int(G[name + "_obs0_crc"][0])
    acts = G[name + "_acts"]
    for t in range(len(acts)):
        obs, rew, d, _ = env.step(torch.tensor([int(acts[t])]))
        assert rew == G[name + "_rewards"][t], t
        assert zlib.crc32(obs.astype(np.float32).tobytes()) == int(G[name + "_obs_crc"][t]), t
        if t % 10 == 9 or t == len(acts) - 1:
            c = _crc_of_snapshot(lambda f: m.engine.get(0, f), m.engine.scalars(0).as_dict())
            assert c == int(G[name + "_state_crc"][t]), t
    assert np.array_equal(m.engine.get(0, "MPF_MAP"), G[name + "_final_map"])
    assert np.array_equal(m.obs[0].cpu().numpy(), G[name + "_final_obs"])
    env.close()
