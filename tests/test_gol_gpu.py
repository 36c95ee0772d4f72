"""GPU parity: golb_* kernels (through the Python mirror of the reference env API) against the
oracle and the committed reference-generated golden vectors.  Bit-exact state and population;
reward bit-exact in fp32 for the multi env."""
import os

import numpy as np
import pytest

from oracle import gol_oracle as go

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "gol_golden.npz"))


def _multi(n, w, max_step=200):
    from gymcity_b200 import GoLMultiEnv
    env = GoLMultiEnv()
    env.configure(map_width=w, num_proc=n, max_step=max_step)
    return env


@pytest.mark.parametrize("key", ["multi_8_16", "multi_4_5", "multi_3_32", "multi_5_8"])
def test_multi_matches_reference_golden(key):
    import torch
    init, acts = G[key + "_init"], G[key + "_acts"]
    n, w = init.shape[0], init.shape[1]
    env = _multi(n, w)
    env.set_state(init)
    st = init.copy()
    for t in range(acts.shape[0]):
        obs, rew, done, info = env.step(torch.from_numpy(acts[t]).reshape(n, 1))
        st, orew, opop = go.multi_step(st, acts[t], w * w * 2 / 3, w * w * 2 / 3, 200, n)
        got = obs[:, 1].cpu().numpy().astype(np.uint8)
        assert np.array_equal(got, G[key + "_states"][t]), (key, t)
        assert np.array_equal(env.get_pop().cpu().numpy(), G[key + "_pops"][t])
        assert np.array_equal(rew.cpu().numpy().reshape(-1), orew), (key, t)   # fp32 bit-exact
        assert obs.shape == (n, 2, w, w) and float(obs[:, 0].min()) == 1.0
    env.close()


@pytest.mark.parametrize("n,w", [(4096, 16), (1000, 7), (513, 32), (64, 3), (300, 12)])
def test_multi_vs_oracle_random(n, w):
    import torch
    rng = np.random.default_rng(n + w)
    init = (rng.random((n, w, w)) < 0.2).astype(np.uint8)
    env = _multi(n, w, max_step=20)
    env.set_state(init)
    st = init
    for t in range(25):
        a = rng.integers(0, w * w, size=n)
        obs, rew, done, _ = env.step(torch.from_numpy(a))
        st, orew, opop = go.multi_step(st, a, w * w * 2 / 3, w * w * 2 / 3, 20, n)
        assert np.array_equal(env.get_state(), st), t
        assert np.array_equal(rew.cpu().numpy().reshape(-1), orew)
        assert bool(done.all().item()) == (t == 20)     # multi_env.py:213-216
    env.close()


def test_multi_1000_ticks_4096_envs_checksum():
    """BASELINE config 2 at full size: 4096 x 16x16, 1000 ticks, against the oracle."""
    import torch
    n, w = 4096, 16
    rng = np.random.default_rng(0)
    init = (rng.random((n, w, w)) < 0.2).astype(np.uint8)
    env = _multi(n, w, max_step=1000)
    env.set_state(init)
    st = init
    for t in range(1000):
        a = rng.integers(0, w * w, size=n)
        env.step(torch.from_numpy(a))
        st, _, opop = go.multi_step(st, a, 1, 1, 1, 1)
        if t % 100 == 99:
            assert np.array_equal(env.get_pop().cpu().numpy(), opop), t
    assert np.array_equal(env.get_state(), st)
    env.close()


@pytest.mark.parametrize("key", ["classic_16", "classic_6"])
def test_classic_matches_reference_golden(key):
    from gymcity_b200 import GameOfLifeEnv
    init, acts = G[key + "_init"], G[key + "_acts"]
    w = init.shape[0]
    env = GameOfLifeEnv()
    env.configure(map_width=w, max_step=100)
    env.set_state(init[None])
    for t in range(acts.shape[0]):
        s, r, term, info = env.step(int(acts[t]))
        assert np.array_equal(s[0], G[key + "_states"][t]), (key, t)
        assert abs(r - G[key + "_raw"][t] * 100 / (100 * w * w)) < 1e-6
        assert term == (G[key + "_raw"][t] < 2 or t == 100)
    env.close()


def test_classic_batched_vs_oracle():
    from gymcity_b200 import GameOfLifeEnv
    import torch
    n, w = 37, 9
    rng = np.random.default_rng(5)
    init = (rng.random((n, w, w)) < 0.3).astype(np.uint8)
    env = GameOfLifeEnv()
    env.configure(map_width=w, max_step=50, num_proc=n)
    env.set_state(init)
    worlds = [go.ClassicWorld(init[i]) for i in range(n)]
    for t in range(40):
        a = rng.integers(0, w * w, size=n)
        obs, rew, done, _ = env.step(torch.from_numpy(a))
        got = obs.cpu().numpy()[:, 0]
        for i in range(n):
            s, r, term = go.classic_step(worlds[i], int(a[i]), 50, t)
            assert np.array_equal(got[i], s), (t, i)
            assert bool(done[i].item()) == term
    env.close()


def test_reset_population_and_determinism():
    env = _multi(2048, 16)
    env.seed(3)
    o1 = env.reset().clone()
    frac = float(o1[:, 1].mean().item())
    assert 0.18 < frac < 0.22
    env.seed(3)
    o2 = env.reset()
    assert bool((o1 == o2).all().item())
    env.close()


def test_step_host_end_to_end():
    n, w = 512, 16
    rng = np.random.default_rng(9)
    init = (rng.random((n, w, w)) < 0.2).astype(np.uint8)
    env = _multi(n, w)
    env.set_state(init)
    a = rng.integers(0, w * w, size=n).astype(np.int64)
    rew = np.zeros(n, np.float32)
    env.step_host(a, rew)
    st, orew, _ = go.multi_step(init, a, w * w * 2 / 3, w * w * 2 / 3, 200, n)
    assert np.array_equal(rew, orew) and np.array_equal(env.get_state(), st)
    env.close()


@pytest.mark.parametrize("n,w,T", [(4096, 16, 33), (100, 8, 7), (77, 32, 5), (50, 12, 9)])
def test_step_n_equals_n_steps(n, w, T):
    """golb_step_n: T ticks in one launch leave the same state / reward / pop / done as T golb_step calls, and the
    per-tick observation and reward sequences are those of the single steps (both planes)."""
    import torch
    rng = np.random.default_rng(n * 3 + w)
    init = (rng.random((n, w, w)) < 0.25).astype(np.uint8)
    acts = torch.from_numpy(rng.integers(0, w * w, (T, n))).cuda()
    a, b = _multi(n, w, max_step=T - 2), _multi(n, w, max_step=T - 2)
    a.set_state(init); b.set_state(init)
    obs_seq = torch.full((T, n, 2, w, w), -1.0, device="cuda")
    rew_seq = torch.full((T, n), -1.0, device="cuda")
    ob, rb, db, _ = b.step_n(acts, obs_seq, rew_seq)
    for t in range(T):
        oa, ra, da, _ = a.step(acts[t])
        assert torch.equal(oa, obs_seq[t]), t
        assert torch.equal(ra.reshape(-1), rew_seq[t]), t
    assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(da, db)
    assert np.array_equal(a.get_state(), b.get_state()) and torch.equal(a.get_pop(), b.get_pop())
    assert a.step_count == b.step_count == T
    oa, ra, da, _ = a.step(acts[0]); ob, rb, db, _ = b.step_n(acts[:1])
    assert torch.equal(oa, ob) and torch.equal(da, db)
    a.close(); b.close()
