"""Extinguisher wrapper (gym_city/wrappers.py:10-141): the device code (CPU harness for the `not gpu` suite,
mpb_env_extinguish on the GPU) against oracle/ref_gym.RefExtinguisher over the reference engine -- age_order,
deletions, shadow map, observation and engine state after every event."""
import numpy as np
import pytest

from oracle.ref_engine import RefEngine, diff_snapshots, ref_available
from oracle.ref_gym import RefEnv, RefExtinguisher
from tests.host_engine import HostGym

pytestmark = pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built")
TYPES = ['age', 'spatial', 'random', 'Monster']


@pytest.mark.parametrize("xt,seed,terrain", [('age', 0, True), ('spatial', 1, True), ('random', 2, False),
                                             ('Monster', 3, True), ('age', 4, False)])
def test_harness_extinction_matches_oracle(xt, seed, terrain):
    size, steps, n_dels = 16, 60, 7
    r = RefEngine()
    env = RefEnv(r, size, max_step=10 ** 6, empty_start=not terrain)
    wrap = RefExtinguisher(env, xt, extinction_prob=0.1, xt_dels=n_dels)
    h = HostGym()
    h.configure(size, size, 10 ** 6)
    h.track_ages()
    env.reset_begin(1000 + seed, seed)
    h.reset_begin(terrain, 1000 + seed, seed)
    if not terrain:
        r.seed(seed)
        h.seed(seed)
    assert np.array_equal(env.reset_finish().astype(np.float32), h.reset_finish())
    rng = np.random.default_rng(seed)
    events = 0
    for t in range(steps):
        a = int(rng.integers(0, env.num_tools * size * size))
        rnd = rng.integers(0, 2 ** 31 - 1, n_dels + 2)
        (o0, r0, d0, _), dels0 = wrap.step(a, rnd)
        o1, r1, d1 = h.step(a)
        assert np.array_equal(o0.astype(np.float32), o1) and r0 == r1, t
        if dels0 is not None:
            events += 1
            dels1 = h.extinguish(TYPES.index(xt), n_dels, rnd)
            assert dels0 == dels1, (t, dels0, dels1)
        assert np.array_equal(env.micro.map.age_order, h.ages()), t
        z, s, c, k = h.shadow()
        assert np.array_equal(z, env.micro.map.zoneMap[-1]) and np.array_equal(s, env.micro.map.static_builds[0])
    assert events == steps // 10
    assert not diff_snapshots(r.snapshot(), h.snapshot())


@pytest.mark.gpu
@pytest.mark.parametrize("xt", TYPES)
def test_gpu_extinguisher_matches_oracle(xt):
    import torch
    from gymcity_b200.micropolis_env import MicropolisEnv
    from gymcity_b200.wrappers import Extinguisher
    n, size, steps, n_dels = 5, 16, 45, 6
    seeds = np.arange(n) + 40
    env = MicropolisEnv(num_envs=n)
    env.setMapSize(size, max_step=10 ** 6, empty_start=False)
    wrap = Extinguisher(env, xt, extinction_prob=0.2, xt_dels=n_dels)
    refs = [RefEnv(RefEngine(), size, max_step=10 ** 6, empty_start=False) for _ in range(n)]
    rwraps = [RefExtinguisher(r, xt, extinction_prob=0.2, xt_dels=n_dels) for r in refs]
    m = env.micro
    m.newMap(seeds + 1000, seeds)
    for i, r in enumerate(refs):
        r.reset_begin(int(seeds[i]) + 1000, int(seeds[i]))
    m._ck(m._L.mpb_env_reset_finish(m._h, None, m.engine._stream()))
    for i, r in enumerate(refs):
        assert np.array_equal(r.reset_finish().astype(np.float32), m.obs[i].cpu().numpy())
    rng = np.random.default_rng(11)
    events = 0
    for t in range(steps):
        acts = rng.integers(0, 20 * size * size, n).astype(np.int64)
        rnd = rng.integers(0, 2 ** 31 - 1, (n, n_dels + 2))
        obs, rew, done, _ = env.step(torch.from_numpy(acts))
        env_fires = env.num_step % wrap.extinction_interval == 0
        dels = wrap.extinguish(xt, rnd) if env_fires else None
        obs = obs.cpu().numpy()
        for i, rw in enumerate(rwraps):
            (o, rr, d, _), dels0 = rw.step(int(acts[i]), rnd[i])
            assert np.array_equal(o.astype(np.float32), obs[i]) and rr == float(rew[i, 0].item()), (t, i)
            assert (dels0 is not None) == env_fires
            if env_fires:
                assert dels0 == int(dels[i]), (t, i, dels0, dels)
            assert np.array_equal(refs[i].micro.map.age_order, wrap.ages(i)), (t, i)
        events += int(env_fires)
    assert events == steps // 5
    for i, r in enumerate(refs):
        assert not diff_snapshots(r.micro.engine.snapshot(), m.engine.snapshot(i))
        z, s, c, k = m.map._shadow(i)
        assert np.array_equal(z, r.micro.map.zoneMap[-1]) and np.array_equal(s, r.micro.map.static_builds[0])
    env.close()


@pytest.mark.gpu
def test_extinguisher_wrapper_step_triggers_on_interval():
    import torch
    from gymcity_b200.micropolis_env import MicropolisEnv
    from gymcity_b200.wrappers import Extinguisher
    n, size = 8, 16
    env = MicropolisEnv(num_envs=n)
    env.seed(3)
    env.setMapSize(size, max_step=10 ** 6, empty_start=False)
    wrap = Extinguisher(env, 'age', extinction_prob=0.25, xt_dels=50, seed=1)
    wrap.reset()
    g = np.random.default_rng(0)
    built = []
    for t in range(8):
        wrap.step(torch.from_numpy(g.integers(0, 20 * size * size, n)))
        built.append(int((wrap.ages(0) > -1).sum()))
    # every 4th step the 50 oldest cells go: nothing aged is left right after an event
    assert built[3] == 0 and built[7] == 0 and max(built) > 0
    env.close()


@pytest.mark.gpu
def test_extinction_follows_each_envs_own_step_count():
    """Extinguisher over the auto-resetting vec env: the event hits env i when ITS num_step is a multiple of the interval
    (episodes staggered by a masked reset), and never the step on which the env was auto-reset (num_step 0: in the
    reference the wrapper sits below the worker's reset)."""
    import torch
    from gymcity_b200.micropolis_env import MicropolisVecEnv
    from gymcity_b200.wrappers import Extinguisher
    n, size = 4, 16
    env = MicropolisVecEnv(n, size, seed=9, max_step=6, empty_start=False)
    wrap = Extinguisher(env, 'age', extinction_prob=0.25, xt_dels=300, seed=1)      # interval 4: every aged cell goes
    wrap.reset()
    g = np.random.default_rng(0)
    road = 8 * size * size                                                            # tool 8 ('Road') always leaves an aged cell
    for t in range(2):
        wrap.step(torch.from_numpy(road + g.integers(0, size * size, n)))
    wrap.reset(mask=np.array([0, 1, 0, 1], bool))                                     # envs 1 and 3 are two steps behind now
    seen = []
    prev = np.array([int((wrap.ages(i) > -1).sum()) for i in range(n)])
    hits = 0
    for t in range(14):
        obs, rew, done, _ = wrap.step(torch.from_numpy(road + g.integers(0, size * size, n)))
        k = env.micro.counters()[:, 4].astype(int)                                    # num_step after the step (0: just reset)
        aged = np.array([int((wrap.ages(i) > -1).sum()) for i in range(n)])
        seen.append((k.copy(), aged.copy()))
        for i in range(n):
            if k[i] > 0 and k[i] % 4 == 0:
                assert aged[i] <= 2 and aged[i] <= prev[i], (t, i, k, aged, prev)     # hit: (all but a stubborn cell or two of) the aged cells are gone
                hits += 1
            elif k[i] > 1:
                assert aged[i] >= prev[i], (t, i, k, aged, prev)                      # not hit: nothing was taken away
        prev = aged
    assert hits >= 8
    ks = np.stack([s[0] for s in seen])
    assert (ks[:, 0] != ks[:, 1]).any() and (ks == 0).any()                           # staggered, and auto-resets happened
    env.close()
