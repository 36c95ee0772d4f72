"""MicropolisPaintEnv (gym_city/envs/paintenv.py, paintcontrol.py): one tool per window cell and step.
Device code on the CPU harness and mpb_env_step_paint on the GPU against oracle/ref_gym.RefPaintEnv over the
reference engine: observation, reward, done, shadow map and engine state."""
import numpy as np
import pytest

from oracle.ref_engine import RefEngine, diff_snapshots, ref_available
from oracle.ref_gym import RefPaintEnv
from tests.host_engine import HostGym

pytestmark = pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built")


def _paint_actions(rng, n, W, H, p_nil=0.7):
    a = rng.integers(0, 19, (n, W, H))
    a[rng.random((n, W, H)) < p_nil] = 19              # mostly 'Nil': a full random repaint every step erases everything
    return a.astype(np.uint8)


@pytest.mark.parametrize("seed,size,terrain", [(0, 12, True), (1, 16, False)])
def test_harness_paint_step_matches_oracle(seed, size, terrain):
    r = RefEngine()
    env = RefPaintEnv(r, size, max_step=10 ** 6, empty_start=not terrain)
    h = HostGym()
    h.configure(size, size, 10 ** 6)
    env.reset_begin(1000 + seed, seed)
    h.reset_begin(terrain, 1000 + seed, seed)
    if not terrain:
        r.seed(seed)
        h.seed(seed)
    assert np.array_equal(env.reset_finish().astype(np.float32), h.reset_finish())
    rng = np.random.default_rng(seed)
    for t in range(14):
        a = _paint_actions(rng, 1, size, size)[0]
        o0, r0, d0, _ = env.step(a)
        o1, r1, d1 = h.step_paint(a)
        assert np.array_equal(o0.astype(np.float32), o1), (t, np.argwhere(o0.astype(np.float32) != o1)[:4])
        assert r0 == r1 and d0 == d1, t
        z, s, c, k = h.shadow()
        assert np.array_equal(z, env.micro.map.zoneMap[-1]) and np.array_equal(s, env.micro.map.static_builds[0])
        assert int(k[0]) == env.micro.map.num_plants and int(k[1]) == env.micro.map.num_roads
    # the last column is never painted (paintcontrol.py:52-54)
    assert not diff_snapshots(r.snapshot(), h.snapshot())


@pytest.mark.gpu
@pytest.mark.parametrize("pre_simt", ["1", "2"])       # 2: the paint action on a thread per env (k_mp_env_pre_t), as batches of >= 4096 envs run it
def test_gpu_paint_env_matches_oracle(pre_simt, monkeypatch):
    import torch
    from gymcity_b200.micropolis_env import MicropolisPaintEnv
    monkeypatch.setenv("MPB_PRE_SIMT", pre_simt)
    n, size, steps = 3, 12, 12
    seeds = np.arange(n) + 60
    env = MicropolisPaintEnv(n, size, max_step=10 ** 6, empty_start=False)
    assert env.action_space.shape == (size, size) and int(env.action_space.high.max()) == 19
    refs = [RefPaintEnv(RefEngine(), size, max_step=10 ** 6, empty_start=False) for _ in range(n)]
    m = env.micro
    m.newMap(seeds + 1000, seeds)
    for i, r in enumerate(refs):
        r.reset_begin(int(seeds[i]) + 1000, int(seeds[i]))
    m._ck(m._L.mpb_env_reset_finish(m._h, None, m.engine._stream()))
    for i, r in enumerate(refs):
        assert np.array_equal(r.reset_finish().astype(np.float32), m.obs[i].cpu().numpy())
    rng = np.random.default_rng(5)
    for t in range(steps):
        a = _paint_actions(rng, n, size, size)
        obs, rew, done, _ = env.step(torch.from_numpy(a))
        obs = obs.cpu().numpy()
        for i, r in enumerate(refs):
            o, rr, d, _ = r.step(a[i])
            assert np.array_equal(o.astype(np.float32), obs[i]), (t, i)
            assert rr == float(rew[i, 0].item()) and d == bool(done[i].item()), (t, i)
    for i, r in enumerate(refs):
        assert not diff_snapshots(r.micro.engine.snapshot(), m.engine.snapshot(i))
        z, s, c, k = m.map._shadow(i)
        assert np.array_equal(z, r.micro.map.zoneMap[-1]) and np.array_equal(s, r.micro.map.static_builds[0])
    env.close()
