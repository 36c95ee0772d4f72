"""GPU parity for the gym layer: gymcity_b200.MicropolisEnv (mpb_env_* kernels) against the oracle
gym layer over the reference engine -- observation (32,W,H), reward, done, shadow map, engine state."""
import numpy as np
import pytest

from oracle.ref_engine import RefEngine, diff_snapshots, ref_available
from oracle.ref_gym import AGENT_TOOLS, RefEnv

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built")]


def _batch(n, size, terrain, max_step, seeds, power_puzzle=False, xs=16, ys=8):
    from gymcity_b200.micropolis_env import MicropolisEnv
    env = MicropolisEnv(num_envs=n)
    env.setMapSize(size, max_step=max_step, empty_start=not terrain, power_puzzle=power_puzzle, MAP_XS=xs, MAP_YS=ys)
    refs = [RefEnv(RefEngine(), size, max_step=max_step, empty_start=not terrain, power_puzzle=power_puzzle,
                   xs=xs, ys=ys) for _ in range(n)]
    m = env.micro
    if terrain:
        m.newMap(seeds + 1000, seeds)
    else:
        m.clearMap()
        m.engine.seed(seeds)
    for i, r in enumerate(refs):
        r.reset_begin(int(seeds[i]) + 1000, int(seeds[i]))
        if not terrain:
            r.micro.engine.seed(int(seeds[i]))
    return env, refs


def _finish(env, refs):
    m = env.micro
    m._ck(m._L.mpb_env_reset_finish(m._h, None, m.engine._stream()))
    obs = m.obs.cpu().numpy()
    for i, r in enumerate(refs):
        assert np.array_equal(r.reset_finish().astype(np.float32), obs[i]), i


def _compare_step(env, refs, acts, sb, t):
    import torch
    obs, rew, done, _ = env.step(torch.from_numpy(acts), static_build=sb)
    obs, rew, done = obs.cpu().numpy(), rew.cpu().numpy().reshape(-1), done.cpu().numpy()
    for i, r in enumerate(refs):
        o, rr, d, _ = r.step(int(acts[i]), sb)
        assert np.array_equal(o.astype(np.float32), obs[i]), (t, i, np.argwhere(o.astype(np.float32) != obs[i])[:4])
        assert rr == rew[i] and d == bool(done[i]), (t, i, rr, rew[i], d, done[i])


@pytest.mark.parametrize("size,terrain,n,steps", [(16, True, 6, 80), (16, False, 4, 60), (32, True, 3, 40),
                                                   ((20, 12), False, 3, 40)])
def test_step_matches_oracle(size, terrain, n, steps):
    seeds = np.arange(n) + 5
    env, refs = _batch(n, size, terrain, steps - 5, seeds)
    _finish(env, refs)
    W, H = refs[0].MAP_X, refs[0].MAP_Y
    rng = np.random.default_rng(1)
    for t in range(steps):
        acts = rng.integers(0, 20 * W * H, n)
        acts[rng.random(n) < 0.05] = -1
        _compare_step(env, refs, acts.astype(np.int64), bool(rng.random() < 0.2), t)
    ctr = env.micro.counters()
    for i, r in enumerate(refs):
        assert not diff_snapshots(r.micro.engine.snapshot(), env.micro.engine.snapshot(i))
        z, s, c, k = env.micro.map._shadow(i)
        assert np.array_equal(z, r.micro.map.zoneMap[-1]) and np.array_equal(s, r.micro.map.static_builds[0])
        assert int(ctr[i, 0]) == r.micro.map.num_plants and int(ctr[i, 3]) == r.micro.total_traffic
    env.close()


def test_power_puzzle_variant():
    """BASELINE config 3 shape: 32x32, Wire-only action space, 5 static Residential + 1 static Nuclear."""
    n, size = 4, 32
    seeds = np.arange(n) + 20
    env, refs = _batch(n, size, True, 50, seeds, power_puzzle=True)
    rng = np.random.default_rng(3)
    res, nuke = AGENT_TOOLS.index('Residential'), AGENT_TOOLS.index('NuclearPowerPlant')
    for k in range(9):
        xs, ys = rng.integers(0, size, n), rng.integers(0, size, n)
        tool = res if k < 5 else nuke
        env.micro.apply_actions(env.micro._flat(tool, xs, ys), True, full_tools=True)
        for i, r in enumerate(refs):
            r.micro.map.addZoneBot(int(xs[i]), int(ys[i]), AGENT_TOOLS[tool], static_build=True)
    _finish(env, refs)
    assert env.action_space.n == size * size
    for t in range(30):
        acts = rng.integers(0, size * size, n).astype(np.int64)
        _compare_step(env, refs, acts, False, t)
    env.close()


def test_full_world_window():
    """BASELINE config 4 addressing: 120x100 window at offset (0, 0)."""
    n = 2
    seeds = np.arange(n) + 30
    env, refs = _batch(n, (120, 100), True, 100, seeds, xs=0, ys=0)
    _finish(env, refs)
    rng = np.random.default_rng(4)
    for t in range(6):
        acts = rng.integers(0, 20 * 12000, n).astype(np.int64)
        _compare_step(env, refs, acts, False, t)
    env.close()


def test_vec_env_auto_reset_and_sharding_invariance():
    """done -> auto reset (subproc_vec_env.py:16-21); per-env seeds derive from the GLOBAL env index so
    a shard [4:8) of an 8-env batch equals envs 4..7 of the full batch."""
    import torch
    from gymcity_b200.micropolis_env import MicropolisVecEnv
    full = MicropolisVecEnv(8, 16, seed=3, max_step=4, empty_start=False)
    shard = MicropolisVecEnv(4, 16, seed=3, max_step=4, empty_start=False, env_offset=4)
    o_full, o_shard = full.reset(), shard.reset()
    assert torch.equal(o_full[4:], o_shard)
    rng = np.random.default_rng(0)
    for t in range(11):
        a = rng.integers(0, 20 * 256, 8)
        of, rf, df, _ = full.step(torch.from_numpy(a))
        os_, rs, ds, _ = shard.step(torch.from_numpy(a[4:]))
        assert torch.equal(of[4:], os_) and torch.equal(rf[4:], rs) and torch.equal(df[4:], ds)
        assert bool(df.all().item()) == (t % 5 == 4), (t, df)
    full.close()
    shard.close()


def test_step_host_end_to_end():
    from gymcity_b200.micropolis_env import MicropolisEnv
    n = 64
    env = MicropolisEnv(num_envs=n)
    env.seed(1)
    env.setMapSize(16, max_step=100, empty_start=False)
    env.reset()
    rng = np.random.default_rng(0)
    a = rng.integers(0, 20 * 256, n).astype(np.int64)
    rew, done = np.zeros(n, np.float32), np.zeros(n, np.uint8)
    env.step_host(a, rew, done)
    assert np.array_equal(rew, env.micro.reward.cpu().numpy().reshape(-1))
    assert not done.any()
    env.close()


def test_1000_tick_seeded_rollout_bit_exact():
    """The north-star parity bar: 1000 env-ticks (200 000 simFrames) of seeded random actions, tile map,
    power map, pollution / land value / crime / density maps, population and RCI demand bit-exact against
    the reference engine, checked every 100 ticks and at the end (observation and reward every tick)."""
    n, size, steps = 4, 16, 1000
    seeds = np.arange(n) + 31
    env, refs = _batch(n, size, True, 10 ** 9, seeds)
    _finish(env, refs)
    rng = np.random.default_rng(7)
    for t in range(steps):
        acts = rng.integers(0, 20 * size * size, n).astype(np.int64)
        _compare_step(env, refs, acts, False, t)
        if (t + 1) % 100 == 0:
            for i, r in enumerate(refs):
                d = diff_snapshots(r.micro.engine.snapshot(), env.micro.engine.snapshot(i))
                assert not d, (t, i, d)
    env.close()


def test_poet_observation_on_gpu():
    """MicropolisEnv.setMapSize(poet=True) (env.py:157-158, 301-313) against the oracle: 38 planes."""
    import torch
    from gymcity_b200.micropolis_env import MicropolisEnv
    n, size = 3, 16
    seeds = np.arange(n) + 70
    env = MicropolisEnv(num_envs=n)
    env.setMapSize(size, max_step=10 ** 6, empty_start=False, poet=True)
    env.set_params({'res_pop': 320, 'com_pop': 77, 'ind_pop': 13, 'traffic': 900, 'num_plants': 3, 'mayor_rating': 61})
    assert env.observation_space.shape == (38, size, size)
    refs = [RefEnv(RefEngine(), size, max_step=10 ** 6, empty_start=False) for _ in range(n)]
    m = env.micro
    m.newMap(seeds + 1000, seeds)
    for i, r in enumerate(refs):
        r.poet = True
        r.city_trgs = dict(env.city_trgs)
        r.reset_begin(int(seeds[i]) + 1000, int(seeds[i]))
    m._ck(m._L.mpb_env_reset_finish(m._h, None, m.engine._stream()))
    for i, r in enumerate(refs):
        assert np.array_equal(r.reset_finish().astype(np.float32), m.obs[i].cpu().numpy())
    rng = np.random.default_rng(2)
    for t in range(25):
        acts = rng.integers(0, 20 * size * size, n).astype(np.int64)
        obs, rew, done, _ = env.step(torch.from_numpy(acts))
        assert tuple(obs.shape) == (n, 38, size, size)
        obs = obs.cpu().numpy()
        for i, r in enumerate(refs):
            o, rr, d, _ = r.step(int(acts[i]))
            assert np.array_equal(o.astype(np.float32), obs[i]), (t, i)
            assert rr == float(rew[i, 0].item())
    env.close()


def test_random_static_start_runs_full_steps_per_env():
    """MicropolisEnv(random_builds=True) reset (env.py:228-241, 276-278): every env takes its own number of full
    static-build steps (action, tick, num_step + 1) before the closing simTick -- mpb_env_step_masked."""
    import torch
    from gymcity_b200.micropolis_env import MicropolisEnv
    n, size = 5, 16
    ks = np.array([0, 3, 1, 5, 2])
    rng = np.random.default_rng(8)
    acts = rng.integers(0, 20 * size * size, (int(ks.max()), n))
    env = MicropolisEnv(num_envs=n)
    env.seed(9)
    env.setMapSize(size, max_step=10 ** 6, empty_start=False, random_builds=True)
    obs0 = env.reset(static_start=(ks, acts)).cpu().numpy()
    from gymcity_b200.sharding import mix
    refs = []
    for i in range(n):
        r = RefEnv(RefEngine(), size, max_step=10 ** 6, empty_start=False)
        r.reset_begin(int(np.int32(mix(9, i, 0, 0))), int(np.int32(mix(9, i, 0, 1))))
        r.random_static_start(ks[i], acts[:, i])
        assert np.array_equal(r.reset_finish().astype(np.float32), obs0[i]), i
        assert r.num_step == ks[i]
        refs.append(r)
    ctr = env.micro.counters()
    assert np.array_equal(ctr[:, 4].astype(int), ks)                       # num_step per env
    assert obs0[:, 31].sum() > 0                                           # static builds exist
    for t in range(10):
        a = rng.integers(0, 20 * size * size, n).astype(np.int64)
        obs, rew, done, _ = env.step(torch.from_numpy(a))
        for i, r in enumerate(refs):
            o, rr, d, _ = r.step(int(a[i]))
            assert np.array_equal(o.astype(np.float32), obs[i].cpu().numpy()) and rr == float(rew[i, 0].item()), (t, i)
    for i, r in enumerate(refs):
        assert not diff_snapshots(r.micro.engine.snapshot(), env.micro.engine.snapshot(i))
    env.close()


def test_soak_48_envs_250_ticks():
    """Breadth for the rare paths (earthquakes, floods, meltdowns, the lane-parallel flood fill and dense scan):
    48 cities x 250 env-ticks of random actions against the reference engine, every observation and reward,
    full state at the end."""
    n, size, steps = 48, 16, 250
    seeds = np.arange(n) + 500
    env, refs = _batch(n, size, True, 10 ** 9, seeds)
    _finish(env, refs)
    rng = np.random.default_rng(77)
    for t in range(steps):
        acts = rng.integers(0, 20 * size * size, n).astype(np.int64)
        _compare_step(env, refs, acts, False, t)
    for i, r in enumerate(refs):
        d = diff_snapshots(r.micro.engine.snapshot(), env.micro.engine.snapshot(i))
        assert not d, (i, d[:5])
    env.close()


def test_masked_reset_leaves_other_envs_untouched():
    """A partial reset (mpb_env_reset_begin / _finish with a mask) ticks only the envs it resets: the reference
    keeps every env in its own worker, and a reset there advances nobody else (env.py:264-279)."""
    import torch
    n, size = 4, 16
    seeds = np.arange(n) + 70
    env, refs = _batch(n, size, True, 1000, seeds)
    _finish(env, refs)
    rng = np.random.default_rng(9)
    for t in range(12):
        _compare_step(env, refs, rng.integers(0, 20 * size * size, n).astype(np.int64), False, t)
    before = [env.micro.engine.snapshot(i) for i in range(n)]
    mask = np.array([1, 0, 0, 1], bool)
    env.micro.newMap(seeds + 2000, seeds + 500, mask=mask)
    mk = np.ascontiguousarray(mask.astype(np.uint8))
    env.micro._ck(env.micro._L.mpb_env_reset_finish(env.micro._h, mk.ctypes.data, env.micro.engine._stream()))
    obs = env.micro.obs.cpu().numpy()
    for i, r in enumerate(refs):
        if mask[i]:
            r.reset_begin(int(seeds[i]) + 2000, int(seeds[i]) + 500)
            assert np.array_equal(r.reset_finish().astype(np.float32), obs[i]), i
        else:
            assert not diff_snapshots(before[i], env.micro.engine.snapshot(i)), i
    for t in range(12):
        _compare_step(env, refs, rng.integers(0, 20 * size * size, n).astype(np.int64), False, 100 + t)
    for i, r in enumerate(refs):
        assert not diff_snapshots(r.micro.engine.snapshot(), env.micro.engine.snapshot(i))
    env.close()


def test_device_auto_reset_matches_reference_workers():
    """MicropolisVecEnv with the auto-reset inside the step (mpb_env_set_auto_reset): episodes of different envs
    end on different steps (a masked reset up front staggers them), each finished env comes back with its reset
    observation and the terminal step's reward / done (subproc_vec_env.py:16-21), its next episode is seeded with
    sharding.mix(seed, global index, episode, which), and nobody else is touched -- against one reference env per
    worker doing exactly that."""
    import torch
    from gymcity_b200.micropolis_env import MicropolisVecEnv
    from gymcity_b200.sharding import mix
    n, size, max_step, seed, off = 6, 16, 7, 11, 40
    env = MicropolisVecEnv(n, size, seed=seed, max_step=max_step, empty_start=False, env_offset=off)
    assert env._device_auto_reset
    refs = [RefEnv(RefEngine(), size, max_step=max_step, empty_start=False) for _ in range(n)]
    episode = np.zeros(n, int)

    def ref_reset(i):
        refs[i].reset_begin(mix(seed, off + i, episode[i], 0), mix(seed, off + i, episode[i], 1))
        episode[i] += 1
        return refs[i].reset_finish().astype(np.float32)

    obs = env.reset().cpu().numpy()
    for i in range(n):
        assert np.array_equal(ref_reset(i), obs[i]), i
    rng = np.random.default_rng(5)
    n_resets, ended = 0, [[] for _ in range(n)]
    for t in range(40):
        if t == 3:                                           # stagger: envs 1 and 4 start a new episode now
            mask = np.array([0, 1, 0, 0, 1, 0], bool)
            obs = env.reset(mask=mask).cpu().numpy()
            for i in np.flatnonzero(mask):
                assert np.array_equal(ref_reset(i), obs[i]), i
        acts = rng.integers(0, 20 * size * size, n).astype(np.int64)
        obs, rew, done, infos = env.step(torch.from_numpy(acts))
        obs, rew, done = obs.cpu().numpy(), rew.cpu().numpy().reshape(-1), done.cpu().numpy()
        assert len(infos) == n and infos[0] == {}
        for i, r in enumerate(refs):
            o, rr, d, _ = r.step(int(acts[i]), False)
            assert rr == rew[i] and d == bool(done[i]), (t, i, rr, rew[i], d, done[i])
            if d:
                o = ref_reset(i)
                n_resets += 1
                ended[i].append(t)
            assert np.array_equal(np.asarray(o, np.float32), obs[i]), (t, i)
    assert n_resets >= 20 and ended[0] != ended[1] and ended[1] == ended[4]   # staggered episodes, several resets each
    assert np.array_equal(env.micro.counters()[:, 16], episode)
    on, served, inline = env.micro.prefetch_stats()                   # most resets copied a map generated ahead of time
    assert on and served + inline == n_resets + 2 + n and served >= n_resets // 2, (served, inline, n_resets)
    for i, r in enumerate(refs):
        assert not diff_snapshots(r.micro.engine.snapshot(), env.micro.engine.snapshot(i))
    env.close()


def test_sticky_obs_buffer_and_per_call_out():
    """set_obs_buffer(X) stays in force for plain step(a) calls; step(a, out=Y) switches to Y."""
    import torch
    from gymcity_b200.micropolis_env import MicropolisVecEnv
    env = MicropolisVecEnv(4, 16, seed=1, max_step=50, empty_start=False)
    env.reset()
    dev = env.micro.obs.device
    X, Y = torch.zeros_like(env.micro.obs), torch.zeros_like(env.micro.obs)
    env.set_obs_buffer(X)
    a = torch.zeros(4, dtype=torch.int64, device=dev)
    o1, *_ = env.step(a)
    assert o1.data_ptr() == X.data_ptr() and float(X.abs().sum()) > 0
    o2, *_ = env.step(a)
    assert o2.data_ptr() == X.data_ptr()
    o3, *_ = env.step(a, out=Y)
    assert o3.data_ptr() == Y.data_ptr() and float(Y.abs().sum()) > 0
    env.set_obs_buffer(None)
    o4, *_ = env.step(a)
    assert o4.data_ptr() == env.micro.obs.data_ptr()
    env.close()


def test_handles_on_threads_do_not_share_state():
    """Library hygiene: the scheduler / profiling state lives in the handle, so several handles driven from several
    threads of one process (one learner process with a handle per GPU) run the rollouts they run alone.  Four handles
    of different batch sizes (8-, 16- and 32-city CTAs, different group geometry) step concurrently on four threads
    and streams; each rollout is compared with the same handle configuration run on its own."""
    import threading
    import torch
    from gymcity_b200.micropolis_env import MicropolisVecEnv

    sizes = [40, 700, 2500, 9600]

    def rollout(n, out, k, barrier=None):
        torch.cuda.set_device(0)
        with torch.cuda.stream(torch.cuda.Stream()):
            env = MicropolisVecEnv(n, 16, seed=40 + k, max_step=6, empty_start=False)
            env.reset()
            rng = np.random.default_rng(k)
            acc = torch.zeros(n, 1, device="cuda")
            if barrier is not None:
                barrier.wait()
            for t in range(14):
                obs, rew, done, _ = env.step(torch.from_numpy(rng.integers(0, 20 * 256, n)))
                acc += rew
            torch.cuda.current_stream().synchronize()
            out[k] = (obs.sum(dim=(1, 2, 3)).cpu().numpy(), acc.cpu().numpy(), env.micro.counters())
            env.close()

    alone, together = {}, {}
    for k, n in enumerate(sizes):
        rollout(n, alone, k)
    barrier = threading.Barrier(len(sizes))
    threads = [threading.Thread(target=rollout, args=(n, together, k, barrier)) for k, n in enumerate(sizes)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    for k in range(len(sizes)):
        for a, b in zip(alone[k], together[k]):
            assert np.array_equal(a, b), k
