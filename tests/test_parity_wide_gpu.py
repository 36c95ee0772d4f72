"""GPU parity beyond the headline configuration ("close the parity holes"): the 24 dense cities the
reference ships, driven with whole-map tools; reference probe envs inside the 8192-env power puzzle; BASELINE
config 1 with the seeds SURVEY names (0..7) over 1000 ticks, terrain and empty start; the rare engine functions
(meltdown, sprite generators, explosions) called directly and followed for 50 frames; every tool on every terrain
through the kernels.  Everything against the reference engine compiled as the oracle, bit-exact."""
import os

import numpy as np
import pytest

from oracle.ref_engine import RefEngine, diff_snapshots, ref_available
from oracle.ref_gym import AGENT_TOOLS, RefEnv

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built")]
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _mk(n):
    from gymcity_b200.engine import EngineBatch
    return EngineBatch(n), [RefEngine() for _ in range(n)]


def _check(b, refs, tag):
    for i, r in enumerate(refs):
        d = diff_snapshots(r.snapshot(), b.snapshot(i))
        assert not d, (tag, i, d[:6])


def test_24_dense_cities_120_ticks_whole_map_tools():
    """BASELINE config 4's dense variant: the maps of the reference's 24 .cty cities (tests/golden/cty_maps.npz, read
    by the reference loader) in a live simulation, 120 env-ticks of MicropolisControl.simTick with a random tool on a
    random tile of the whole 120x100 world per tick -- big power grids, traffic, R/C/I growth, airports / seaports and
    their sprites, disasters.  Full engine state every 10 ticks."""
    z = np.load(os.path.join(GOLD, "cty_maps.npz"))
    maps = z["maps"]
    n = maps.shape[0]
    assert n == 24
    b, refs = _mk(n)
    ss = np.arange(n) + 5
    b.generate(ss + 1000, ss)
    for i, r in enumerate(refs):
        r.generate(int(ss[i]) + 1000)
        r.seed(int(ss[i]))
        r.set("MPF_MAP", maps[i])
        b.set(i, "MPF_MAP", maps[i])
    rng = np.random.default_rng(24)
    seen = set()
    for t in range(120):
        tools = rng.integers(0, 20, n)
        tools[tools == 5] = 7                                   # the query tool is a GUI read
        xs, ys = rng.integers(0, 120, n), rng.integers(0, 100, n)
        res = b.toolDown(tools, xs, ys)
        for i, r in enumerate(refs):
            assert r.toolDown(int(tools[i]), int(xs[i]), int(ys[i])) == int(res[i]), (t, i)
            r.setFunds(2000000)
            r.controlSimTick()
        b.setFunds(2000000)
        b.controlSimTick()
        if t % 10 == 9:
            _check(b, refs, "tick %d" % t)
            for r in refs[::4]:
                seen |= set(s[0] for s in r.sprites())
    _check(b, refs, "final")
    assert len(seen) >= 3, seen                                  # trains / planes / ships / copters are live in these cities
    assert int(((maps & 0x0400) != 0).sum()) > 5000              # ~5 100 zone centres over the 24 maps (badnews, haight, yokohama ...)
    assert max(r.scalars().cityPop for r in refs) > 0
    b.close()


def test_power_puzzle_8192_envs_reference_probes():
    """BASELINE config 3 at full size with reference probes: envs 0, 4097 and 8191 of the 8192-env puzzle batch against
    the oracle gym layer over the reference engine -- same derived seeds, same set-up builds (5 static Residential, 4
    static NuclearPowerPlant attempts at positions drawn here), 25 Wire actions."""
    import torch
    from gymcity_b200.micropolis_env import MicropolisVecEnv
    from gymcity_b200.sharding import mix
    n, size, seed = 8192, 32, 11
    probes = [0, 4097, 8191]
    env = MicropolisVecEnv(n, size, seed=seed, max_step=10 ** 6, empty_start=False, power_puzzle=True)
    assert env.action_space.n == size * size
    rng = np.random.default_rng(8)
    res, nuke = AGENT_TOOLS.index('Residential'), AGENT_TOOLS.index('NuclearPowerPlant')
    setup = []
    for k in range(9):
        xs, ys = rng.integers(0, size, n), rng.integers(0, size, n)
        setup.append((res if k < 5 else nuke, xs, ys))
    refs = []
    for i in probes:
        r = RefEnv(RefEngine(), size, max_step=10 ** 6, empty_start=False, power_puzzle=True)
        r.reset_begin(mix(seed, i, 0, 0), mix(seed, i, 0, 1))
        for tool, xs, ys in setup:
            r.micro.map.addZoneBot(int(xs[i]), int(ys[i]), AGENT_TOOLS[tool], static_build=True)
        refs.append((i, r, r.reset_finish().astype(np.float32)))
    obs = env.reset(setup_actions=[(env.micro._flat(tool, xs, ys), True, True) for tool, xs, ys in setup])
    for i, r, o0 in refs:
        assert np.array_equal(o0, obs[i].cpu().numpy()), i
    for t in range(25):
        a = rng.integers(0, size * size, n)
        obs, rew, done, _ = env.step(torch.from_numpy(a))
        for i, r, _ in refs:
            o, rr, d, _ = r.step(int(a[i]))
            assert np.array_equal(np.asarray(o, np.float32), obs[i].cpu().numpy()), (t, i)
            assert rr == float(rew[i, 0].item()) and d == bool(done[i].item()), (t, i)
    for i, r, _ in refs:
        assert not diff_snapshots(r.micro.engine.snapshot(), env.micro.engine.snapshot(i)), i
    assert not env.micro.counters()[:, 15].any()
    env.close()


@pytest.mark.parametrize("terrain", [True, False])
def test_config1_seeds_0_to_7_1000_ticks(terrain):
    """BASELINE config 1 with the seeds SURVEY 8d names: 8 envs, seeds 0..7, 16x16 window, 1000 seeded random-action
    env-ticks each, every reward and done, the observation and the full engine state every 100 ticks.  terrain=False
    is the reference's default empty_start (clearMap: the simulation does not run), carried the same length."""
    import torch
    from gymcity_b200.micropolis_env import MicropolisEnv
    n, size = 8, 16
    seeds = np.arange(n)
    env = MicropolisEnv(num_envs=n)
    env.setMapSize(size, max_step=10 ** 6, empty_start=not terrain)
    refs = [RefEnv(RefEngine(), size, max_step=10 ** 6, empty_start=not terrain) for _ in range(n)]
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
    m._ck(m._L.mpb_env_reset_finish(m._h, None, m.engine._stream()))
    obs = m.obs.cpu().numpy()
    for i, r in enumerate(refs):
        assert np.array_equal(r.reset_finish().astype(np.float32), obs[i]), i
    rng = np.random.default_rng(1000)
    for t in range(1000):
        acts = rng.integers(0, 20 * size * size, n).astype(np.int64)
        obs, rew, done, _ = env.step(torch.from_numpy(acts))
        rew, done = rew.cpu().numpy().reshape(-1), done.cpu().numpy()
        full = t % 100 == 99
        if full:
            obs = obs.cpu().numpy()
        for i, r in enumerate(refs):
            o, rr, d, _ = r.step(int(acts[i]), False)
            assert rr == rew[i] and d == bool(done[i]), (t, i, rr, rew[i])
            if full:
                assert np.array_equal(np.asarray(o, np.float32), obs[i]), (t, i)
                assert not diff_snapshots(r.micro.engine.snapshot(), m.engine.snapshot(i)), (t, i)
    assert not m.counters()[:, 15].any()
    env.close()


def test_rare_functions_then_50_frames():
    """makeMeltdown, the four sprite generators and makeExplosion called directly on a built, live city, each followed
    by 50 simFrames (the sprites move, the fire / radiation fields are scanned) -- functions a random rollout reaches
    only by luck."""
    n = 2
    b, refs = _mk(n)
    ss = np.arange(n) + 60
    b.generate(ss + 500, ss)
    for i, r in enumerate(refs):
        r.generate(int(ss[i]) + 500)
        r.seed(int(ss[i]))
    NUKE, AIR, PORT, RAIL, ROAD, RES, WIRE, DOZE = 14, 15, 12, 8, 9, 0, 6, 7
    acts = []
    for x in range(30, 90):
        for y in range(30, 70):
            acts.append((DOZE, x, y))
    acts += [(NUKE, 40, 40), (AIR, 50, 40), (PORT, 60, 40), (RES, 40, 50), (RES, 44, 50)]
    acts += [(RAIL, x, 48) for x in range(32, 80)] + [(ROAD, x, 46) for x in range(32, 80)]
    acts += [(WIRE, x, 45) for x in range(38, 66)]
    for tool, x, y in acts:
        res = b.toolDown([tool] * n, [x] * n, [y] * n)
        for i, r in enumerate(refs):
            assert r.toolDown(tool, x, y) == int(res[i])
    for k in range(3):
        for r in refs:
            r.setFunds(2000000)
            r.controlSimTick()
        b.setFunds(2000000)
        b.controlSimTick()
    _check(b, refs, "built")
    for fn, a, bb in [("GENERATE_TRAIN", 50, 48), ("GENERATE_PLANE", 52, 42), ("GENERATE_COPTER", 52, 42),
                      ("GENERATE_SHIP", 0, 0), ("MAKE_EXPLOSION", 45, 47), ("MAKE_MELTDOWN", 0, 0),
                      ("MAKE_EXPLOSION", 61, 41), ("GENERATE_TRAIN", 60, 48)]:
        sc = [r.scalars() for r in refs]
        for i, r in enumerate(refs):                             # generateTrain wants totalPop > 20, sprites a free slot
            sc[i].totalPop = max(sc[i].totalPop, 100)
            r.set_scalars(sc[i])
            b.set_scalars(i, sc[i])
        b.call(fn, a, bb)
        for r in refs:
            r.call(fn, a, bb)
        _check(b, refs, fn)
        b.simFrames(50)
        for r in refs:
            r.simFrames(50)
        _check(b, refs, fn + " + 50 frames")
    assert any(len(r.sprites()) > 0 for r in refs)
    rad = sum(int(((np.asarray(r.get("MPF_MAP")) & 1023) == 52).sum()) for r in refs)
    assert rad > 0, "the meltdown left no radiation"
    b.close()


def test_tools_every_tool_every_terrain_on_the_gpu():
    """toolDown results and map effects for every tool over water / woods / dirt / built tiles and at the world edges
    (prepareBuildingSite bounds, connectTile neighbours), 6000 clicks on each of 4 terrains through k_mp_tool."""
    n = 4
    b, refs = _mk(n)
    ss = np.arange(n) + 7
    b.generate(ss, ss + 3)
    for i, r in enumerate(refs):
        r.generate(int(ss[i]))
        r.seed(int(ss[i]) + 3)
    rng = np.random.default_rng(5)
    for k in range(6000):
        tools = rng.integers(0, 20, n)
        if rng.random() < 0.1:
            xs, ys = rng.choice([0, 1, 118, 119], n), rng.choice([0, 1, 98, 99], n)
        else:
            xs, ys = rng.integers(-1, 121, n), rng.integers(-1, 101, n)
        res = b.toolDown(tools, xs, ys)
        for i, r in enumerate(refs):
            assert r.toolDown(int(tools[i]), int(xs[i]), int(ys[i])) == int(res[i]), (k, i, tools[i], xs[i], ys[i])
        if k % 1000 == 999:
            _check(b, refs, "tools %d" % k)
    _check(b, refs, "tools final")
    b.close()


def test_speed_divider_frames():
    """simLoop's speed divider (main.cpp:403-425: at simSpeed 1 every 5th frame simulates, at 2 every 3rd, at 3 every
    one).  The gym path always runs at 3 and takes the frame kernel's fast path; envs at 1, 2 and 0 (paused) take the
    general one (sim_loop_head / sim_loop_tail) inside the same lockstep CTA."""
    n = 4
    b, refs = _mk(n)
    ss = np.arange(n) + 90
    b.generate(ss + 10, ss)
    for i, r in enumerate(refs):
        r.generate(int(ss[i]) + 10)
        r.seed(int(ss[i]))
        sc = r.scalars()
        sc.simSpeed = [1, 2, 3, 0][i]
        r.set_scalars(sc)
        b.set_scalars(i, sc)
    for k in range(5):
        b.simFrames(77)
        for r in refs:
            r.simFrames(77)
        _check(b, refs, "frames %d" % k)
    b.close()


def test_dense_city_through_the_gym_layer():
    """BASELINE config 4, dense variant, as bench.py runs it: a reference city loaded into the engine under a
    120x100 window, the TileMap shadow re-read from it (MicropolisControl.updateMap, corecontrol.py:193-203), random
    actions over all 20 tools -- observation, reward, done and engine state against the oracle gym layer."""
    import torch
    from gymcity_b200.micropolis_env import MicropolisEnv
    maps = np.load(os.path.join(GOLD, "cty_maps.npz"))["maps"]
    pick = [1, 7, 23]                                             # badnews, haight, yokohama: the densest
    n, size = len(pick), (120, 100)
    seeds = np.arange(n) + 70
    env = MicropolisEnv(num_envs=n)
    env.setMapSize(size, max_step=100, empty_start=False, MAP_XS=0, MAP_YS=0)
    refs = [RefEnv(RefEngine(), size, max_step=100, empty_start=False, xs=0, ys=0) for _ in range(n)]
    m = env.micro
    m.newMap(seeds + 1000, seeds)
    for i, r in enumerate(refs):
        r.reset_begin(int(seeds[i]) + 1000, int(seeds[i]))
        r.micro.engine.set("MPF_MAP", maps[pick[i]])
        r.micro.updateMap()
        m.engine.set(i, "MPF_MAP", maps[pick[i]])
    m.updateMap()
    m._ck(m._L.mpb_env_reset_finish(m._h, None, m.engine._stream()))
    obs = m.obs.cpu().numpy()
    for i, r in enumerate(refs):
        assert np.array_equal(r.reset_finish().astype(np.float32), obs[i]), i
    rng = np.random.default_rng(4)
    for t in range(5):
        acts = rng.integers(0, 20 * 12000, n).astype(np.int64)
        obs, rew, done, _ = env.step(torch.from_numpy(acts))
        obs, rew, done = obs.cpu().numpy(), rew.cpu().numpy().reshape(-1), done.cpu().numpy()
        for i, r in enumerate(refs):
            o, rr, d, _ = r.step(int(acts[i]), False)
            assert np.array_equal(np.asarray(o, np.float32), obs[i]), (t, i)
            assert rr == rew[i] and d == bool(done[i]), (t, i)
    for i, r in enumerate(refs):
        assert not diff_snapshots(r.micro.engine.snapshot(), m.engine.snapshot(i)), i
    env.close()
