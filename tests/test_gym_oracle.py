"""CPU: the gym layer.  (1) oracle/ref_gym.RefTileMap against the reference's own TileMap class
imported by path (only where /root/reference exists); (2) the device gym-layer code (CPU harness)
against the oracle gym layer over the reference engine: observation, reward, done, shadow map."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest

from oracle.ref_engine import RefEngine, diff_snapshots, ref_available
from oracle.ref_gym import RefControl, RefEnv, RefTileMap
from tests.host_engine import HostGym

pytestmark = pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built")
REF_TILEMAP = "/root/reference/gym_city/envs/tilemap.py"


def _reference_tilemap_cls():
    gym = types.ModuleType("gym")
    gym.spaces = types.ModuleType("gym.spaces")
    sys.modules.setdefault("gym", gym)
    sys.modules.setdefault("gym.spaces", gym.spaces)
    spec = importlib.util.spec_from_file_location("ref_tilemap", REF_TILEMAP)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)

    class Adapter(m.TileMap):
        def __init__(self, micro, X, Y):
            micro.env = types.SimpleNamespace(num_step=0)
            super().__init__(micro, X, Y, paint=False)
    return Adapter, m


@pytest.mark.skipif(not os.path.exists(REF_TILEMAP), reason="reference tree not present")
@pytest.mark.parametrize("seed,size", [(0, 16), (1, 12), (2, 24)])
def test_ref_tilemap_restatement_matches_reference_class(seed, size):
    Adapter, m = _reference_tilemap_cls()
    for i in range(1024):
        from oracle.ref_gym import zone_from_int
        assert zone_from_int(i) == m.zoneFromInt(i)
    engines = [RefEngine(), RefEngine()]
    ctl = [RefControl(engines[0], size, size, tilemap_cls=RefTileMap),
           RefControl(engines[1], size, size, tilemap_cls=Adapter)]
    for e in engines:
        e.clearMap()
        e.seed(seed)
        sc = e.scalars()
        sc.enableDisasters = 0          # floods make the reference class raise KeyError
        e.set_scalars(sc)
    for c in ctl:
        c.env = types.SimpleNamespace(num_step=0)
        c.clearMap()
        c.map.init_age_array()              # Extinguisher's age tracking (tilemap.py:201-206)
    rng = np.random.default_rng(seed)
    for t in range(400):
        a = [int(rng.integers(0, 20)), int(rng.integers(0, size)), int(rng.integers(0, size))]
        sb = bool(rng.random() < 0.25)
        for c in ctl:
            c.env.num_step = t
            c.takeAction(a, sb)
            c.engine.setFunds(2000000)
            if t % 3 == 0:
                c.simTick()
        a_, b_ = ctl[0].map, ctl[1].map
        assert np.array_equal(a_.zoneMap, b_.zoneMap), t
        assert np.array_equal(a_.static_builds, b_.static_builds), t
        assert np.array_equal(a_.age_order, b_.age_order), t
        assert (a_.num_plants, a_.num_roads, a_.n_struct_tiles) == (b_.num_plants, b_.num_roads, b_.n_struct_tiles)
        ca = [[a_.centers[x][y] for y in range(size)] for x in range(size)]
        cb = [[b_.centers[x][y] for y in range(size)] for x in range(size)]
        assert ca == cb, t
    assert not diff_snapshots(engines[0].snapshot(), engines[1].snapshot())


def _rollout(seed, steps, size, terrain, static_p, xs=16, ys=8):
    r = RefEngine()
    env = RefEnv(r, size, max_step=steps - 3, empty_start=not terrain, xs=xs, ys=ys)
    W, H = env.MAP_X, env.MAP_Y
    h = HostGym()
    h.configure(W, H, steps - 3, xs=xs, ys=ys)
    rng = np.random.default_rng(seed)
    env.reset_begin(1000 + seed, seed)
    h.reset_begin(terrain, 1000 + seed, seed)
    if not terrain:
        r.seed(seed)
        h.seed(seed)
    o0, o1 = env.reset_finish(), h.reset_finish()
    assert np.array_equal(o0.astype(np.float32), o1)
    done_seen = False
    for t in range(steps):
        a = int(rng.integers(0, env.num_tools * W * H)) if rng.random() > 0.05 else -1
        sb = bool(rng.random() < static_p)
        o0, r0, d0, _ = env.step(a, sb)
        o1, r1, d1 = h.step(a, sb)
        assert np.array_equal(o0.astype(np.float32), o1), (t, np.argwhere(o0.astype(np.float32) != o1)[:4])
        assert r0 == r1 and d0 == d1, (t, r0, r1, d0, d1)
        done_seen |= d0
        z, s, c, k = h.shadow()
        assert np.array_equal(z, env.micro.map.zoneMap[-1]) and np.array_equal(s, env.micro.map.static_builds[0])
        assert int(k[0]) == env.micro.map.num_plants and int(k[1]) == env.micro.map.num_roads
        assert int(k[3]) == env.micro.total_traffic and int(k[5]) == env.micro.map.tilemap_error
    assert not diff_snapshots(r.snapshot(), h.snapshot())
    assert done_seen                      # max_step reached: done = num_step >= max_step (env.py:517-520)
    return env


@pytest.mark.parametrize("seed,size,terrain,static_p", [(0, 16, True, 0.0), (1, 16, False, 0.3),
                                                       (2, 32, True, 0.1), (3, (20, 12), False, 0.2)])
def test_device_gym_layer_matches_oracle(seed, size, terrain, static_p):
    _rollout(seed, 120, size, terrain, static_p)


def test_full_world_window_at_origin():
    """BASELINE config 4 addressing: 120x100 window at offset (0,0)."""
    _rollout(5, 12, (120, 100), True, 0.0, xs=0, ys=0)


def test_poet_observation_planes():
    """MicropolisEnv(poet=True) (env.py:301-313): 38 planes, pops / ranges and targets / ranges."""
    seed, size = 7, 16
    r = RefEngine()
    env = RefEnv(r, size, max_step=10 ** 6, empty_start=False)
    env.poet = True
    env.city_trgs = {'res_pop': 320, 'com_pop': 77, 'ind_pop': 13, 'traffic': 900, 'num_plants': 3, 'mayor_rating': 61}
    h = HostGym()
    h.configure(size, size, 10 ** 6)
    h.set_poet(env.param_ranges, [env.city_trgs[k] for k in ('res_pop', 'com_pop', 'ind_pop', 'traffic', 'num_plants',
                                                              'mayor_rating')])
    env.reset_begin(1000 + seed, seed)
    h.reset_begin(True, 1000 + seed, seed)
    o0, o1 = env.reset_finish(), h.reset_finish()
    assert o0.shape == (38, size, size) and np.array_equal(o0.astype(np.float32), o1)
    rng = np.random.default_rng(seed)
    for t in range(60):
        a = int(rng.integers(0, 20 * size * size))
        o0, r0, d0, _ = env.step(a)
        o1, r1, d1 = h.step(a)
        assert np.array_equal(o0.astype(np.float32), o1), (t, np.argwhere(o0.astype(np.float32) != o1)[:4])
        assert r0 == r1
    assert float(o1[31, 0, 0]) == np.float32(320 / 750) and float(o1[36, 3, 3]) == np.float32(61 / 100)


# ---- the restated env / control arithmetic against the reference's own functions, lifted out of its files ------
REF_ENV = "/root/reference/gym_city/envs/env.py"
REF_CONTROL = "/root/reference/gym_city/envs/corecontrol.py"


def _lift(path, class_name, names):
    """The named methods of `class_name` in the reference file `path`, compiled on their own (the modules import gtk /
    the SWIG engine at the top and cannot be imported here) into a bare class."""
    import ast
    tree = ast.parse(open(path).read())
    cls = next(n for n in ast.walk(tree) if isinstance(n, ast.ClassDef) and n.name == class_name)
    fns = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert sorted(f.name for f in fns) == sorted(names), [f.name for f in fns]
    mod = ast.Module(body=[ast.ClassDef(name="Lifted", bases=[], keywords=[], body=fns, decorator_list=[])], type_ignores=[])
    ast.fix_missing_locations(mod)
    ns = {"np": np, "print": lambda *a, **k: None}
    exec(compile(mod, path, "exec"), ns)
    return ns["Lifted"]


@pytest.mark.skipif(not os.path.exists(REF_ENV), reason="reference tree not present")
def test_ref_env_reward_and_action_map_match_reference_functions():
    """RefEnv.getReward / decode against MicropolisEnv.getReward / mapIntsToActions (env.py:209-219, 345-387)."""
    L = _lift(REF_ENV, "MicropolisEnv", ["getReward", "mapIntsToActions"])
    rng = np.random.default_rng(0)
    env = RefEnv(RefEngine(), 16, max_step=10)
    ref = L()
    ref.city_trgs, ref.weights = dict(env.city_trgs), dict(env.weights)
    for k in range(4000):
        scale = [1, 10, 1000][k % 3]
        last = {m: int(rng.integers(-scale, 3 * scale)) for m in env.city_trgs}
        cur = {m: (last[m] if rng.random() < 0.3 else int(rng.integers(-scale, 3 * scale))) for m in env.city_trgs}
        if k % 7 == 0:                                       # exactly on / beyond the target
            m = list(env.city_trgs)[k % 6]
            cur[m] = env.city_trgs[m] + int(rng.integers(-1, 2))
        env.last_city_metrics = ref.last_city_metrics = last
        env.city_metrics = ref.city_metrics = cur
        assert env.getReward() == ref.getReward(), (last, cur)
    for T, W, H in [(20, 16, 16), (1, 32, 32), (20, 20, 12)]:
        ref = L()
        ref.num_tools, ref.MAP_X, ref.MAP_Y = T, W, H
        ref.intsToActions, ref.actionsToInts = {}, np.zeros((T, W, H), int)
        ref.mapIntsToActions()
        e = RefEnv(RefEngine(), (W, H), max_step=10, power_puzzle=(T == 1))
        assert e.num_tools == T
        for i in range(T * W * H):
            assert e.decode(i) == ref.intsToActions[i] and ref.actionsToInts[tuple(e.decode(i))] == i


@pytest.mark.skipif(not os.path.exists(REF_CONTROL), reason="reference tree not present")
@pytest.mark.parametrize("W,H,xs,ys", [(16, 16, 16, 8), (32, 32, 16, 8), (120, 100, 0, 0), (20, 12, 30, 40)])
def test_ref_control_density_maps_match_reference_functions(W, H, xs, ys):
    """RefControl.getDensityMaps against MicropolisControl.getDensityMaps / fillDensityMap (corecontrol.py:205-239)
    run over the same engine state: the swapped x / y offsets (MAP_YS added to the row index), the half-resolution
    lookups and total_traffic."""
    L = _lift(REF_CONTROL, "MicropolisControl", ["fillDensityMap", "getDensityMaps"])
    r = RefEngine()
    r.generate(77)
    r.seed(7)
    rng = np.random.default_rng(W + H)
    r.set("MPF_TRAFFIC_DENSITY", rng.integers(0, 256, 3000).astype(np.uint8))
    r.set("MPF_POP_DENSITY", rng.integers(0, 256, 3000).astype(np.uint8))
    pw = (rng.random(12000) < 0.3).astype(np.uint8)
    r.set("MPF_POWER_GRID", pw)
    traffic = np.asarray(r.get("MPF_TRAFFIC_DENSITY")).reshape(60, 50)
    pop = np.asarray(r.get("MPF_POP_DENSITY")).reshape(60, 50)
    power = np.asarray(r.get("MPF_POWER_GRID")).reshape(120, 100)

    class Eng:                                                 # stubs.cpp:382-505: Map::get / worldGet, 0 off the map
        @staticmethod
        def getTrafficDensity(x, y): return int(traffic[x, y]) if 0 <= x < 60 and 0 <= y < 50 else 0
        @staticmethod
        def getPopulationDensity(x, y): return int(pop[x, y]) if 0 <= x < 60 and 0 <= y < 50 else 0
        @staticmethod
        def getPowerGrid(x, y): return int(power[x, y]) if 0 <= x < 120 and 0 <= y < 100 else 0
    ref = L()
    ref.engine, ref.MAP_X, ref.MAP_Y, ref.MAP_XS, ref.MAP_YS = Eng, W, H, xs, ys
    want = ref.getDensityMaps()
    ctl = RefControl(r, W, H, xs=xs, ys=ys)
    got = ctl.getDensityMaps()
    assert np.array_equal(want, got)
    assert ctl.total_traffic == ref.total_traffic
