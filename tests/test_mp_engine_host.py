"""CPU: the engine headers the CUDA kernels are built from (compiled for one host thread by
tests/host_engine.py) against the reference engine compiled as the oracle -- bit-exact maps, scalars
and sprites.  The GPU suite (test_mp_engine_gpu.py) repeats this through the mpb_* C-ABI."""
import numpy as np
import pytest

from oracle.ref_engine import RefEngine, diff_snapshots, ref_available
from tests.host_engine import HostEngine

pytestmark = pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built")


def _pair():
    return RefEngine(), HostEngine()


def _same(r, h, tag):
    d = diff_snapshots(r.snapshot(), h.snapshot())
    assert not d, (tag, d[:6])


def test_construct_matches_reference_constructor_sequence():
    r, h = _pair()
    _same(r, h, "construct")


@pytest.mark.parametrize("seed", range(8))
def test_terrain_generation(seed):
    r, h = _pair()
    for e in (r, h):
        e.generate(100 + seed)
        e.seed(seed)
    _same(r, h, "terrain")


@pytest.mark.parametrize("seed,terrain", [(0, True), (1, True), (2, False)])
def test_random_action_rollout_400_env_ticks(seed, terrain):
    r, h = _pair()
    rng = np.random.default_rng(seed)
    for e in (r, h):
        if terrain:
            e.generate(1000 + seed)
        else:
            e.clearMap()
        e.seed(seed)
    for t in range(400):
        tool = int(rng.integers(0, 20))
        tool = 7 if tool == 5 else tool
        x, y = int(rng.integers(0, 16)) + 16, int(rng.integers(0, 16)) + 8
        assert r.toolDown(tool, x, y) == h.toolDown(tool, x, y), t
        for e in (r, h):
            e.setFunds(2000000)
            e.controlSimTick()
        if t % 20 == 19:
            _same(r, h, "tick %d" % t)
    _same(r, h, "final")


def _city_actions(rng, x0=20, y0=10, x1=100, y1=88):
    R, C, I, FIRE, POL, WIRE, RAIL, ROAD, STAD, PARK, PORT, COAL, NUKE, AIR = 0, 1, 2, 3, 4, 6, 8, 9, 10, 11, 12, 13, 14, 15
    acts = []
    roads = list(range(y0, y1, 7))
    for ry in roads:
        acts += [(ROAD, x, ry) for x in range(x0, x1)]
    for ry in roads[:-1]:
        for zy in (ry + 2, ry + 5):
            for x in range(x0 + 1, x1 - 2, 3):
                p = rng.random()
                t = R if p < 0.5 else (C if p < 0.72 else (I if p < 0.92 else (POL if p < 0.94 else (FIRE if p < 0.96 else PARK))))
                acts.append((t, x, zy))
    for ry in roads:
        acts += [(WIRE, x0 + 5, ry), (WIRE, x0 + 40, ry)]
    acts += [(COAL, x0 + 2, y0 - 3), (NUKE, x0 + 8, y0 - 3), (COAL, x0 + 14, y0 - 3)]
    acts += [(WIRE, x, y0 - 1) for x in range(x0, x0 + 20)]
    acts += [(RAIL, x, y1 + 2) for x in range(x0 - 2, x1 + 2)]
    acts += [(STAD, x1 + 4, y0 + 10), (AIR, x1 + 6, y0 + 20), (PORT, x1 + 4, y0 + 34)]
    acts += [(WIRE, x1 + 1, y) for y in range(y0, y0 + 40)]
    return acts


@pytest.mark.parametrize("seed,disasters,terrain", [(0, True, False), (1, False, False), (2, True, True)])
def test_built_city_rollout(seed, disasters, terrain):
    """Zone growth, traffic walks, power-stack walk, hospitals/churches, stadium, airport, seaport,
    trains / planes / copters / monsters / tornadoes / explosions, fires, floods, earthquakes."""
    r, h = _pair()
    rng = np.random.default_rng(seed)
    for e in (r, h):
        e.generate(1000 + seed)
        if not terrain:
            e.clearMap()
        e.seed(seed)
        if not disasters:
            sc = e.scalars()
            sc.enableDisasters = 0
            e.set_scalars(sc)
    acts, ai, seen = _city_actions(rng), 0, set()
    for t in range(260):
        for k in range(40):
            if ai < len(acts):
                tool, x, y = acts[ai]
                ai += 1
            elif k < 1 and rng.random() < 0.15:
                tool, x, y = int(rng.integers(0, 20)), int(rng.integers(15, 110)), int(rng.integers(5, 95))
                tool = 7 if tool == 5 else tool
            else:
                break
            assert r.toolDown(tool, x, y) == h.toolDown(tool, x, y)
        for e in (r, h):
            e.setFunds(2000000)
            e.controlSimTick()
        seen |= set(s[0] for s in r.sprites())
        if t % 10 == 9:
            _same(r, h, "tick %d" % t)
    _same(r, h, "final")
    assert seen, "scenario produced no sprites"


def test_single_functions_on_injected_state():
    r, h = _pair()
    for e in (r, h):
        e.generate(45)
        e.seed(40)
    rng = np.random.default_rng(2)
    for f in ["MPF_TRAFFIC_DENSITY", "MPF_POP_DENSITY", "MPF_LAND_VALUE", "MPF_POLLUTION_DENSITY", "MPF_CRIME_RATE"]:
        a = rng.integers(0, 256, 3000).astype(np.uint8)
        r.set(f, a)
        h.set(f, a)
    for f in ["MPF_RATE_OF_GROWTH", "MPF_FIRE_STATION", "MPF_POLICE_STATION"]:
        a = rng.integers(-250, 1200, 195).astype(np.int16)
        r.set(f, a)
        h.set(f, a)
    for fn in ["DEC_TRAFFIC", "DEC_ROG", "FIRE_ANALYSIS", "CRIME_SCAN", "POP_DENSITY_SCAN", "PTLV_SCAN",
               "SET_VALVES", "TAKE10", "TAKE120", "COLLECT_TAX", "CITY_EVAL", "SEND_MESSAGES", "ANIMATE",
               "MAKE_EARTHQUAKE", "MAKE_FLOOD", "SET_FIRE", "MAKE_TORNADO", "MAKE_MONSTER", "MOVE_OBJECTS",
               "DO_DISASTERS", "POWER_SCAN"]:
        r.call(fn)
        h.call(fn)
        _same(r, h, fn)


def test_tools_every_tool_every_terrain():
    """toolDown results and map effects for every tool over water / woods / dirt / built tiles and
    at the world edges (prepareBuildingSite bounds, connectTile neighbours)."""
    r, h = _pair()
    for e in (r, h):
        e.generate(7)
        e.seed(3)
    rng = np.random.default_rng(5)
    for k in range(6000):
        tool = int(rng.integers(0, 20))
        if rng.random() < 0.1:
            x, y = int(rng.choice([0, 1, 118, 119])), int(rng.choice([0, 1, 98, 99]))
        else:
            x, y = int(rng.integers(-1, 121)), int(rng.integers(-1, 101))
        assert r.toolDown(tool, x, y) == h.toolDown(tool, x, y), (k, tool, x, y)
        if k % 500 == 499:
            _same(r, h, "tools %d" % k)
    _same(r, h, "tools final")


def test_scan_period_reciprocal_is_exact():
    """mp_sim.h scan_due replaces `simCycle % period == 0` (simulate.cpp:124-135, periods 1..20, simCycle 0..1023) by a
    multiply with ceil(2^16 / period) and a shift; the identity is checked here over the whole domain."""
    for d in range(1, 21):
        r = (65535 + d) // d
        for c in range(1024):
            assert (c - ((c * r) >> 16) * d == 0) == (c % d == 0), (c, d)
