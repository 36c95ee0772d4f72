"""GPU parity: the CUDA Micropolis engine (mpb_* C-ABI) against the reference engine compiled as
the oracle (oracle/_ref/libmicropolis_ref.so), bit-exact on every state array, scalar and sprite."""
import numpy as np
import pytest

from oracle.ref_engine import RefEngine, diff_snapshots, ref_available

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built")]


def _mk(n):
    from gymcity_b200.engine import EngineBatch
    return EngineBatch(n), [RefEngine() for _ in range(n)]


def _check(b, refs, tag, every=None):
    for i, r in enumerate(refs):
        d = diff_snapshots(r.snapshot(), b.snapshot(i))
        assert not d, (tag, i, d[:6])


def test_construct_and_terrain():
    n = 12
    b, refs = _mk(n)
    _check(b, refs, "construct")
    ts, ss = np.arange(n) + 100, np.arange(n)
    b.generate(ts, ss)
    for i, r in enumerate(refs):
        r.generate(int(ts[i]))
        r.seed(int(ss[i]))
    _check(b, refs, "generate")
    b.close()


def test_terrain_generator_256_seeds():
    """generateSomeCity on all 32 lanes (generate_map_coop: plops a lane per cell, smoothRiver by word with ranked coin
    flips, tree splashes 32 steps at a time, smoothTrees as a wavefront) against the reference generator: map, RNG
    state and everything doSimInit derives from them.  256 seeds cover islands (1 in 10), walks along the map border
    and rejected draws (about 1.5 per map)."""
    n = 256
    b, refs = _mk(n)
    ts, ss = np.arange(n) * 7919 + 1, np.arange(n)
    b.generate(ts, ss)
    for i, r in enumerate(refs):
        r.generate(int(ts[i]))
        r.seed(int(ss[i]))
    _check(b, refs, "generate256")
    b.close()


@pytest.mark.parametrize("terrain,steps,n", [(True, 120, 8), (False, 60, 4)])
def test_random_action_rollout(terrain, steps, n):
    """BASELINE config 1 shape: 16x16 window at (16,8), uniform random tools, env-ticks of
    2 x 100 simFrames + animateTiles."""
    b, refs = _mk(n)
    ss = np.arange(n) + (7 if terrain else 3)
    if terrain:
        b.generate(ss + 1000, ss)
    else:
        b.clearMap()
        b.seed(ss)
    for i, r in enumerate(refs):
        if terrain:
            r.generate(int(ss[i]) + 1000)
        else:
            r.clearMap()
        r.seed(int(ss[i]))
    rng = np.random.default_rng(11)
    for t in range(steps):
        tools = rng.integers(0, 20, n)
        tools[tools == 5] = 7
        xs, ys = rng.integers(0, 16, n) + 16, rng.integers(0, 16, n) + 8
        res = b.toolDown(tools, xs, ys)
        for i, r in enumerate(refs):
            assert r.toolDown(int(tools[i]), int(xs[i]), int(ys[i])) == int(res[i]), (t, i)
            r.setFunds(2000000)
            r.controlSimTick()
        b.setFunds(2000000)
        b.controlSimTick()
        if t % 10 == 9 or t < 3:
            _check(b, refs, "step %d" % t)
    _check(b, refs, "final")
    b.close()


def _city_actions(rng, x0=20, y0=10, x1=100, y1=88):
    R, C, I, FIRE, POL, WIRE, RAIL, ROAD, STAD, PARK, PORT, COAL, NUKE, AIR = 0, 1, 2, 3, 4, 6, 8, 9, 10, 11, 12, 13, 14, 15
    acts = []
    roads = list(range(y0, y1, 7))
    for ry in roads:
        for x in range(x0, x1):
            acts.append((ROAD, x, ry))
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


@pytest.mark.parametrize("disasters", [True, False])
def test_built_city_rollout(disasters):
    """A laid-out 80x78 city (roads, R/C/I, stations, plants, rail, stadium, airport, seaport):
    exercises zone growth, traffic, power walk, sprites and disasters."""
    n = 3
    b, refs = _mk(n)
    ss = np.arange(n)
    b.clearMap()
    b.seed(ss)
    if not disasters:
        b.setEnableDisasters(False)
    for i, r in enumerate(refs):
        r.clearMap()
        r.seed(int(ss[i]))
        if not disasters:
            sc = r.scalars()
            sc.enableDisasters = 0
            r.set_scalars(sc)
    acts = [_city_actions(np.random.default_rng(i)) for i in range(n)]
    ai = 0
    alen = max(len(a) for a in acts)
    steps = alen // 40 + 60
    for t in range(steps):
        for k in range(40):
            if ai >= alen:
                break
            cur = [a[ai] if ai < len(a) else (-1, 0, 0) for a in acts]
            ai += 1
            res = b.toolDown([c[0] for c in cur], [c[1] for c in cur], [c[2] for c in cur])
            for i, r in enumerate(refs):
                if cur[i][0] >= 0:
                    assert r.toolDown(*cur[i]) == int(res[i])
        for r in refs:
            r.setFunds(2000000)
            r.controlSimTick()
        b.setFunds(2000000)
        b.controlSimTick()
        if t % 15 == 14:
            _check(b, refs, "step %d" % t)
    _check(b, refs, "final")
    assert max(r.scalars().totalPop for r in refs) >= 0
    b.close()


def test_single_functions_injected_state():
    """Run one engine function on identical injected state and diff."""
    n = 2
    b, refs = _mk(n)
    ss = np.arange(n) + 40
    b.generate(ss + 5, ss)
    for i, r in enumerate(refs):
        r.generate(int(ss[i]) + 5)
        r.seed(int(ss[i]))
    rng = np.random.default_rng(2)
    for i, r in enumerate(refs):
        for f, hi in [("MPF_TRAFFIC_DENSITY", 256), ("MPF_POP_DENSITY", 256), ("MPF_LAND_VALUE", 256),
                      ("MPF_POLLUTION_DENSITY", 256), ("MPF_CRIME_RATE", 256)]:
            a = rng.integers(0, hi, 3000).astype(np.uint8)
            r.set(f, a)
            b.set(i, f, a)
        for f in ["MPF_RATE_OF_GROWTH", "MPF_FIRE_STATION", "MPF_POLICE_STATION"]:
            a = rng.integers(-250, 1200, 195).astype(np.int16)
            r.set(f, a)
            b.set(i, f, a)
    for fn in ["DEC_TRAFFIC", "DEC_ROG", "FIRE_ANALYSIS", "CRIME_SCAN", "POP_DENSITY_SCAN", "PTLV_SCAN",
               "SET_VALVES", "TAKE10", "TAKE120", "COLLECT_TAX", "CITY_EVAL", "SEND_MESSAGES", "ANIMATE",
               "MAKE_EARTHQUAKE", "MAKE_FLOOD", "SET_FIRE", "MAKE_TORNADO", "MAKE_MONSTER"]:
        b.call(fn)
        for r in refs:
            r.call(fn)
        _check(b, refs, fn)
    b.close()


def test_power_scan_paths_match_reference():
    """doPowerScan's three device paths -- bitmap flood fill (large grid, ample capacity), literal walk
    (small grid, or capacity too close: one coal plant feeding > 160 conductive tiles) and the memo (nothing
    conductive changed between scans) -- against the reference engine: power grid, PWRBIT of every tile and
    the zone power census after full simulation cycles."""
    TOOL_WIRE, TOOL_RES, TOOL_COAL, TOOL_NUKE, TOOL_DOZE = 6, 0, 13, 14, 7
    layouts = []
    # 0: nuclear plant + a long serpentine of wire with zones hanging off it (fill path)
    a = [(TOOL_NUKE, 20, 20)]
    for row in range(6):
        y = 26 + 4 * row
        a += [(TOOL_WIRE, x, y) for x in range(18, 60)]
        a += [(TOOL_WIRE, 59 if row % 2 == 0 else 18, y + k) for k in range(1, 4)]
        a += [(TOOL_RES, x, y + 2) for x in range(22, 56, 8)]
    a += [(TOOL_WIRE, 20, y) for y in range(23, 27)]
    layouts.append(a)
    # 1: coal plant + ~200 wire tiles: 4 T + seeds + 256 >= 700, the literal walk (and its cut-off margin)
    a = [(TOOL_COAL, 30, 30)]
    a += [(TOOL_WIRE, x, 34) for x in range(10, 110)] + [(TOOL_WIRE, x, 36) for x in range(10, 110)]
    a += [(TOOL_WIRE, 10, 35), (TOOL_WIRE, 30, 33), (TOOL_WIRE, 30, 32)]
    layouts.append(a)
    # 2: small grid (below the fill threshold)
    layouts.append([(TOOL_COAL, 40, 40)] + [(TOOL_WIRE, x, 44) for x in range(38, 50)] + [(TOOL_RES, 50, 46)])
    n = len(layouts)
    b, refs = _mk(n)
    # a live simulation (generateSomeCity; a clearMap-only engine does not tick) on bare ground
    b.generate(np.arange(n) + 50, np.arange(n) + 3)
    bare = np.zeros(12000, np.uint16)
    for i, r in enumerate(refs):
        r.generate(i + 50)
        r.seed(i + 3)
        r.set("MPF_MAP", bare)
        b.set(i, "MPF_MAP", bare)
    for k in range(max(len(a) for a in layouts)):
        t = np.array([a[k][0] if k < len(a) else -1 for a in layouts])
        x = np.array([a[k][1] if k < len(a) else 0 for a in layouts])
        y = np.array([a[k][2] if k < len(a) else 0 for a in layouts])
        b.toolDown(t, x, y)
        for i, r in enumerate(refs):
            if t[i] >= 0:
                r.toolDown(int(t[i]), int(x[i]), int(y[i]))
    for rep in range(4):                     # 4 x 200 frames = 50 simulation cycles: 10 power scans, most of them memo hits
        b.setFunds(2000000)
        b.controlSimTick()
        for r in refs:
            r.setFunds(2000000)
            r.controlSimTick()
        _check(b, refs, "power rep %d" % rep)
        if rep == 1:                         # cut the grid in the middle: the memo must notice
            b.toolDown(np.full(n, TOOL_DOZE), np.array([40, 60, 44]), np.array([26, 34, 44]))
            for r, (x, y) in zip(refs, [(40, 26), (60, 34), (44, 44)]):
                r.toolDown(TOOL_DOZE, x, y)
    pw = [int(b.get(i, "MPF_POWER_GRID").sum()) for i in range(n)]
    assert pw[0] > 150 and pw[1] > 50 and pw[2] > 10, pw
    b.close()
