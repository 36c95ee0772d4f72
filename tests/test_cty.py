""".cty wire format (fileio.cpp:160-330): the parser against the reference engine's own loadFileDir on the
city files the reference ships as fixtures (committed copies are not needed: the parser is also checked on a
synthetic file), the writer through a round trip and through the reference loader; on the GPU, env state ->
file -> reference engine and file -> env."""
import os

import numpy as np
import pytest

from gymcity_b200 import cty
from oracle.ref_engine import RefEngine, ref_available

CITIES = "/root/reference/gym_city/envs/micropolis/MicropolisCore/cities"


def _synthetic(seed):
    rng = np.random.default_rng(seed)
    f = {n: rng.integers(-3000, 3000, 240).astype(np.int16) for n in cty.HIST}
    f["MPF_MISC_HIST"] = rng.integers(-100, 100, 120).astype(np.int16)
    f["MPF_MAP"] = rng.integers(0, 65535, 12000).astype(np.uint16)
    return f


def test_round_trip_and_scalar_words(tmp_path):
    f = _synthetic(0)
    cty._put32(f["MPF_MISC_HIST"], 50, 1234567)
    cty._put32(f["MPF_MISC_HIST"], 8, 4800)
    f["MPF_MISC_HIST"][52:58] = [1, 0, 1, 0, 9, 3]
    cty._put32(f["MPF_MISC_HIST"], 58, 65536)
    cty._put32(f["MPF_MISC_HIST"], 60, 32768)
    cty._put32(f["MPF_MISC_HIST"], 62, 0)
    p = tmp_path / "a.cty"
    cty.write_cty(p, f)
    assert os.path.getsize(p) == cty.FILE_BYTES
    g = cty.read_cty(p)
    for k in f:
        assert np.array_equal(np.asarray(f[k]).view(np.uint16), np.asarray(g[k]).view(np.uint16)), k
    m = cty.decode_misc(g["MPF_MISC_HIST"])
    assert (m["totalFunds"], m["cityTime"], m["cityTax"], m["simSpeed"], m["autoBulldoze"], m["autoBudget"]) == \
        (1234567, 4800, 9, 3, 1, 0)
    assert (m["policePercent"], m["firePercent"], m["roadPercent"]) == (1.0, 0.5, 0.0)
    raw = open(p, "rb").read()
    assert raw[2 * (1440 + 50): 2 * (1440 + 52)] == (1234567).to_bytes(4, "big")      # 32-bit big-endian in the file


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built")
def test_writer_is_read_by_the_reference_loader(tmp_path):
    f = _synthetic(1)
    p = tmp_path / "b.cty"
    cty.write_cty(p, f)
    r = RefEngine()
    assert r.loadFileDir(p)
    s = r.snapshot()
    for k in cty.HIST + ("MPF_MISC_HIST", "MPF_MAP"):
        assert np.array_equal(np.asarray(s[k]).view(np.uint16), np.asarray(f[k]).view(np.uint16)), k


@pytest.mark.skipif(not (ref_available() and os.path.isdir(CITIES)), reason="reference city files not present")
def test_parser_matches_reference_loader_on_shipped_cities():
    names = sorted(n for n in os.listdir(CITIES) if n.endswith(".cty"))
    assert len(names) >= 20
    for n in names:
        r = RefEngine()
        assert r.loadFileDir(os.path.join(CITIES, n)), n
        s, g = r.snapshot(), cty.read_cty(os.path.join(CITIES, n))
        for k in cty.HIST + ("MPF_MISC_HIST", "MPF_MAP"):
            assert np.array_equal(np.asarray(s[k]).view(np.uint16), np.asarray(g[k]).view(np.uint16)), (n, k)
        m = cty.decode_misc(g["MPF_MISC_HIST"])
        assert m["cityTime"] >= 0 and -(2 ** 31) <= m["totalFunds"] < 2 ** 31, n


@pytest.mark.gpu
@pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built")
def test_env_state_to_file_to_reference_and_back(tmp_path):
    import torch
    from gymcity_b200.micropolis_env import MicropolisEnv
    n = 3
    env = MicropolisEnv(num_envs=n)
    env.seed(2)
    env.setMapSize(16, max_step=10 ** 6, empty_start=False)
    env.reset()
    rng = np.random.default_rng(0)
    for t in range(12):
        env.step(torch.from_numpy(rng.integers(0, 20 * 256, n)))
    eng = env.micro.engine
    p = tmp_path / "env1.cty"
    cty.save_env(eng, 1, p)
    r = RefEngine()
    assert r.loadFileDir(p)
    s, mine = r.snapshot(), eng.snapshot(1)
    for k in cty.HIST + ("MPF_MAP",):
        assert np.array_equal(np.asarray(s[k]).view(np.uint16), np.asarray(mine[k]).view(np.uint16)), k
    m = cty.decode_misc(cty.read_cty(p)["MPF_MISC_HIST"])
    sc = eng.scalars(1)
    assert (m["totalFunds"], m["cityTime"], m["cityTax"]) == (sc.totalFunds, sc.cityTime, sc.cityTax)
    # and back into another env of the batch: map, histories and the decoded scalars arrive; the derived
    # bitmaps follow the map (a tick afterwards must not trip over stale ones)
    cty.load_env(eng, 2, p)
    a, b = eng.snapshot(1), eng.snapshot(2)
    for k in cty.HIST + ("MPF_MAP",):
        assert np.array_equal(a[k], b[k]), k
    assert eng.scalars(2).totalFunds == sc.totalFunds and eng.scalars(2).cityTime == sc.cityTime
    env.step(torch.from_numpy(rng.integers(0, 20 * 256, n)))
    env.close()
