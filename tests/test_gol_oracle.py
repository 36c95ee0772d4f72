"""CPU: the GoL oracle (oracle/gol_oracle.py) against golden vectors produced by the reference's
own worlds (tests/golden/make_gol_golden.py), plus known-answer patterns."""
import os

import numpy as np
import pytest

from oracle import gol_oracle as go

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "gol_golden.npz"))


@pytest.mark.parametrize("key", ["multi_8_16", "multi_4_5", "multi_3_32", "multi_5_8"])
def test_multi_oracle_matches_reference_world_pytorch(key):
    st = G[key + "_init"].copy()
    acts, states, pops = G[key + "_acts"], G[key + "_states"], G[key + "_pops"]
    n, w = st.shape[0], st.shape[1]
    for t in range(acts.shape[0]):
        st, rew, pop = go.multi_step(st, acts[t], w * w * 2 / 3, w * w * 2 / 3, 200, n)
        assert np.array_equal(st, states[t]), (key, t)
        assert np.array_equal(pop, pops[t])


@pytest.mark.parametrize("key", ["classic_16", "classic_6"])
def test_classic_oracle_matches_reference_world(key):
    init, acts = G[key + "_init"], G[key + "_acts"]
    world = go.ClassicWorld(init)
    for t in range(acts.shape[0]):
        s, r, term = go.classic_step(world, int(acts[t]), 100, t)
        assert np.array_equal(s, G[key + "_states"][t]), (key, t)
        assert abs(r - G[key + "_raw"][t] * 100 / (100 * init.shape[0] ** 2)) < 1e-12


def test_known_patterns():
    b = np.zeros((8, 8), np.uint8)
    b[3, 2:5] = 1                                   # blinker
    b1 = go.tick_torus(b)
    assert b1[2:5, 3].sum() == 3 and b1.sum() == 3
    assert np.array_equal(go.tick_torus(b1), b)
    blk = np.zeros((6, 6), np.uint8)
    blk[0, 0] = blk[0, 5] = blk[5, 0] = blk[5, 5] = 1   # block across the torus seam
    assert np.array_equal(go.tick_torus(blk), blk)
    g = np.zeros((8, 8), np.uint8)
    g[0, 1] = g[1, 2] = g[2, 0] = g[2, 1] = g[2, 2] = 1  # glider: period 4, shift (1,1)
    s = g
    for _ in range(4):
        s = go.tick_torus(s)
    assert np.array_equal(s, np.roll(np.roll(g, 1, 0), 1, 1))


def test_torch_world_port_matches_bit_oracle():
    """The torch-CPU restatement of world_pytorch.World used as bench.py's Game-of-Life CPU baseline."""
    rng = np.random.default_rng(3)
    n, w = 64, 16
    st = (rng.random((n, w, w)) < 0.3).astype(np.uint8)
    tw = go.TorchWorldPort(st)
    for t in range(20):
        a = rng.integers(0, w * w, n)
        st, rew, pop = go.multi_step(st, a, w * w * 2 / 3, w * w * 2 / 3, 200, n)
        s2, r2, p2 = tw.step(a, w * w * 2 / 3, w * w * 2 / 3, 200)
        assert np.array_equal(s2[:, 0].numpy().astype(np.uint8), st), t
        assert np.array_equal(p2.numpy().astype(np.int64), pop)
        assert np.allclose(r2.numpy(), rew, rtol=1e-6)
