"""BASELINE.json's batch sizes on the GPU.  The small parity tests run one CTA of a handful of envs; the
scheduler (heavy-first permutation, heavy env group, 32-env CTAs in frame lockstep, env groups on
streams) only comes into play with thousands of envs.  Property used: an env's trajectory depends on its
global index, its seeds and its actions only -- so env i of a full-size batch must equal the same env
stepped in a two-env batch at env_offset = i, bit for bit, and (terrain path) the reference engine driven
by the oracle gym layer with the same seeds."""
import numpy as np
import pytest

from oracle.ref_engine import RefEngine, diff_snapshots, ref_available
from oracle.ref_gym import RefEnv

pytestmark = pytest.mark.gpu


def _probe_indices(n):
    return sorted({0, 1, 31, 32, 33, n // 64 - 1, n // 64, n // 3 + 5, n // 2, n - 2})


@pytest.mark.parametrize("size,n,xs,ys,steps", [(16, 32768, 16, 8, 45), ((120, 100), 2048, 0, 0, 8)])
def test_full_batch_envs_equal_small_batches_and_reference(size, n, xs, ys, steps):
    import torch
    from gymcity_b200.micropolis_env import MicropolisVecEnv
    from gymcity_b200.sharding import mix
    kw = dict(seed=5, max_step=10 ** 6, empty_start=False, MAP_XS=xs, MAP_YS=ys)
    big = MicropolisVecEnv(n, size, **kw)
    W, H = big.MAP_X, big.MAP_Y
    idx = _probe_indices(n)
    small = [MicropolisVecEnv(2, size, env_offset=i, **kw) for i in idx]
    refs = []
    if ref_available():
        for i in idx[:3]:
            r = RefEnv(RefEngine(), size, max_step=10 ** 6, empty_start=False, xs=xs, ys=ys)
            r.reset_begin(int(np.int32(mix(5, i, 0, 0))), int(np.int32(mix(5, i, 0, 1))))
            refs.append((i, r, r.reset_finish()))
    ob = big.reset()
    for k, i in enumerate(idx):
        assert torch.equal(ob[i], small[k].reset()[0]), i
    for i, r, o0 in refs:
        assert np.array_equal(o0.astype(np.float32), ob[i].cpu().numpy()), i
    rng = np.random.default_rng(9)
    nA = 20 * W * H
    for t in range(steps):
        a = rng.integers(0, nA, n)
        ob, rb, db, _ = big.step(torch.from_numpy(a))
        for k, i in enumerate(idx):
            os_, rs, ds, _ = small[k].step(torch.from_numpy(a[i:i + 2]))
            assert torch.equal(ob[i], os_[0]) and torch.equal(rb[i], rs[0]) and torch.equal(db[i], ds[0]), (t, i)
        for i, r, _ in refs:
            o, rr, d, _ = r.step(int(a[i]))
            assert np.array_equal(o.astype(np.float32), ob[i].cpu().numpy()) and rr == float(rb[i, 0].item()), (t, i)
    for k, i in enumerate(idx):
        assert not diff_snapshots(big.micro.engine.snapshot(i), small[k].micro.engine.snapshot(0)), i
    for i, r, _ in refs:
        assert not diff_snapshots(r.micro.engine.snapshot(), big.micro.engine.snapshot(i)), i
    # every env advanced by the same number of steps
    ctr = big.micro.counters()
    assert int(ctr[:, 4].min()) == steps == int(ctr[:, 4].max())
    big.close()
    for s in small:
        s.close()


def test_power_puzzle_8192_envs_is_deterministic_and_well_formed():
    """BASELINE config 3: 32x32, 8192 envs, Wire-only actions over 5 static Residential + 1 static Nuclear."""
    import torch
    from gymcity_b200.micropolis_env import MicropolisVecEnv
    n, size = 8192, 32
    out = []
    for rep in range(2):
        env = MicropolisVecEnv(n, size, seed=11, max_step=10 ** 6, empty_start=False, power_puzzle=True)
        assert env.action_space.n == size * size
        env.reset()
        rng = np.random.default_rng(4)
        for t in range(4):
            obs, rew, done, _ = env.step(torch.from_numpy(rng.integers(0, size * size, n)))
        ctr = env.micro.counters()
        out.append((obs.sum(dim=(2, 3)).cpu().numpy(), rew.cpu().numpy().copy(), ctr.copy()))
        static_cells = obs[:, 31].sum(dim=(1, 2)).cpu().numpy()
        # static builds never shrink below one zone and never exceed 5 x 9 + 16 cells
        assert static_cells.min() >= 9 and static_cells.max() <= 5 * 9 + 16
        assert int(ctr[:, 0].max()) <= 16                                        # num_plants counts cells: one 4x4 plant
        env.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    assert np.array_equal(out[0][2], out[1][2])


@pytest.mark.parametrize("n,steps", [(32768, 45), (18950, 12), (2047, 12), (8197, 12)])
def test_results_do_not_depend_on_the_schedule(n, steps, monkeypatch):
    """Every env of a 32768-env batch, 45 aged steps (flood fields, sprites, the dense scan path and the
    heavy group are all live by then): the default schedule (3 env groups on streams, heavy-first order, 32
    envs per CTA in frame lockstep) against the plainest one (one group, index order, one free-running warp
    per CTA) -- observations, rewards, dones and bookkeeping counters of ALL envs must be identical."""
    import torch
    from gymcity_b200.micropolis_env import MicropolisVecEnv
    # 32 768: the headline batch; 18 950: just above the one-group threshold, last CTA partial; 2 047 / 8 197: three groups
    # of 8- / 16-env chunks with a partial last chunk and a heavy group that is not a multiple of the chunk

    def run(envvars):
        for k, v in envvars.items():
            monkeypatch.setenv(k, v)
        env = MicropolisVecEnv(n, 16, seed=21, max_step=10 ** 6, empty_start=False)
        env.reset()
        rng = np.random.default_rng(3)
        rsum = torch.zeros(n, 1, device="cuda")
        snaps = []
        for t in range(steps):
            obs, rew, done, _ = env.step(torch.from_numpy(rng.integers(0, 20 * 256, n)))
            rsum += rew
            if t % 15 == 14 or t == steps - 1:
                snaps.append((obs.clone(), done.clone()))
        ctr = env.micro.counters()
        maps = [env.micro.engine.get(i, "MPF_MAP").copy() for i in (0, 777, n - 1)]
        assert int(ctr[:, 4].min()) == steps == int(ctr[:, 4].max())          # every env stepped exactly `steps` times
        env.close()
        for k in envvars:
            monkeypatch.delenv(k)
        return snaps, rsum, ctr, maps

    a = run({})
    # ... and the action kernel with a warp per env (k_mp_env_pre) instead of four envs to a warp (k_mp_env_pre_t)
    b = run({"MPB_GROUPS": "1", "MPB_FEPC": "1", "MPB_HEAVY_FEPC": "1", "MPB_SCHED": "0", "MPB_PRE_SIMT": "0"})
    for (oa, da), (ob, db) in zip(a[0], b[0]):
        assert torch.equal(oa, ob) and torch.equal(da, db)
    assert torch.equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    for ma, mb in zip(a[3], b[3]):
        assert np.array_equal(ma, mb)
