"""Vec-env adapter + rollout storage (envs.py:270-444, storage.py:9-110 of the reference): storage
arithmetic on the CPU; on the GPU the observation kernel writing straight into RolloutStorage.obs."""
import types

import numpy as np
import pytest
import torch

from gymcity_b200 import spaces
from gymcity_b200.envs import RolloutStorage


def test_rollout_storage_returns_and_bookkeeping():
    T, N = 5, 3
    rs = RolloutStorage(T, N, (2, 4, 4), spaces.Discrete(7), 8)
    g = torch.Generator().manual_seed(0)
    for t in range(T):
        slot = rs.next_obs()
        obs = torch.rand(N, 2, 4, 4, generator=g)
        if t % 2:
            slot.copy_(obs)                                   # "the env wrote it already"
            obs = slot
        rs.insert(obs, torch.zeros(N, 8), torch.randint(0, 7, (N, 1), generator=g), torch.rand(N, 1, generator=g),
                  torch.rand(N, 1, generator=g), torch.rand(N, 1, generator=g),
                  (torch.rand(N, 1, generator=g) > 0.3).float())
        assert torch.equal(rs.obs[t + 1], obs)
    assert rs.step == 0 and rs.actions.dtype == torch.long
    nv, gamma, tau = torch.rand(N, 1, generator=g), 0.97, 0.9
    rs.compute_returns(nv, False, gamma, tau)
    ret = nv.clone()
    for t in reversed(range(T)):
        ret = ret * gamma * rs.masks[t + 1] + rs.rewards[t]
        assert torch.allclose(rs.returns[t], ret)
    rs.compute_returns(nv, True, gamma, tau)
    gae, nxt = torch.zeros(N, 1), nv
    for t in reversed(range(T)):
        delta = rs.rewards[t] + gamma * nxt * rs.masks[t + 1] - rs.value_preds[t]
        gae = delta + gamma * tau * rs.masks[t + 1] * gae
        assert torch.allclose(rs.returns[t], gae + rs.value_preds[t])
        nxt = rs.value_preds[t]
    last = rs.obs[-1].clone()
    rs.after_update()
    assert torch.equal(rs.obs[0], last)


@pytest.mark.gpu
def test_make_vec_envs_writes_observations_into_rollout_storage():
    from gymcity_b200.envs import make_vec_envs
    N, T = 6, 4
    args = types.SimpleNamespace(map_width=16, max_step=6, random_terrain=True, power_puzzle=False)
    dev = torch.device("cuda", 0)
    a = make_vec_envs('MicropolisEnv-v0', 3, N, 0.99, None, False, dev, True, args=args)
    b = make_vec_envs('MicropolisEnv-v0', 3, N, 0.99, None, False, dev, True, args=args)
    assert a.observation_space.shape == (32, 16, 16) and a.action_space.n == 20 * 256
    rs = RolloutStorage(T, N, a.observation_space.shape, a.action_space, 1, device=dev)
    o_a, o_b = a.reset(), b.reset()
    assert torch.equal(o_a, o_b)
    rs.obs[0].copy_(o_a)
    rng = np.random.default_rng(0)
    resets = 0
    for it in range(3):
        for t in range(T):
            act = torch.from_numpy(rng.integers(0, 20 * 256, (N, 1))).to(dev)
            obs, rew, done, infos = a.step(act, out=rs.next_obs())
            obs_b, rew_b, done_b, _ = b.step(act)
            assert obs.data_ptr() == rs.next_obs().data_ptr()            # written in place
            assert torch.equal(obs, obs_b) and torch.equal(rew, rew_b) and np.array_equal(done, done_b)
            assert rew.shape == (N, 1) and done.shape == (N,) and len(infos) == N
            resets += int(done.any())
            masks = torch.from_numpy(1.0 - done.astype(np.float32)).reshape(N, 1).to(dev)
            rs.insert(obs, torch.zeros(N, 1, device=dev), act, torch.zeros(N, 1, device=dev), torch.zeros(N, 1, device=dev),
                      rew, masks)
            assert torch.equal(rs.obs[t + 1], obs_b)
        rs.after_update()
    assert resets >= 1                                                   # max_step = 6: the auto-reset path ran
    a.close(); b.close()


@pytest.mark.gpu
def test_make_vec_envs_extinction_and_gol():
    from gymcity_b200.envs import make_vec_envs
    dev = torch.device("cuda", 0)
    args = types.SimpleNamespace(map_width=16, max_step=100, random_terrain=True, extinction_type='age',
                                 extinction_prob=0.5)
    v = make_vec_envs('MicropolisEnv-v0', 1, 4, 0.99, None, False, dev, True, args=args)
    v.reset()
    rng = np.random.default_rng(1)
    for t in range(4):
        v.step(torch.from_numpy(rng.integers(0, 20 * 256, (4, 1))).to(dev))
    assert int((v.env.ages(0) > -1).sum()) == 0                          # an event fired on the last step
    v.close()
    gargs = types.SimpleNamespace(map_width=16, max_step=20, num_proc=64, prob_life=20)
    gol = make_vec_envs('GoLMultiEnv-v0', 0, 1, 0.99, None, False, dev, True, args=gargs)
    obs = gol.reset()
    assert tuple(obs.shape) == (64, 2, 16, 16)
    gol.close()


@pytest.mark.gpu
def test_make_vec_envs_paint_env():
    """MicropolisPaintEnv-v0 through make_vec_envs (envs.py:74): Box action space, one tool per window cell per step."""
    from gymcity_b200.envs import make_vec_envs
    dev = torch.device("cuda", 0)
    N = 3
    args = types.SimpleNamespace(map_width=12, max_step=5, random_terrain=True)
    v = make_vec_envs('MicropolisPaintEnv-v0', 2, N, 0.99, None, False, dev, True, args=args)
    assert tuple(v.action_space.shape) == (12, 12)
    obs = v.reset()
    assert tuple(obs.shape) == (N, 32, 12, 12)
    rng = np.random.default_rng(0)
    ended = 0
    for t in range(8):
        a = torch.from_numpy(rng.integers(0, 20, (N, 12, 12))).to(dev)
        obs, rew, done, infos = v.step(a)
        assert tuple(obs.shape) == (N, 32, 12, 12) and rew.shape == (N, 1) and done.shape == (N,)
        ended += int(done.sum())
    assert ended == N                                                    # max_step = 5: one auto-reset each
    v.close()


@pytest.mark.gpu
def test_views_keep_the_handle_alive_until_released():
    """A tensor handed out by an env (DLPack view of the handle's memory) stays valid after close() and after the env
    object is gone: the handle is destroyed when the last view is released (gymcity_b200/dlpack.py)."""
    import gc
    from gymcity_b200 import GoLMultiEnv
    from gymcity_b200.micropolis_env import MicropolisVecEnv
    env = MicropolisVecEnv(4, 16, seed=1, max_step=50, empty_start=False)
    obs = env.reset()
    want = obs.clone()
    eng = env.micro.engine
    env.close()
    assert eng._h is not None and eng._close_pending               # deferred: `obs` is still alive
    junk = [torch.empty(1 << 20, device="cuda") for _ in range(8)]  # would reuse freed memory
    assert torch.equal(obs, want)
    del env, junk
    gc.collect()
    assert torch.equal(obs, want)
    del obs, want
    gc.collect()
    assert eng._h is None                                           # the last view is gone: destroyed
    g = GoLMultiEnv()
    g.configure(map_width=16, num_proc=8, max_step=5)
    o = g.reset()
    keep = o.clone()
    g.close()
    assert g._h is not None
    assert torch.equal(o, keep)
    del o
    gc.collect()
    assert g._h is None


def test_frame_stack_matches_the_reference_wrapper_logic():
    """VecPyTorchFrameStack(nstack = 3) against a literal per-env restatement of envs.py:425-443 on a scripted venv."""
    from gymcity_b200.envs import VecPyTorchFrameStack
    N, C, W = 5, 2, 3
    g = torch.Generator().manual_seed(1)
    script = [(torch.rand(N, C, W, W, generator=g), np.array([(i + t) % 4 == 0 for i in range(N)])) for t in range(9)]

    class Venv:
        num_envs, action_space = N, spaces.Discrete(4)
        observation_space = spaces.Box(-1, 1, (C, W, W), dtype=np.float32)
        t = 0
        def reset(self):
            self.t = 0
            return script[0][0]
        def step_async(self, a):
            self.t += 1
        def step_wait(self):
            o, d = script[self.t]
            return o, torch.zeros(N, 1), d, [{} for _ in range(N)]
        def close(self):
            pass

    fs = VecPyTorchFrameStack(Venv(), 3, torch.device("cpu"))
    assert fs.observation_space.shape == (3 * C, W, W)
    ref = torch.zeros(N, 3 * C, W, W)
    ref[:, -C:] = script[0][0]
    assert torch.equal(fs.reset(), ref)
    for t in range(1, 9):
        obs, rew, news, infos = fs.step(torch.zeros(N, 1, dtype=torch.long))
        ref[:, :-C] = ref[:, C:].clone()
        for i, new in enumerate(script[t][1]):
            if new:
                ref[i] = 0
        ref[:, -C:] = script[t][0]
        assert torch.equal(obs, ref) and np.array_equal(news, script[t][1])


@pytest.mark.gpu
def test_make_vec_envs_frame_stack_2():
    from gymcity_b200.envs import make_vec_envs
    N = 4
    args = types.SimpleNamespace(map_width=16, max_step=5, random_terrain=True)
    dev = torch.device("cuda", 0)
    a = make_vec_envs('MicropolisEnv-v0', 3, N, 0.99, None, False, dev, True, num_frame_stack=2, args=args)
    b = make_vec_envs('MicropolisEnv-v0', 3, N, 0.99, None, False, dev, True, args=args)
    assert a.observation_space.shape == (64, 16, 16)
    o, prev = a.reset(), b.reset().clone()
    assert torch.equal(o[:, 32:], prev) and not o[:, :32].any()
    rng = np.random.default_rng(0)
    for t in range(8):
        act = torch.from_numpy(rng.integers(0, 20 * 256, (N, 1))).to(dev)
        o, _, done, _ = a.step(act)
        ob, _, done_b, _ = b.step(act)
        assert np.array_equal(done, done_b) and torch.equal(o[:, 32:], ob)
        keep = torch.from_numpy(~done).to(dev).reshape(N, 1, 1, 1)
        assert torch.equal(o[:, :32], prev * keep)
        prev = ob.clone()
    a.close(); b.close()
