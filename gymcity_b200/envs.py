"""Vec-env adapter and rollout storage for the batched env -- what the reference's training loop
(train.py:108, envs.py:270-308 make_vec_envs, envs.py:347-444 VecPyTorch / VecPyTorchFrameStack,
storage.py:9-110 RolloutStorage) sees, with the batch living on the GPU:

* `make_vec_envs(...)` takes the reference's arguments and returns ONE object that plays
  SubprocVecEnv + VecPyTorch + VecPyTorchFrameStack(1): `reset()` -> obs (N,32,W,H) float CUDA tensor,
  `step(actions (N,1))` -> (obs, reward (N,1) float, done (N,) bool ndarray, infos), auto-reset of
  finished envs (subproc_vec_env.py:16-21).
* `RolloutStorage` keeps the reference's fields and methods; `step(actions, out=rollouts.next_obs())`
  makes the env's observation kernel write straight into `rollouts.obs[step + 1]`, and `insert`
  then skips its copy of the (N,32,W,H) batch (mpb_env_set_obs_buffer).
"""
import numpy as np

from . import spaces


class BatchedVecEnv:
    def __init__(self, env, device):
        self.venv = self.env = env
        self.device = device
        self.num_envs = env.num_envs
        self.observation_space = spaces.Box(-1, 1, (env.micro.obs_channels, env.MAP_X, env.MAP_Y), dtype=np.float32)
        self.action_space = env.action_space
        self._pending = None

    def reset(self):                                         # VecPyTorch.reset + VecPyTorchFrameStack.reset
        return self.env.reset()

    def step_async(self, actions, out=None):                 # VecPyTorch.step_async (envs.py:364-366)
        # one flat action per env; the paint env takes a tool per window cell, (N, W, H)
        self._pending = (actions if getattr(self.env, "paint_actions", False) else actions.reshape(-1), out)

    def step_wait(self):                                     # VecPyTorch.step_wait (envs.py:368-376)
        actions, out = self._pending
        self._pending = None
        obs, reward, done, infos = self.env.step(actions, out=out)
        return obs, reward.float().reshape(-1, 1), done.cpu().numpy(), infos

    def step(self, actions, out=None):
        self.step_async(actions, out)
        return self.step_wait()

    def get_param_bounds(self):
        return self.env.get_param_bounds()

    def set_param_bounds(self, bounds):
        return self.env.set_param_bounds(bounds)

    def set_params(self, params):
        return self.env.set_params(params)

    def close(self):
        self.env.close()


class VecPyTorchFrameStack:
    """envs.py:406-456 for nstack > 1: the last `nstack` observations of every env along the channel axis, newest
    last; an env that finished (and was auto-reset) keeps only the observation of its new episode.  Batch tensors
    on the env's device, no per-env loop."""

    def __init__(self, venv, nstack, device=None):
        import torch
        self.venv, self.nstack = venv, int(nstack)
        self.num_envs, self.action_space, self.device = venv.num_envs, venv.action_space, device
        wos = venv.observation_space
        self.shape_dim0 = wos.shape[0]
        self.observation_space = spaces.Box(-1, 1, (wos.shape[0] * self.nstack,) + tuple(wos.shape[1:]), dtype=np.float32)
        self.stacked_obs = torch.zeros((venv.num_envs,) + self.observation_space.shape, device=device)

    def _push(self, obs, news=None):
        import torch
        c = self.shape_dim0
        self.stacked_obs[:, :-c] = self.stacked_obs[:, c:].clone()
        if news is not None:
            keep = 1.0 - torch.as_tensor(np.asarray(news), dtype=torch.float32).to(self.stacked_obs.device)
            self.stacked_obs[:, :-c] *= keep.reshape(-1, 1, 1, 1)
        self.stacked_obs[:, -c:] = obs
        return self.stacked_obs

    def step_async(self, actions):
        self.venv.step_async(actions)

    def step_wait(self):
        obs, rews, news, infos = self.venv.step_wait()
        return self._push(obs, news), rews, news, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def reset(self):
        obs = self.venv.reset()
        self.stacked_obs.zero_()
        return self._push(obs)

    def close(self):
        self.venv.close()

    def get_param_bounds(self):
        return self.venv.get_param_bounds()

    def set_param_bounds(self, bounds):
        return self.venv.set_param_bounds(bounds)

    def set_params(self, params):
        return self.venv.set_params(params)


def make_vec_envs(env_name, seed, num_processes, gamma, log_dir, add_timestep, device, allow_early_resets,
                  num_frame_stack=None, args=None):
    """envs.py:270-308 for the envs of this package; `args` carries the reference's CLI fields
    (map_width, max_step, random_terrain, power_puzzle, random_builds, extinction_type, extinction_prob,
    prob_life, render ...), missing ones take the reference's defaults (arguments.py)."""
    g = lambda k, d=None: getattr(args, k, d) if args is not None else d
    dev_index = 0
    if hasattr(device, "index") and device.index is not None:
        dev_index = device.index
    name = env_name.lower()
    if 'golmultienv' in name:                               # one env object that holds all the boards (envs.py:273-283)
        from .gol_env import GoLMultiEnv
        env = GoLMultiEnv()
        env.configure(map_width=g('map_width', 16), prob_life=g('prob_life', 20), max_step=g('max_step', 200),
                      num_proc=g('num_proc', num_processes), device=dev_index)
        env.seed(seed)
        return env
    if 'micropolis' not in name:
        raise NotImplementedError("%s: only MicropolisEnv-v0, MicropolisPaintEnv-v0 and GoLMultiEnv-v0 are on the tick path" % env_name)
    from .micropolis_env import MicropolisPaintEnv, MicropolisVecEnv
    cls = MicropolisPaintEnv if 'paint' in name else MicropolisVecEnv      # envs.py:74, gym_city/envs/paintenv.py
    env = cls(num_processes, g('map_width', 20), device=dev_index, seed=seed,
                           max_step=g('max_step', None), empty_start=not g('random_terrain', True),
                           power_puzzle=bool(g('power_puzzle', False)), random_builds=bool(g('random_builds', False)),
                           poet=bool(g('poet', False)),
                           env_offset=g('env_offset', 0))
    if g('extinction_type') is not None:                    # envs.py:255-257
        from .wrappers import Extinguisher
        env = Extinguisher(env, g('extinction_type'), g('extinction_prob', 0.1), seed=seed)
    venv = BatchedVecEnv(env, device)
    if num_frame_stack not in (None, 1):                    # envs.py:301-303
        return VecPyTorchFrameStack(venv, num_frame_stack, device)
    return venv


class RolloutStorage:
    """storage.py:9-110 on the device.  Shapes as in the reference: obs (T+1, N, C, W, H), rewards /
    value_preds / action_log_probs (T, N, 1), returns / masks (T+1, N, 1), actions (T, N, 1) long."""

    def __init__(self, num_steps, num_processes, obs_shape, action_space, recurrent_hidden_state_size, args=None,
                 device="cpu"):
        import torch
        T, N = num_steps, num_processes
        z = lambda *shape, **kw: torch.zeros(*shape, device=device, **kw)
        self.args = args
        self.obs = z(T + 1, N, *obs_shape)
        if isinstance(recurrent_hidden_state_size, tuple):
            self.recurrent_hidden_states = z(T + 1, 2, N, *recurrent_hidden_state_size)
        else:
            self.recurrent_hidden_states = z(T + 1, N, recurrent_hidden_state_size)
        self.rewards, self.value_preds, self.returns = z(T, N, 1), z(T, N, 1), z(T + 1, N, 1)
        self.action_log_probs = z(T, N, 1)
        discrete = action_space.__class__.__name__ == 'Discrete'
        self.actions = z(T, N, 1 if discrete else action_space.shape[0], dtype=torch.long if discrete else torch.float32)
        self.masks = torch.ones(T + 1, N, 1, device=device)
        self.num_steps, self.step = T, 0

    def to(self, device):
        for k in ('obs', 'recurrent_hidden_states', 'rewards', 'value_preds', 'returns', 'action_log_probs', 'actions',
                  'masks'):
            setattr(self, k, getattr(self, k).to(device))

    def next_obs(self):
        """The slot the next `insert` fills: pass it as `out=` to the env's step."""
        return self.obs[self.step + 1]

    def insert(self, obs, recurrent_hidden_states, actions, action_log_probs, value_preds, rewards, masks):
        slot = self.obs[self.step + 1]
        if obs.data_ptr() != slot.data_ptr():               # the env already wrote into the slot: no copy
            slot.copy_(obs)
        self.recurrent_hidden_states[self.step + 1] = recurrent_hidden_states
        self.actions[self.step].copy_(actions)
        self.action_log_probs[self.step].copy_(action_log_probs)
        self.value_preds[self.step].copy_(value_preds)
        self.rewards[self.step].copy_(rewards)
        self.masks[self.step + 1].copy_(masks)
        self.step = (self.step + 1) % self.num_steps

    def after_update(self):
        for t in (self.obs, self.recurrent_hidden_states, self.masks):
            t[0].copy_(t[-1])

    def compute_returns(self, next_value, use_gae, gamma, tau):
        T = self.rewards.size(0)
        if use_gae:
            vp = self.value_preds
            gae = 0
            nxt = next_value
            for t in reversed(range(T)):
                delta = self.rewards[t] + gamma * nxt * self.masks[t + 1] - vp[t]
                gae = delta + gamma * tau * self.masks[t + 1] * gae
                self.returns[t] = gae + vp[t]
                nxt = vp[t]
        else:
            self.returns[-1] = next_value
            for t in reversed(range(T)):
                self.returns[t] = self.returns[t + 1] * gamma * self.masks[t + 1] + self.rewards[t]
