"""Game-of-Life envs with the reference's API, stepping on the GPU through the golb_* C-ABI.

GoLMultiEnv   mirrors game_of_life/envs/multi_env.py:21-265 (batched, XOR-toggle action, torus tick,
              target-population reward, obs = [target plane, state]).
GameOfLifeEnv mirrors game_of_life/envs/env.py:24-203 over world.py (single env semantics incl. the
              stale-neighbour behaviour of World.build_cell, world.py:108-130), batched over
              `num_proc` independent boards (num_proc=1 reproduces the reference object exactly).
"""
import ctypes as C
from collections import OrderedDict

import numpy as np

from . import spaces
from ._lib import lib
from .dlpack import ViewOwner, device_tensor

GOLB_ROWS, GOLB_OBS, GOLB_REWARD, GOLB_POP, GOLB_DONE = range(5)


def _stream_ptr(device):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _GolBase(ViewOwner):
    classic = 0

    def __init__(self):
        self.num_tools = 1
        self._h = None
        self._L = None

    def _create(self, num_proc, map_width, max_step, device):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("gymcity_b200 Game of Life needs a CUDA device (no CPU fallback)")
        self._L = lib()
        self.close()
        self.device = int(device)
        h = self._L.golb_create(int(num_proc), int(map_width), int(max_step), self.classic,
                                self.device)
        if not h:
            raise RuntimeError("golb_create failed (n=%d, W=%d, max_step=%s)"
                               % (num_proc, map_width, max_step))
        self._h = C.c_void_p(h)
        self.num_proc = int(num_proc)
        self.map_width = int(map_width)
        self.max_step = int(max_step)
        n, w = self.num_proc, self.map_width
        P = lambda which: self._L.golb_ptr(self._h, which)
        self._reward = device_tensor(P(GOLB_REWARD), "float32", (n, 1), self.device, owner=self)
        self._pop = device_tensor(P(GOLB_POP), "int32", (n,), self.device, owner=self)
        self._done = device_tensor(P(GOLB_DONE), "uint8", (n,), self.device, owner=self)
        self._rows = device_tensor(P(GOLB_ROWS), "int32", (n, w), self.device, owner=self)
        self.intsToActions = [[0, i // w, i % w] for i in range(w * w)]
        self.actionsToInts = np.arange(w * w).reshape(1, w, w)
        self.action_space = spaces.Discrete(self.num_tools * w * w)

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError(self._L.golb_last_error(self._h).decode())

    def _destroy_native(self):
        if getattr(self, "_h", None):
            self._L.golb_destroy(self._h)
            self._h = None

    def close(self):
        """Frees the handle once no tensor handed out by this env is alive any more (dlpack.ViewOwner)."""
        for k in ("_obs", "_reward", "_pop", "_done", "_rows"):
            if hasattr(self, k):
                setattr(self, k, None)
        ViewOwner.close(self)

    def __del__(self):
        try:
            self._destroy_native()
        except Exception:
            pass

    def seed(self, seed=None):
        self._seed = 0 if seed is None else int(seed)
        self._episode = 0
        return [self._seed]

    def set_state(self, cells):
        """cells: (N, W, W) or (N,1,W,W) array of 0/1 -- explicit boards (tests / fixtures)."""
        a = np.ascontiguousarray(np.asarray(cells).reshape(self.num_proc, self.map_width,
                                                           self.map_width), dtype=np.uint8)
        self._ck(self._L.golb_set_state(self._h, a.ctypes.data, _stream_ptr(self.device)))
        self.step_count = 0

    def get_state(self):
        out = np.zeros((self.num_proc, self.map_width, self.map_width), dtype=np.uint8)
        self._ck(self._L.golb_get_state(self._h, out.ctypes.data))
        return out

    def _reset_world(self):
        seed = (getattr(self, "_seed", 0) * 1000003 + getattr(self, "_episode", 0)) & 0xffffffff
        self._episode = getattr(self, "_episode", 0) + 1
        self._ck(self._L.golb_reset(self._h, seed, float(self.prob_life), _stream_ptr(self.device)))
        self.step_count = 0

    def _actions_dev(self, a):
        import torch
        if not torch.is_tensor(a):
            a = np.asarray(a).reshape(-1)
            # intsToActions[a] raises for an action outside the board (env.py:143, multi_env.py:182); device
            # tensors cannot be checked without a sync -- the kernels skip the build for such an action
            lim = self.map_width * self.map_width
            if a.size and (np.abs(a).max() >= lim or (not self.classic and a.min() < 0)):
                raise IndexError("action out of range for a %dx%d board" % (self.map_width, self.map_width))
            a = torch.as_tensor(a, dtype=torch.int64)
        a = a.reshape(-1).to(device="cuda:%d" % self.device, dtype=torch.int64).contiguous()
        assert a.numel() == self.num_proc
        return a


class GoLMultiEnv(_GolBase):
    classic = 0

    def configure(self, render=False, map_width=16, prob_life=20, max_step=200, num_proc=1,
                  record=None, cuda=True, device=0):
        self.prob_life = prob_life / 100            # world_pytorch.py:31
        self._create(num_proc, map_width, max_step, device)
        n, w = self.num_proc, self.map_width
        max_pop = w * w * 2 / 3                     # multi_env.py:57
        self.param_bounds = OrderedDict({'pop': (0, max_pop)})
        self.param_ranges = [abs(ub - lb) for lb, ub in self.param_bounds.values()]
        self.params = OrderedDict({'pop': max_pop})
        self.num_params = 1
        self.observation_space = spaces.Box(0, 1, (n, 2, w, w), dtype=np.int64)
        self._obs = device_tensor(self._L.golb_ptr(self._h, GOLB_OBS), "float32", (n, 2, w, w),
                                  self.device, owner=self)
        self.step_count = 0
        self.set_params(self.params)

    def get_param_bounds(self):
        return self.param_bounds

    def set_param_bounds(self, bounds):
        self.param_bounds = bounds

    def set_params(self, params):
        self.params = params
        v = float(np.float32(list(params.values())[0]))
        self._ck(self._L.golb_set_target(self._h, v, float(np.float32(self.param_ranges[0]))))

    def get_pop(self):
        return self._pop

    def get_obs(self):
        return self._obs

    def reset(self):
        self._reset_world()
        return self._obs

    def step(self, a):
        a = self._actions_dev(a)
        self._ck(self._L.golb_step(self._h, a.data_ptr(), _stream_ptr(self.device)))
        self.step_count += 1
        return self._obs, self._reward, self._done.bool(), [{}]

    def step_n(self, actions, obs_seq=None, reward_seq=None):
        """T steps in one launch (golb_step_n) for scripted / random-action rollouts: actions (T, N) int64 CUDA
        tensor; obs_seq (T, N, 2, W, W) / reward_seq (T, N) float32 CUDA tensors receive every tick's observation /
        reward when given.  Returns what the T-th step() would have."""
        import torch
        a = actions.to(device="cuda:%d" % self.device, dtype=torch.int64).contiguous()
        T = a.shape[0]
        assert a.dim() == 2 and a.shape[1] == self.num_proc
        for t, shape in ((obs_seq, (T, self.num_proc, 2, self.map_width, self.map_width)), (reward_seq, (T, self.num_proc))):
            assert t is None or (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == shape)
        self._ck(self._L.golb_step_n(self._h, a.data_ptr(), T, obs_seq.data_ptr() if obs_seq is not None else None,
                                     reward_seq.data_ptr() if reward_seq is not None else None, _stream_ptr(self.device)))
        self.step_count += T
        return self._obs, self._reward, self._done.bool(), [{}]

    def step_host(self, actions_np, reward_out):
        """End-to-end call with HOST buffers (int64 actions in, float32 rewards out)."""
        self._ck(self._L.golb_step_host(self._h, actions_np.ctypes.data, reward_out.ctypes.data,
                                        _stream_ptr(self.device)))
        self.step_count += 1


class GameOfLifeEnv(_GolBase):
    classic = 1

    def configure(self, render=False, map_width=16, prob_life=20, record=None, max_step=None,
                  num_proc=1, device=0):
        # world.py:85,90: alive = randint(0,100) <= prob_life  ->  p = (prob_life + 1) / 100
        self.prob_life = (prob_life + 1) / 100
        if max_step is None:
            raise ValueError("max_step is required (env.py:174 divides by it)")
        self._create(num_proc, map_width, max_step, device)
        n, w = self.num_proc, self.map_width
        self.size = w
        self.observation_space = spaces.Box(0, 1, (1, w, w), dtype=np.int64)
        self._obs = device_tensor(self._L.golb_ptr(self._h, GOLB_OBS), "uint8", (n, 1, w, w),
                                  self.device, owner=self)
        self.step_count = 0

    def reset(self):
        self._reset_world()
        return self._state_out()

    def _state_out(self):
        if self.num_proc == 1:
            return self._obs[0].cpu().numpy()
        return self._obs

    def step(self, a):
        a = self._actions_dev(a)
        self._ck(self._L.golb_step(self._h, a.data_ptr(), _stream_ptr(self.device)))
        self.step_count += 1
        if self.num_proc == 1:
            return (self._state_out(), float(self._reward[0, 0].item()),
                    bool(self._done[0].item()), {})
        return self._obs, self._reward, self._done.bool(), [{}]
