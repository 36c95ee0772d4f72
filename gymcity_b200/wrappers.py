"""Wrappers around the batched MicropolisEnv -- the reference's gym_city/wrappers.py.

Extinguisher (wrappers.py:10-141): intermittent extinction events.  Every env of the batch tracks
TileMap.age_order on the device (the num_step of each build, -1 for natural cells); every
`1 / extinction_prob` steps each env is hit by one event, executed by one kernel for the whole batch
(mpb_env_extinguish): 'age' clears the n_dels oldest cells, 'spatial' clears rings around a random
centre, 'random' clears random aged cells, 'Monster' calls the engine's makeMonster().

Deviation, documented: the reference draws the centre / the random cells from numpy's GLOBAL stream;
here they come from the wrapper's own generator (seeded by `seed`), one row of draws per env.
"""
import numpy as np

XT_TYPES = {'age': 0, 'spatial': 1, 'random': 2, 'Monster': 3}


class Extinguisher:
    def __init__(self, env, extinction_type=None, extinction_prob=0.1, xt_dels=25, seed=0):
        self.env = env
        self.np_random = np.random.default_rng(seed)
        self.set_extinction_type(extinction_type, extinction_prob, xt_dels)

    def __getattr__(self, name):                       # gym.Wrapper attribute pass-through
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env

    def set_extinction_type(self, extinction_type, extinction_prob, extinction_dels):   # wrappers.py:24-33
        self.extinction_type = extinction_type
        self.extinction_prob = extinction_prob
        self.n_dels = int(extinction_dels)
        self.extinction_interval = -1 if extinction_prob == 0 else 1 / extinction_prob
        m = self.env.micro
        m._ck(m._L.mpb_env_track_ages(m._h, m.engine._stream()))       # TileMap.init_age_array()

    def reset(self, *a, **kw):
        return self.env.reset(*a, **kw)

    def step(self, a, **kw):                           # wrappers.py:35-40
        """`if self.num_step % self.extinction_interval == 0: self.extinguish(...)` with every env's OWN num_step, decided
        on the device (episodes of different envs end on different steps; an env the step has just auto-reset is not hit:
        in the reference the event precedes the worker's reset, which replaces the city)."""
        out = self.env.step(a, **kw)
        self.extinguish(self.extinction_type, due_interval=self.extinction_interval)
        return out

    def extinguish(self, extinction_type='age', rnd=None, mask=None, due_interval=None):   # wrappers.py:42-55
        """One extinction event in every env (or the masked ones, or -- due_interval -- the envs whose num_step is a
        multiple of it); returns the deletions per env.  rnd: optional int32 [num_envs, n_dels + 2] draws (tests pass
        the oracle's)."""
        if extinction_type not in XT_TYPES:
            return None
        m = self.env.micro
        n = self.env.num_envs
        if rnd is None:
            rnd = self.np_random.integers(0, 2 ** 31 - 1, size=(n, self.n_dels + 2))
        rnd = np.ascontiguousarray(rnd, np.int32).reshape(n, self.n_dels + 2)
        dels = np.zeros(n, np.int32)
        if due_interval is not None:
            m._ck(m._L.mpb_env_extinguish_due(m._h, XT_TYPES[extinction_type], self.n_dels, rnd.ctypes.data,
                                              float(due_interval), dels.ctypes.data, m.engine._stream()))
            return dels
        mk = None if mask is None else np.ascontiguousarray(np.asarray(mask).astype(np.uint8).reshape(n))
        m._ck(m._L.mpb_env_extinguish(m._h, XT_TYPES[extinction_type], self.n_dels, rnd.ctypes.data,
                                      None if mk is None else mk.ctypes.data, dels.ctypes.data, m.engine._stream()))
        return dels

    def ages(self, env=0):
        m = self.env.micro
        out = np.zeros(self.env.MAP_X * self.env.MAP_Y, np.int32)
        m._ck(m._L.mpb_env_get_ages(m._h, int(env), out.ctypes.data))
        return out.reshape(self.env.MAP_X, self.env.MAP_Y)
