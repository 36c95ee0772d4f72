"""MicropolisEnv / MicropolisControl with the reference's API, stepping N cities per call on the GPU.

Mirrors gym_city/envs/env.py (MicropolisEnv: setMapSize / seed / reset / step / close / set_params /
get_param_bounds, attributes micro, num_step, max_step, city_metrics, city_trgs, weights,
action_space, observation_space) and gym_city/envs/corecontrol.py (MicropolisControl: takeAction,
simTick, getDensityMaps, getTile, doSimTool, clearMap, newMap, setFunds/getFunds, get*Pop, tools,
engineTools, MAP_XS = 16, MAP_YS = 8), but every call acts on a batch: `num_envs` independent cities
resident on one GPU, one fused kernel launch per step.  With num_envs = 1 (the default) the return
types are the reference's (numpy observation, float reward, bool done, dict info); with
num_envs > 1 they are the vec-env's (envs.py:347-375: obs (N,32,W,H) float32 tensor, reward (N,1),
done (N,) bool, infos list) and live on the device (zero-copy DLPack views of the handle's buffers).

Seeding: the reference engine reseeds itself from the wall clock on construction and inside
generateSomeCity (random.cpp:155-164, initialize.cpp:76-78), so it is not reproducible as shipped.
Here seed(s) fixes, per env i of the batch (global index env_offset + i), the terrain seed and the
simRandom seed of every episode; the same (seed, global env index, episode) gives the same rollout
on any number of GPUs.
"""
import ctypes as C
from collections import OrderedDict

import numpy as np

from . import spaces
from .dlpack import device_tensor
from .engine import EngineBatch, FIELD_ID

ENGINE_TOOLS = ['Residential', 'Commercial', 'Industrial', 'FireDept', 'PoliceDept', 'Query', 'Wire', 'Clear',
                'Rail', 'Road', 'Stadium', 'Park', 'Seaport', 'CoalPowerPlant', 'NuclearPowerPlant', 'Airport',
                'Net', 'Water', 'Land', 'Forest']                                   # corecontrol.py:61-80
AGENT_TOOLS = ['Residential', 'Commercial', 'Industrial', 'FireDept', 'PoliceDept', 'Clear', 'Wire', 'Rail',
               'Road', 'Stadium', 'Park', 'Seaport', 'CoalPowerPlant', 'NuclearPowerPlant', 'Airport', 'Net',
               'Water', 'Land', 'Forest', 'Nil']                                    # corecontrol.py:85-106
ZONES = ['Residential', 'Commercial', 'Industrial', 'Seaport', 'Stadium', 'PoliceDept', 'FireDept', 'Airport',
         'NuclearPowerPlant', 'CoalPowerPlant', 'Road', 'Rail', 'Park', 'Wire', 'Rubble', 'Net', 'Water', 'Land',
         'Forest', 'Church', 'Hospital', 'Fire', 'RoadWire', 'RailWire', 'Bridge', 'RoadRail', 'WaterWire',
         'Radar', 'RailBridge']                                                     # tilemap.py:85-125
NUM_FEATURES = 22
MPB_ENV_OBS, MPB_ENV_REWARD, MPB_ENV_DONE = 0, 1, 2
METRICS = ['res_pop', 'com_pop', 'ind_pop', 'traffic', 'num_plants', 'mayor_rating']


from .sharding import mix as _mix  # noqa: E402


class _ShadowMap:
    """Read-only view of the device TileMap shadow (tilemap.py attributes the env layer reads)."""

    def __init__(self, control):
        self._c = control
        self.zones = ZONES
        self.num_zones = len(ZONES)
        self.num_features = NUM_FEATURES

    def _shadow(self, env=0):
        c = self._c
        n = c.MAP_X * c.MAP_Y
        z, s, ce, k = np.zeros(n, np.uint8), np.zeros(n, np.uint8), np.zeros(n, np.int16), np.zeros(16)
        c._ck(c._L.mpb_env_get_shadow(c._h, int(env), z.ctypes.data, s.ctypes.data, ce.ctypes.data, k.ctypes.data))
        shp = (c.MAP_X, c.MAP_Y)
        return z.reshape(shp), s.reshape(shp), ce.reshape(shp), k

    def zone_ids(self, env=0):
        return self._shadow(env)[0]

    @property
    def static_builds(self):
        return self._shadow(0)[1][None].astype(int)

    @property
    def num_plants(self):
        return int(self._shadow(0)[3][0])

    @property
    def num_roads(self):
        return int(self._shadow(0)[3][1])

    def counters(self, env=0):
        k = self._shadow(env)[3]
        names = ["num_plants", "num_roads", "n_struct_tiles", "total_traffic", "num_step", "tilemap_error",
                 "reward", "done", "funds"] + METRICS + ["sprite_overflow", "episode"]
        return dict(zip(names, [float(v) for v in k]))


class MicropolisControl:
    """Batched stand-in for corecontrol.MicropolisControl over the mpb_* C-ABI."""

    def __init__(self, env, MAP_W=12, MAP_H=12, PADDING=0, gui=False, rank=None, power_puzzle=False,
                 paint=False, num_envs=1, device=0, MAP_XS=16, MAP_YS=8, max_step=None):
        if PADDING != 0:
            raise NotImplementedError("PADDING != 0 breaks the reference's own observation concat "
                                      "(env.py:340, tilemap.py:548-557); only 0 is supported")
        if gui:
            raise NotImplementedError("GTK rendering is outside the tick path")
        env.micro = self
        self.env = env
        self.MAP_X, self.MAP_Y, self.PADDING = int(MAP_W), int(MAP_H), 0
        self.MAP_XS, self.MAP_YS = int(MAP_XS), int(MAP_YS)
        self.engineTools = ENGINE_TOOLS
        self.tools = ['Wire'] if power_puzzle else AGENT_TOOLS
        self.num_tools = len(self.tools)
        self.num_envs = int(num_envs)
        self.engine = EngineBatch(self.num_envs, device)
        self._L, self._h, self._ck = self.engine._L, self.engine._h, self.engine._ck
        self._bind()
        tz = np.array([(-1 if t == 'Nil' else (-2 if t == 'Clear' else ZONES.index(t))) for t in self.tools], np.int8)
        te = np.array([(0 if t == 'Nil' else ENGINE_TOOLS.index(t)) for t in self.tools], np.int8)
        trg = np.array([env.city_trgs[m] for m in METRICS], np.float64)
        wt = np.array([env.weights[m] for m in METRICS], np.float64)
        ms = self.MAP_X * self.MAP_Y if max_step is None else int(max_step)
        self._ck(self._L.mpb_env_configure(self._h, self.MAP_X, self.MAP_Y, self.MAP_XS, self.MAP_YS,
                                           self.num_tools, tz.ctypes.data, te.ctypes.data, ms,
                                           trg.ctypes.data, wt.ctypes.data))
        self.map = _ShadowMap(self)
        self.zones, self.num_zones = ZONES, len(ZONES)
        self.init_funds = 2000000
        self.win1 = None
        self.player_builds = []
        n, W, H = self.num_envs, self.MAP_X, self.MAP_Y
        P = lambda w: self._L.mpb_env_ptr(self._h, w)
        dev = self.engine.device
        self.obs_channels = 32
        self.obs = device_tensor(P(MPB_ENV_OBS), "float32", (n, 32, W, H), dev, owner=self.engine)
        self.reward = device_tensor(P(MPB_ENV_REWARD), "float32", (n, 1), dev, owner=self.engine)
        self.done = device_tensor(P(MPB_ENV_DONE), "uint8", (n,), dev, owner=self.engine)

    def set_poet(self, param_ranges):
        """env.py:301-313 poet observation: 38 planes (pops / ranges, targets / ranges); rebinds self.obs."""
        r = np.ascontiguousarray(param_ranges, np.float64)
        self.obs_channels = self._ck(self._L.mpb_env_set_poet(self._h, 1, r.ctypes.data))
        self.obs = device_tensor(self._L.mpb_env_ptr(self._h, MPB_ENV_OBS), "float32",
                                 (self.num_envs, self.obs_channels, self.MAP_X, self.MAP_Y), self.engine.device,
                                 owner=self.engine)

    def _bind(self):
        L, vp, i32 = self._L, C.c_void_p, C.c_int
        sig = {
            "mpb_env_configure": (i32, [vp, i32, i32, i32, i32, i32, vp, vp, i32, vp, vp]),
            "mpb_env_set_targets": (i32, [vp, vp, vp]), "mpb_env_set_max_step": (i32, [vp, i32]),
            "mpb_env_reset_begin": (i32, [vp, i32, vp, vp, vp, vp]), "mpb_env_reset_finish": (i32, [vp, vp, vp]),
            "mpb_env_action": (i32, [vp, vp, i32, i32, vp, vp]), "mpb_env_get_counters": (i32, [vp, vp]), "mpb_env_step": (i32, [vp, vp, i32, vp]),
            "mpb_env_step_host": (i32, [vp, vp, vp, vp, vp]), "mpb_env_ptr": (vp, [vp, i32]),
            "mpb_env_get_shadow": (i32, [vp, i32, vp, vp, vp, vp]),
            "mpb_env_track_ages": (i32, [vp, vp]), "mpb_env_extinguish": (i32, [vp, i32, i32, vp, vp, vp, vp]),
            "mpb_env_extinguish_due": (i32, [vp, i32, i32, vp, C.c_double, vp, vp]),
            "mpb_env_get_ages": (i32, [vp, i32, vp]), "mpb_env_set_obs_buffer": (i32, [vp, vp]),
            "mpb_env_step_paint": (i32, [vp, vp, vp]), "mpb_env_set_poet": (i32, [vp, i32, vp]),
            "mpb_env_step_masked": (i32, [vp, vp, i32, vp, vp]),
            "mpb_env_set_seed": (i32, [vp, C.c_longlong, i32]), "mpb_env_set_auto_reset": (i32, [vp, i32, i32]),
            "mpb_env_set_num_step": (i32, [vp, vp, vp]), "mpb_env_update_map": (i32, [vp, vp, vp]), "mpb_env_prefetch_stats": (i32, [vp, vp, vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args

    # ---- corecontrol.py API (batched: scalars broadcast, arrays are per env)
    def _flat(self, tool_idx, x, y):
        return (np.asarray(tool_idx, np.int64) * self.MAP_X * self.MAP_Y
                + np.asarray(x, np.int64) * self.MAP_Y + np.asarray(y, np.int64))

    def _actions(self, flat):
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(flat, np.int64).reshape(-1), (self.num_envs,)))
        return a

    def apply_actions(self, flat, static_build=False, full_tools=False):
        """addZoneBot for one flat action per env (< 0 = none), no tick."""
        flat = self._actions(flat)
        ok = np.zeros(self.num_envs, np.int32)
        self._ck(self._L.mpb_env_action(self._h, flat.ctypes.data, int(bool(static_build)),
                                        int(bool(full_tools)), ok.ctypes.data, self.engine._stream()))
        return ok

    def takeAction(self, a, static_build=False):            # corecontrol.py:332-338
        return self.apply_actions(self._flat(a[0], a[1], a[2]), static_build)

    def doBotTool(self, x, y, tool, static_build=False):    # corecontrol.py:276-278 (tool by NAME)
        if tool == 'Nil':
            return np.ones(self.num_envs, np.int32)
        return self.apply_actions(self._flat(AGENT_TOOLS.index(tool), x, y), static_build, full_tools=True)

    def counters(self):
        """(N, 17) array: num_plants, num_roads, n_struct_tiles, total_traffic, num_step, tilemap_error,
        reward, done, funds, metrics[6], sprite_overflow, episode for every env."""
        out = np.zeros((self.num_envs, 17), np.float64)
        self._ck(self._L.mpb_env_get_counters(self._h, out.ctypes.data))
        return out

    def prefetch_stats(self):
        """(active, resets served from a prepared terrain, resets that generated theirs in line)."""
        a, b = C.c_uint(), C.c_uint()
        on = self._ck(self._L.mpb_env_prefetch_stats(self._h, C.byref(a), C.byref(b)))
        return bool(on), a.value, b.value

    def doSimTool(self, x, y, tool):                        # corecontrol.py:296-301
        return self.engine.toolDown(self.engineTools.index(tool), np.asarray(x) + self.MAP_XS,
                                    np.asarray(y) + self.MAP_YS)

    def getTile(self, x, y, env=0):                         # corecontrol.py:302-305
        wx, wy = x + self.MAP_XS, y + self.MAP_YS
        if not (0 <= wx < 120 and 0 <= wy < 100):
            return 0
        return int(self.engine.get(env, "MPF_MAP")[wx * 100 + wy]) & 1023

    def simTick(self):                                      # corecontrol.py:148-152
        self.engine.controlSimTick()

    def clearMap(self, mask=None):                          # corecontrol.py:188-192
        m, p = EngineBatch._mask(mask, self.num_envs)
        self._ck(self._L.mpb_env_reset_begin(self._h, 0, None, None, p, self.engine._stream()))

    def newMap(self, terrain_seeds, sim_seeds, mask=None):  # corecontrol.py:184-186 (+ explicit seeds)
        a, b = self.engine._i32(terrain_seeds), self.engine._i32(sim_seeds)
        m, p = EngineBatch._mask(mask, self.num_envs)
        self._ck(self._L.mpb_env_reset_begin(self._h, 1, a.ctypes.data, b.ctypes.data, p, self.engine._stream()))

    def newMapDerived(self, mask=None):
        """clearMap + generateSomeCity with the seeds of each env's next episode, derived on the device from
        (seed, global env index, episode) -- mpb_env_set_seed; sharding.mix is the same hash on the host."""
        m, p = EngineBatch._mask(mask, self.num_envs)
        self._ck(self._L.mpb_env_reset_begin(self._h, 1, None, None, p, self.engine._stream()))

    def updateMap(self, mask=None):                         # corecontrol.py:193-203
        """The TileMap shadow of the window re-read from the engine map (after loading a city into the engine)."""
        m, p = EngineBatch._mask(mask, self.num_envs)
        self._ck(self._L.mpb_env_update_map(self._h, p, self.engine._stream()))

    def getDensityMaps(self):                               # corecontrol.py:215-239 -> obs planes 22..24
        return self.obs[:, 22:25]

    def setFunds(self, funds): self.engine.setFunds(funds)
    def getFunds(self, env=0): return int(self.engine.scalars(env).totalFunds)
    def getResPop(self, env=0): return int(self.engine.scalars(env).resPop)
    def getComPop(self, env=0): return int(self.engine.scalars(env).comPop)
    def getIndPop(self, env=0): return int(self.engine.scalars(env).indPop)
    def getTotPop(self, env=0): return int(self.engine.scalars(env).totalPop)

    @property
    def total_traffic(self):
        return int(self.map.counters(0)["total_traffic"])

    def close(self):
        """Drops this object's own views, then frees the handle -- at once, or when the last tensor a caller still holds
        (an observation, a reward) is gone (dlpack.ViewOwner): no view outlives the memory it aliases."""
        self.obs = self.reward = self.done = None
        self.engine.close()


class MicropolisEnv:
    def __init__(self, MAP_X=20, MAP_Y=20, PADDING=0, num_envs=1, device=0, env_offset=0):
        self.num_envs, self.device, self.env_offset = int(num_envs), int(device), int(env_offset)
        self.num_episode = 0
        self.city_trgs = OrderedDict({'res_pop': 500, 'com_pop': 50, 'ind_pop': 50, 'traffic': 2000,
                                      'num_plants': 14, 'mayor_rating': 100})          # env.py:31-38
        self.trg_param_vals = np.array([v for v in self.city_trgs.values()])
        self.param_bounds = OrderedDict({'res_pop': (0, 750), 'com_pop': (0, 100), 'ind_pop': (0, 100),
                                         'traffic': (0, 2000), 'num_plants': (0, 100), 'mayor_rating': (0, 100)})
        self.weights = OrderedDict({'res_pop': 1, 'com_pop': 1, 'ind_pop': 1, 'traffic': 1, 'num_plants': 0,
                                    'mayor_rating': 0})                                  # env.py:49-56
        self.num_params = 6
        self.param_ranges = [abs(ub - lb) for lb, ub in self.param_bounds.values()]
        self.max_loss = sum(r * w for r, w in zip(self.param_ranges, self.weights.values()))
        self.city_metrics = {}
        self.max_reward = 100
        self._seed = 0
        self.auto_reset = True

    def seed(self, seed=None):                              # env.py:91-98
        self._seed = 0 if seed is None else int(seed)
        self.np_random = np.random.default_rng(self._seed)
        if hasattr(self, 'micro'):
            self.micro._ck(self.micro._L.mpb_env_set_seed(self.micro._h, self._seed, self.env_offset))
        return [self._seed]

    def setMapSize(self, size, max_step=None, rank=0, print_map=False, PADDING=0, static_builds=True,
                   parallel_gui=False, render_gui=False, empty_start=True, simple_reward=False,
                   power_puzzle=False, record=False, traffic_only=False, random_builds=False, poet=False,
                   num_envs=None, device=None, env_offset=None, MAP_XS=16, MAP_YS=8, **kwargs):
        """env.py:100-139 (setMapSize / pre_gui / post_gui) with the same keyword arguments."""
        if record:
            raise NotImplementedError
        if num_envs is not None:
            self.num_envs = int(num_envs)
        if device is not None:
            self.device = int(device)
        if env_offset is not None:
            self.env_offset = int(env_offset)
        self.PADDING, self.rank, self.render_gui = PADDING, rank, render_gui
        self.random_builds, self.traffic_only = random_builds, traffic_only
        self.MAP_X, self.MAP_Y = (size, size) if isinstance(size, int) else (size[0], size[1])
        self.max_step = self.MAP_X * self.MAP_Y if max_step is None else max_step
        self.empty_start, self.simple_reward, self.power_puzzle = empty_start, simple_reward, power_puzzle
        self.static_builds, self.poet, self.print_map = True, poet, print_map
        if hasattr(self, 'micro'):
            self.micro.close()
        self.micro = MicropolisControl(self, self.MAP_X, self.MAP_Y, PADDING, rank=rank,
                                       power_puzzle=power_puzzle, gui=render_gui, num_envs=self.num_envs,
                                       device=self.device, MAP_XS=MAP_XS, MAP_YS=MAP_YS, max_step=self.max_step)
        self.micro._ck(self.micro._L.mpb_env_set_seed(self.micro._h, self._seed, self.env_offset))
        self._device_auto_reset = False
        self.num_step = 0
        self.minFunds = 0
        self.initFunds = self.micro.init_funds
        self.num_tools, self.num_zones = self.micro.num_tools, self.micro.num_zones
        self.num_obs_channels = NUM_FEATURES + 6 + 3 + 1
        if poet:                                               # env.py:157-158, 305-312
            self.micro.set_poet(self.param_ranges)
            self.num_obs_channels += len(self.city_trgs)
        self.action_space = spaces.Discrete(self.num_tools * self.MAP_X * self.MAP_Y)
        self.observation_space = spaces.Box(-1, 1, (self.num_obs_channels, self.MAP_X, self.MAP_Y), dtype=float)
        self.intsToActions = _IntsToActions(self.num_tools, self.MAP_X, self.MAP_Y)
        self.actionsToInts = np.arange(self.num_tools * self.MAP_X * self.MAP_Y).reshape(
            self.num_tools, self.MAP_X, self.MAP_Y)
        self.city_metrics = self.get_city_metrics()
        self.last_city_metrics = self.city_metrics
        self.win1 = None

    # ---- helpers
    def _single(self):
        return self.num_envs == 1

    def _out(self):
        m = self.micro
        if self._single():
            return (m.obs[0].cpu().numpy().astype(np.float64), float(m.reward[0, 0].item()),
                    bool(m.done[0].item()) and bool(self.auto_reset), {})
        done = m.done.bool()
        if not self.auto_reset:                              # env.py:519-520: terminal = (...) and self.auto_reset
            done = done & False
        return self._cur_obs(), m.reward, done, _Infos(self.num_envs)

    def _cur_obs(self):
        obs = getattr(self, "_obs_out", None)
        return self.micro.obs if obs is None else obs

    def episode_seeds(self, which, episodes=None):
        """The seeds the device derives for the NEXT reset of every env (which = 0 terrain, 1 simulation):
        sharding.mix(seed, global env index, episode of the env, which).  For tests and reference runs."""
        ep = self.micro.counters()[:, 16].astype(np.int64) if episodes is None else np.broadcast_to(episodes, (self.num_envs,))
        g = np.arange(self.num_envs) + self.env_offset
        return np.array([_mix(self._seed, int(i), int(k), which) for i, k in zip(g, ep)], np.int32)

    def get_param_bounds(self):
        return self.param_bounds

    def set_param_bounds(self, bounds):
        self.param_bounds = bounds

    def set_params(self, trgs):                             # env.py:419-424
        for k, v in trgs.items():
            self.city_trgs[k] = v
        self.trg_param_vals = np.array([v for v in self.city_trgs.values()])
        trg = np.array([self.city_trgs[m] for m in METRICS], np.float64)
        self.micro._ck(self.micro._L.mpb_env_set_targets(self.micro._h, trg.ctypes.data, None))

    def get_city_metrics(self, env=0):                      # env.py:426-440
        c = self.micro.map.counters(env)
        return {m: c[m] for m in METRICS}

    def getRating(self, env=0):
        return int(self.micro.engine.scalars(env).cityYes)

    # ---- episode control
    def powerPuzzle(self, mask=None):                       # env.py:253-262
        """5 static Residential zones and static NuclearPowerPlants until one sticks, at uniform random
        window positions per env (the reference draws them from numpy's global RNG)."""
        n, m = self.num_envs, self.micro
        live = np.ones(n, bool) if mask is None else np.asarray(mask, bool).copy()
        res, nuke = AGENT_TOOLS.index('Residential'), AGENT_TOOLS.index('NuclearPowerPlant')
        for _ in range(5):
            flat = m._flat(res, self.np_random.integers(0, self.MAP_X, n), self.np_random.integers(0, self.MAP_Y, n))
            flat[~live] = -1
            m.apply_actions(flat, True, full_tools=True)
        for _ in range(200):
            need = live & (m.counters()[:, 0] == 0)
            if not need.any():
                break
            flat = m._flat(nuke, self.np_random.integers(0, self.MAP_X, n), self.np_random.integers(0, self.MAP_Y, n))
            flat[~need] = -1
            m.apply_actions(flat, True, full_tools=True)

    def reset(self, mask=None, setup_actions=None, static_start=None):
        """env.py:264-292.  `mask` (N,) selects the envs to reset (None = all).  `setup_actions` is an
        optional list of (flat_action (N,), static_build) applied between clearMap/newMap and the first
        simTick -- the hook the reference uses for powerPuzzle / randomStaticStart (env.py:275-278)."""
        m = self.micro
        if self.empty_start:
            m.clearMap(mask)
        else:
            m.newMapDerived(mask)                            # seeds: episode_seeds(), derived in the kernel
        self.num_step = 0
        if self.power_puzzle and setup_actions is None:
            self.powerPuzzle(mask)
        for item in (setup_actions or []):
            flat, sb = item[0], item[1]
            full = item[2] if len(item) > 2 else False
            flat = m._actions(flat).copy()
            if mask is not None:
                flat[~np.asarray(mask, bool)] = -1
            m.apply_actions(flat, sb, full_tools=full)
        if self.random_builds or static_start is not None:
            self._random_static_start(mask, *(static_start or ()))
        mk, p = EngineBatch._mask(mask, self.num_envs)
        m._ck(m._L.mpb_env_reset_finish(m._h, p, m.engine._stream()))
        self.num_episode += 1
        if self._single():
            self.city_metrics = self.get_city_metrics()
            self.last_city_metrics = self.city_metrics
            self.state = m.obs[0].cpu().numpy().astype(np.float64)
            return self.state
        return self._cur_obs()

    def _random_static_start(self, mask=None, ks=None, actions=None):   # env.py:228-241
        """randomStaticStart: every env being reset takes its own number k of FULL env steps (random action
        with static_build=True, simulation tick, num_step + 1) before the reset's closing simTick; the
        envs that are done with theirs, and the ones not being reset, sit the step out
        (mpb_env_step_masked).  ks (N,) / actions (max k, N): the draws, for tests; by default k ~
        U{0..W*H/10} and uniform actions from the env's generator."""
        m, n = self.micro, self.num_envs
        num_static = int(self.MAP_X * self.MAP_Y / 10)
        m.setFunds(m.init_funds)
        if ks is None:
            ks = self.np_random.integers(0, num_static + 1, n) if num_static > 0 else np.zeros(n, np.int64)
        ks = np.asarray(ks).reshape(n)
        live = np.ones(n, bool) if mask is None else np.asarray(mask, bool).reshape(n)
        for j in range(int(ks.max()) if n else 0):
            a = (self.np_random.integers(0, self.action_space.n, n) if actions is None else np.asarray(actions[j])).astype(np.int64)
            mk = np.ascontiguousarray(((j < ks) & live).astype(np.uint8))
            if not mk.any():
                continue
            a = np.ascontiguousarray(a)
            m._ck(m._L.mpb_env_step_masked(m._h, a.ctypes.data, 1, mk.ctypes.data, m.engine._stream()))

    def set_num_steps(self, num_step):
        """Per-env MicropolisEnv.num_step (N,): episodes of different envs end on different steps."""
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(num_step, np.int32), (self.num_envs,)))
        self.micro._ck(self.micro._L.mpb_env_set_num_step(self.micro._h, a.ctypes.data, self.micro.engine._stream()))

    def set_obs_buffer(self, out=None):
        """Observations of the following steps / resets go to `out` (float32 CUDA tensor (N,32,W,H),
        contiguous -- e.g. RolloutStorage.obs[step + 1]) instead of the env's own buffer; None restores it."""
        m = self.micro
        if out is None:
            m._ck(m._L.mpb_env_set_obs_buffer(m._h, None))
            self._obs_out = None
            return
        import torch
        assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous()
        assert tuple(out.shape) == (self.num_envs, self.micro.obs_channels, self.MAP_X, self.MAP_Y)
        assert out.device.index == self.device
        m._ck(m._L.mpb_env_set_obs_buffer(m._h, out.data_ptr()))
        self._obs_out = out

    def step(self, a, static_build=False, out=None):         # env.py:448-540
        import torch
        m = self.micro
        if out is not None:                                   # an explicit set_obs_buffer() stays in force otherwise
            self.set_obs_buffer(out)
        if torch.is_tensor(a):
            act = a.reshape(-1).to(device="cuda:%d" % self.device, dtype=torch.int64).contiguous()
        else:
            act = torch.as_tensor(np.asarray(a, np.int64).reshape(-1)).to("cuda:%d" % self.device)
        assert act.numel() == self.num_envs
        m._ck(m._L.mpb_env_step(m._h, act.data_ptr(), int(bool(static_build)), m.engine._stream()))
        self.num_step += 1
        out = self._out()
        if self._single():
            self.state = out[0]
            self.last_city_metrics = self.city_metrics
            self.city_metrics = self.get_city_metrics()
        return out

    def step_host(self, actions_np, reward_out, done_out):
        """End to end with HOST buffers: int64 actions in, float32 reward / uint8 done out."""
        m = self.micro
        m._ck(m._L.mpb_env_step_host(m._h, actions_np.ctypes.data, reward_out.ctypes.data,
                                     done_out.ctypes.data, m.engine._stream()))
        self.num_step += 1

    def close(self):
        if hasattr(self, 'micro'):
            self.micro.close()

    def render(self, mode='human'):
        raise NotImplementedError("GTK rendering is outside the tick path")


class _IntsToActions:
    """intsToActions[i] -> [z, x, y] with i = z*W*H + x*H + y (env.py:209-219), without the dict."""

    def __init__(self, T, W, H):
        self.T, self.W, self.H = T, W, H

    def __len__(self):
        return self.T * self.W * self.H

    def __getitem__(self, i):
        i = int(i)
        n = self.W * self.H
        return [i // n, (i % n) // self.H, i % self.H]


class _Infos:
    """The per-env info dicts of a step (the reference's workers return {} each): a sequence of N empty dicts
    that costs nothing until someone indexes it (a list of 32 768 fresh dicts per step is ~2 ms of Python)."""

    def __init__(self, n):
        self._n, self._d = n, None

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if self._d is None:
            self._d = [{} for _ in range(self._n)]
        return self._d[i]

    def __iter__(self):
        return (self[i] for i in range(self._n))


class MicropolisVecEnv(MicropolisEnv):
    """The batch as the vec env (envs.py:347-375, 406-444): reset() -> obs, step(actions (N,1)) ->
    (obs, reward (N,1), done (N,), infos), with the worker's auto-reset (subproc_vec_env.py:16-21)."""

    def __init__(self, num_envs, size, device=0, env_offset=0, seed=0, **kwargs):
        super().__init__(num_envs=num_envs, device=device, env_offset=env_offset)
        self.seed(seed)
        self.setMapSize(size, **kwargs)

    def setMapSize(self, size, **kwargs):
        super().setMapSize(size, **kwargs)
        self.set_device_auto_reset(self.num_envs > 1 and not self.power_puzzle and not self.random_builds)

    def set_device_auto_reset(self, on):
        """The worker's auto-reset (subproc_vec_env.py:16-21) inside the step, on the device: no `done.any()`
        round trip, no host-side reset call.  Off for the variants whose reset needs host-drawn set-up builds
        (powerPuzzle, randomStaticStart): those keep the host path below."""
        self._device_auto_reset = bool(on)
        self.micro._ck(self.micro._L.mpb_env_set_auto_reset(self.micro._h, int(bool(on) and self.auto_reset),
                                                            int(not self.empty_start)))

    def step(self, a, static_build=False, out=None):
        if self._device_auto_reset:
            if not self.auto_reset:                                      # switched off after setMapSize
                self.set_device_auto_reset(False)
            else:
                return super().step(a, static_build, out=out)            # done envs come back already reset
        obs, reward, done, infos = super().step(a, static_build, out=out)
        if self.num_envs == 1:
            if done:
                obs = self.reset()
            return obs, reward, done, infos
        if self.auto_reset and bool(done.any().item()):
            mask = done.cpu().numpy()
            done = done.clone()
            reward = reward.clone()
            self.reset(mask=mask)
            obs = self._cur_obs()
        return obs, reward, done, infos


class MicropolisPaintEnv(MicropolisVecEnv):
    """gym_city/envs/paintenv.py: the agent names a tool for EVERY window cell each step.  Batched:
    step(a) takes (N, W, H) tool indices (uint8 / integer tensor or array); the action space is the
    reference's Box(0, num_tools - 1, (W, H))."""
    paint_actions = True

    def setMapSize(self, size, **kwargs):
        super().setMapSize(size, **kwargs)
        hi = np.full((self.MAP_X, self.MAP_Y), self.num_tools - 1)
        self.action_space = spaces.Box(np.zeros_like(hi), hi, dtype=np.int64)     # paintenv.py:36-39

    def step(self, a, static_build=False, out=None):
        import torch
        m = self.micro
        dev = "cuda:%d" % self.device
        if not torch.is_tensor(a):
            a = torch.as_tensor(np.asarray(a))
        t = a.to(device=dev).reshape(self.num_envs, self.MAP_X * self.MAP_Y).to(torch.uint8).contiguous()
        if out is not None:
            self.set_obs_buffer(out)
        m._ck(m._L.mpb_env_step_paint(m._h, t.data_ptr(), m.engine._stream()))
        self.num_step += 1
        obs, reward, done, infos = self._out()
        if self._device_auto_reset and self.auto_reset:
            return obs, reward, done, infos
        if self.num_envs > 1 and self.auto_reset and bool(done.any().item()):
            mask = done.cpu().numpy()
            done, reward = done.clone(), reward.clone()
            self.reset(mask=mask)
            obs = self._cur_obs()
        return obs, reward, done, infos
