"""CPU harness (TEST INFRASTRUCTURE): builds tests/hostbuild/host_engine.cpp -- the same engine
headers the CUDA kernels compile, built with g++ for one host thread -- into
tests/_hostbuild/libmph.so and wraps it with the oracle's EngineBase ABI so the CPU suite can diff
the engine logic against the reference engine.  gymcity_b200 itself never loads this."""
import ctypes as C
import glob
import os
import subprocess

from oracle.ref_engine import EngineBase

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "hostbuild", "host_engine.cpp")
LIB = os.path.join(HERE, "_hostbuild", "libmph.so")


def build_host_engine():
    if os.environ.get("MPH_LIB"):                    # the ASan / UBSan build (tools/sanitize_host.sh)
        return os.environ["MPH_LIB"]
    deps = [SRC] + glob.glob(os.path.join(ROOT, "gymcity_b200", "csrc", "mp_*.h")) \
        + glob.glob(os.path.join(ROOT, "include", "*.h"))
    if os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-o", LIB, SRC])
    return LIB


class HostEngine(EngineBase):
    def __init__(self):
        super().__init__(C.CDLL(build_host_engine()), "mph_")


import numpy as np  # noqa: E402

from oracle.ref_gym import AGENT_TOOLS, ENGINE_TOOLS  # noqa: E402

ZONES = ['Residential', 'Commercial', 'Industrial', 'Seaport', 'Stadium', 'PoliceDept', 'FireDept', 'Airport',
         'NuclearPowerPlant', 'CoalPowerPlant', 'Road', 'Rail', 'Park', 'Wire', 'Rubble', 'Net', 'Water', 'Land',
         'Forest', 'Church', 'Hospital', 'Fire', 'RoadWire', 'RailWire', 'Bridge', 'RoadRail', 'WaterWire', 'Radar',
         'RailBridge']
TRG = [500, 50, 50, 2000, 14, 100]
WEIGHT = [1, 1, 1, 1, 0, 0]


def tool_tables(tools):
    tz = [(-1 if t == 'Nil' else (-2 if t == 'Clear' else ZONES.index(t))) for t in tools]
    te = [(0 if t == 'Nil' else ENGINE_TOOLS.index(t)) for t in tools]
    return np.array(tz, np.int8), np.array(te, np.int8)


class HostGym(HostEngine):
    """One env of the gym layer on the CPU harness (mph_env_*)."""

    def configure(self, W, H, max_step, power_puzzle=False, xs=16, ys=8):
        L = self._lib
        self.W, self.H = W, H
        tools = ['Wire'] if power_puzzle else AGENT_TOOLS
        tz, te = tool_tables(tools)
        trg, wt = np.array(TRG, np.float64), np.array(WEIGHT, np.float64)
        L.mph_env_configure.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.mph_env_configure(self._h, W, H, xs, ys, len(tools), tz.ctypes.data, te.ctypes.data, max_step,
                            trg.ctypes.data, wt.ctypes.data)
        L.mph_env_reset_begin.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.mph_env_reset_finish.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mph_env_action.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int]
        L.mph_env_step.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mph_env_get_shadow.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        self.obs = np.zeros((32, W, H), np.float32)
        self.rew = np.zeros(1, np.float32)
        self.done = np.zeros(1, np.uint8)

    def reset_begin(self, terrain, tseed=0, sseed=0):
        self._lib.mph_env_reset_begin(self._h, int(terrain), int(tseed), int(sseed))

    def reset_finish(self):
        self._lib.mph_env_reset_finish(self._h, self.obs.ctypes.data, self.rew.ctypes.data, self.done.ctypes.data)
        return self.obs.copy()

    def action(self, a, static_build=False, full=False):
        return self._lib.mph_env_action(self._h, int(a), int(static_build), int(full))

    def step(self, a, static_build=False):
        self._lib.mph_env_step(self._h, int(a), int(static_build), self.obs.ctypes.data, self.rew.ctypes.data,
                               self.done.ctypes.data)
        return self.obs.copy(), float(self.rew[0]), bool(self.done[0])

    def set_poet(self, ranges, trg=None):
        r = np.ascontiguousarray(ranges, np.float64)
        self._lib.mph_env_set_poet.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        self._lib.mph_env_set_poet(self._h, 1, r.ctypes.data)
        if trg is not None:
            t = np.ascontiguousarray(trg, np.float64)
            self._lib.mph_env_set_targets.argtypes = [C.c_void_p, C.c_void_p]
            self._lib.mph_env_set_targets(self._h, t.ctypes.data)
        self.obs = np.zeros((38, self.W, self.H), np.float32)

    def step_paint(self, tools):
        t = np.ascontiguousarray(tools, np.uint8).reshape(-1)
        assert t.size == self.W * self.H
        self._lib.mph_env_step_paint.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        self._lib.mph_env_step_paint(self._h, t.ctypes.data, self.obs.ctypes.data, self.rew.ctypes.data,
                                     self.done.ctypes.data)
        return self.obs.copy(), float(self.rew[0]), bool(self.done[0])

    def track_ages(self):
        self._lib.mph_env_track_ages.argtypes = [C.c_void_p]
        self._lib.mph_env_track_ages(self._h)

    def extinguish(self, type_id, n_dels, rnd):
        r = np.ascontiguousarray(rnd, np.int32)
        assert r.size >= n_dels + 2
        self._lib.mph_env_extinguish.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        return self._lib.mph_env_extinguish(self._h, int(type_id), int(n_dels), r.ctypes.data)

    def ages(self):
        out = np.zeros(self.W * self.H, np.int32)
        self._lib.mph_env_get_ages.argtypes = [C.c_void_p, C.c_void_p]
        self._lib.mph_env_get_ages(self._h, out.ctypes.data)
        return out.reshape(self.W, self.H)

    def shadow(self):
        n = self.W * self.H
        z, s, c, k = np.zeros(n, np.uint8), np.zeros(n, np.uint8), np.zeros(n, np.int16), np.zeros(15)
        self._lib.mph_env_get_shadow(self._h, z.ctypes.data, s.ctypes.data, c.ctypes.data, k.ctypes.data)
        return z.reshape(self.W, self.H), s.reshape(self.W, self.H), c.reshape(self.W, self.H), k
