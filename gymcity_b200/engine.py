"""EngineBatch: ctypes view of the mpb_* C-ABI -- n independent Micropolis engines on one GPU.

Method names follow the engine methods the reference's MicropolisControl calls through SWIG
(gym_city/envs/corecontrol.py: toolDown, simTick, clearMap, generateMap, setFunds, ...), batched.
"""
import ctypes as C
import os
import re

import numpy as np

from ._lib import lib
from .dlpack import ViewOwner

WORLD_W, WORLD_H = 120, 100
_HDR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include", "mp_snapshot.h")


def _parse_header():
    src = re.sub(r"/\*.*?\*/", "", open(_HDR).read(), flags=re.S)

    def xlist(macro):
        m = re.search(r"#define %s\(X\)(.*?)\n\n" % macro, src, flags=re.S)
        return re.findall(r"X\((\w+)\)", m.group(1))
    m = re.search(r"enum mp_field \{(.*?)\};", src, flags=re.S)
    fields = [t.strip().split("=")[0].strip() for t in m.group(1).split(",") if t.strip()]
    m = re.search(r"typedef struct mp_sprite \{(.*?)\} mp_sprite;", src, flags=re.S)
    spr = re.findall(r"(\w+)\s*[,;]", m.group(1).replace("int32_t", ""))
    return xlist("MP_SCALAR_INT_FIELDS"), xlist("MP_SCALAR_FLOAT_FIELDS"), fields, spr


INT_FIELDS, FLOAT_FIELDS, FIELD_NAMES, SPRITE_FIELDS = _parse_header()
FIELD_ID = {n: i for i, n in enumerate(FIELD_NAMES)}
FIELD_DTYPE = {
    "MPF_MAP": (np.uint16, 12000), "MPF_POP_DENSITY": (np.uint8, 3000),
    "MPF_TRAFFIC_DENSITY": (np.uint8, 3000), "MPF_POLLUTION_DENSITY": (np.uint8, 3000),
    "MPF_LAND_VALUE": (np.uint8, 3000), "MPF_CRIME_RATE": (np.uint8, 3000),
    "MPF_TERRAIN_DENSITY": (np.uint8, 750), "MPF_POWER_GRID": (np.uint8, 12000),
    "MPF_RATE_OF_GROWTH": (np.int16, 195), "MPF_FIRE_STATION": (np.int16, 195),
    "MPF_FIRE_STATION_EFFECT": (np.int16, 195), "MPF_POLICE_STATION": (np.int16, 195),
    "MPF_POLICE_STATION_EFFECT": (np.int16, 195), "MPF_COM_RATE": (np.int16, 195),
    "MPF_RES_HIST": (np.int16, 240), "MPF_COM_HIST": (np.int16, 240), "MPF_IND_HIST": (np.int16, 240),
    "MPF_MONEY_HIST": (np.int16, 240), "MPF_POLLUTION_HIST": (np.int16, 240),
    "MPF_CRIME_HIST": (np.int16, 240), "MPF_MISC_HIST": (np.int16, 120),
    "MPF_POWER_STACK": (np.int16, 6000),
}
PARITY_FIELDS = [n for n in FIELD_NAMES if n not in ("MPF_COUNT", "MPF_POWER_STACK")]
FN = {n: i for i, n in enumerate([
    "MAP_SCAN", "POWER_SCAN", "PTLV_SCAN", "CRIME_SCAN", "POP_DENSITY_SCAN", "FIRE_ANALYSIS",
    "DEC_TRAFFIC", "DEC_ROG", "SET_VALVES", "CLEAR_CENSUS", "TAKE10", "TAKE120", "COLLECT_TAX",
    "CITY_EVAL", "SEND_MESSAGES", "DO_DISASTERS", "MOVE_OBJECTS", "ANIMATE", "SIMULATE",
    "DO_SIM_INIT", "MAKE_FLOOD", "MAKE_TORNADO", "MAKE_EARTHQUAKE", "MAKE_MONSTER", "SET_FIRE",
    "MAKE_EXPLOSION", "GENERATE_SHIP", "GENERATE_PLANE", "GENERATE_COPTER", "GENERATE_TRAIN",
    "MAKE_MELTDOWN"])}


class Scalars(C.Structure):
    _fields_ = ([(n, C.c_int64) for n in INT_FIELDS]
                + [("problemVotes", C.c_int64 * 10), ("problemOrder", C.c_int64 * 4)]
                + [(n, C.c_float) for n in FLOAT_FIELDS])

    def as_dict(self):
        d = {n: int(getattr(self, n)) for n in INT_FIELDS}
        d["problemVotes"] = [int(v) for v in self.problemVotes]
        d["problemOrder"] = [int(v) for v in self.problemOrder]
        for n in FLOAT_FIELDS:
            d[n] = float(getattr(self, n))
        return d


class Sprite(C.Structure):
    _fields_ = [(n, C.c_int32) for n in SPRITE_FIELDS]


_bound = False


def _bind(L):
    global _bound
    if _bound:
        return
    vp, i32 = C.c_void_p, C.c_int
    sig = {
        "mpb_create": (vp, [i32, i32]), "mpb_destroy": (None, [vp]),
        "mpb_last_error": (C.c_char_p, [vp]), "mpb_num_envs": (i32, [vp]),
        "mpb_env_stride": (i32, []), "mpb_launches_per_tick": (i32, [vp]), "mpb_launches_per_env_step": (i32, [vp]), "mpb_state_bytes": (i32, []), "mpb_blocks": (vp, [vp]),
        "mpb_sync": (i32, [vp]), "mpb_read_cycles": (i32, [vp, vp]), "mpb_group_times": (i32, [vp, vp]), "mpb_frame_profile": (i32, [vp, i32, vp]), "mpb_read_schedule": (i32, [vp, vp]), "mpb_seed": (i32, [vp, vp, vp]),
        "mpb_set_funds": (i32, [vp, i32, vp]), "mpb_set_passes": (i32, [vp, i32, vp]),
        "mpb_set_disasters": (i32, [vp, i32, vp]),
        "mpb_clear_map": (i32, [vp, vp, vp]), "mpb_generate": (i32, [vp, vp, vp, vp, vp]),
        "mpb_tool": (i32, [vp, vp, vp, vp]), "mpb_tool_host": (i32, [vp, vp, vp, vp]),
        "mpb_sim_frames": (i32, [vp, i32, vp]), "mpb_sim_tick": (i32, [vp, i32, vp]),
        "mpb_control_sim_tick": (i32, [vp, vp]), "mpb_call": (i32, [vp, i32, i32, i32, vp]),
        "mpb_get_field": (i32, [vp, i32, i32, vp, i32]), "mpb_set_field": (i32, [vp, i32, i32, vp, i32]),
        "mpb_get_scalars": (i32, [vp, i32, vp]), "mpb_set_scalars": (i32, [vp, i32, vp]),
        "mpb_get_sprites": (i32, [vp, i32, vp, i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _bound = True


class EngineBatch(ViewOwner):
    def __init__(self, n_envs, device=0):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("gymcity_b200 Micropolis needs a CUDA device (no CPU fallback)")
        self._L = lib()
        _bind(self._L)
        h = self._L.mpb_create(int(n_envs), int(device))
        if not h:
            raise RuntimeError("mpb_create failed (n_envs=%d, device=%d)" % (n_envs, device))
        self._h = C.c_void_p(h)
        self.n = int(n_envs)
        self.device = int(device)

    def _destroy_native(self):
        """mpb_destroy.  close() (dlpack.ViewOwner) calls it at once when no DLPack view of the handle's buffers is
        alive, otherwise when the last one is released."""
        if getattr(self, "_h", None):
            self._L.mpb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self._destroy_native()
        except Exception:
            pass

    def _ck(self, rc):
        if rc < 0:
            raise RuntimeError(self._L.mpb_last_error(self._h).decode())
        return rc

    def _stream(self):
        import torch
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _i32(self, a):
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.int32), (self.n,)))
        return a

    @staticmethod
    def _mask(mask, n):
        if mask is None:
            return None, None
        m = np.ascontiguousarray(np.asarray(mask).astype(np.uint8).reshape(n))
        return m, m.ctypes.data

    # --- batched engine methods
    def seed(self, seeds):
        a = self._i32(seeds)
        self._ck(self._L.mpb_seed(self._h, a.ctypes.data, self._stream()))

    def setFunds(self, v): self._ck(self._L.mpb_set_funds(self._h, int(v), self._stream()))
    def setPasses(self, v): self._ck(self._L.mpb_set_passes(self._h, int(v), self._stream()))
    def setEnableDisasters(self, v): self._ck(self._L.mpb_set_disasters(self._h, int(bool(v)), self._stream()))

    def clearMap(self, mask=None):
        m, p = self._mask(mask, self.n)
        self._ck(self._L.mpb_clear_map(self._h, p, self._stream()))

    def generate(self, terrain_seeds, sim_seeds, mask=None):
        a, b = self._i32(terrain_seeds), self._i32(sim_seeds)
        m, p = self._mask(mask, self.n)
        self._ck(self._L.mpb_generate(self._h, a.ctypes.data, b.ctypes.data, p, self._stream()))

    def toolDown(self, tools, xs, ys):
        txy = np.ascontiguousarray(np.stack([self._i32(tools), self._i32(xs), self._i32(ys)], 1))
        res = np.zeros(self.n, np.int32)
        self._ck(self._L.mpb_tool_host(self._h, txy.ctypes.data, res.ctypes.data, self._stream()))
        return res

    def simFrames(self, n): self._ck(self._L.mpb_sim_frames(self._h, int(n), self._stream()))
    def simTick(self): self._ck(self._L.mpb_sim_tick(self._h, 0, self._stream()))
    def tickEngine(self): self._ck(self._L.mpb_sim_tick(self._h, 1, self._stream()))
    def controlSimTick(self): self._ck(self._L.mpb_control_sim_tick(self._h, self._stream()))
    def call(self, fn, a=0, b=0): self._ck(self._L.mpb_call(self._h, FN[fn], int(a), int(b), self._stream()))
    def sync(self): self._ck(self._L.mpb_sync(self._h))

    def group_times(self):
        out = np.zeros(8, np.float32)
        g = self._ck(self._L.mpb_group_times(self._h, out.ctypes.data))
        return out[:g]

    def frame_profile(self, enable=True, read=False):
        out = np.zeros((self.n, 256), np.uint32) if read else None
        self._ck(self._L.mpb_frame_profile(self._h, int(enable), out.ctypes.data if read else None))
        return out

    def read_schedule(self):
        out = np.zeros(self.n, np.int32)
        g = self._ck(self._L.mpb_read_schedule(self._h, out.ctypes.data))
        return out, g

    def read_cycles(self):
        out = np.zeros(self.n, np.uint64)
        self._ck(self._L.mpb_read_cycles(self._h, out.ctypes.data))
        return out

    # --- per-env state access
    def get(self, env, field):
        dt, n = FIELD_DTYPE[field]
        out = np.zeros(n, dtype=dt)
        r = self._ck(self._L.mpb_get_field(self._h, int(env), FIELD_ID[field], out.ctypes.data, out.nbytes))
        assert r == out.nbytes, (field, r)
        return out

    def set(self, env, field, arr):
        dt, n = FIELD_DTYPE[field]
        a = np.ascontiguousarray(arr, dtype=dt).reshape(-1)
        assert a.size == n
        self._ck(self._L.mpb_set_field(self._h, int(env), FIELD_ID[field], a.ctypes.data, a.nbytes))

    def scalars(self, env):
        s = Scalars()
        self._ck(self._L.mpb_get_scalars(self._h, int(env), C.byref(s)))
        return s

    def set_scalars(self, env, s):
        self._ck(self._L.mpb_set_scalars(self._h, int(env), C.byref(s)))

    def sprites(self, env):
        buf = (Sprite * 64)()
        n = self._ck(self._L.mpb_get_sprites(self._h, int(env), buf, 64))
        return [tuple(int(getattr(buf[i], f)) for f in SPRITE_FIELDS) for i in range(n)]

    def snapshot(self, env):
        d = {f: self.get(env, f) for f in PARITY_FIELDS}
        d["scalars"] = self.scalars(env).as_dict()
        d["sprites"] = self.sprites(env)
        return d
