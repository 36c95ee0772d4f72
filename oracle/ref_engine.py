"""ctypes wrapper over oracle/_ref/libmicropolis_ref.so -- TEST INFRASTRUCTURE ONLY.

The library is the UNMODIFIED reference Micropolis engine (plus the one problemTable pad, see
oracle/build_ref.sh) behind the extern "C" shim in oracle/ref_shim.cpp.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.environ.get("MPREF_LIB", os.path.join(HERE, "_ref", "libmicropolis_ref.so"))     # MPREF_LIB: the ASan / UBSan build (tools/sanitize_host.sh)
SNAPSHOT_H = os.path.join(HERE, "..", "include", "mp_snapshot.h")

WORLD_W, WORLD_H = 120, 100


def _parse_snapshot_header():
    """Builds the ctypes mirror of mp_scalars / mp_sprite / enum mp_field from the header."""
    src = open(SNAPSHOT_H).read()
    src_nc = re.sub(r"/\*.*?\*/", "", src, flags=re.S)

    def xlist(macro):
        m = re.search(r"#define %s\(X\)(.*?)\n\n" % macro, src_nc, flags=re.S)
        return re.findall(r"X\((\w+)\)", m.group(1))

    ints = xlist("MP_SCALAR_INT_FIELDS")
    floats = xlist("MP_SCALAR_FLOAT_FIELDS")
    m = re.search(r"enum mp_field \{(.*?)\};", src_nc, flags=re.S)
    fields = [t.strip().split("=")[0].strip() for t in m.group(1).split(",") if t.strip()]
    m = re.search(r"typedef struct mp_sprite \{(.*?)\} mp_sprite;", src_nc, flags=re.S)
    spr = re.findall(r"(\w+)\s*[,;]", m.group(1).replace("int32_t", ""))
    return ints, floats, fields, spr


INT_FIELDS, FLOAT_FIELDS, FIELD_NAMES, SPRITE_FIELDS = _parse_snapshot_header()
FIELD_ID = {n: i for i, n in enumerate(FIELD_NAMES)}


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

    def as_tuple(self):
        return tuple(int(getattr(self, n)) for n in SPRITE_FIELDS)


FIELD_DTYPE = {
    "MPF_MAP": (np.uint16, 12000),
    "MPF_POP_DENSITY": (np.uint8, 3000), "MPF_TRAFFIC_DENSITY": (np.uint8, 3000),
    "MPF_POLLUTION_DENSITY": (np.uint8, 3000), "MPF_LAND_VALUE": (np.uint8, 3000),
    "MPF_CRIME_RATE": (np.uint8, 3000), "MPF_TERRAIN_DENSITY": (np.uint8, 750),
    "MPF_POWER_GRID": (np.uint8, 12000),
    "MPF_RATE_OF_GROWTH": (np.int16, 195), "MPF_FIRE_STATION": (np.int16, 195),
    "MPF_FIRE_STATION_EFFECT": (np.int16, 195), "MPF_POLICE_STATION": (np.int16, 195),
    "MPF_POLICE_STATION_EFFECT": (np.int16, 195), "MPF_COM_RATE": (np.int16, 195),
    "MPF_RES_HIST": (np.int16, 240), "MPF_COM_HIST": (np.int16, 240),
    "MPF_IND_HIST": (np.int16, 240), "MPF_MONEY_HIST": (np.int16, 240),
    "MPF_POLLUTION_HIST": (np.int16, 240), "MPF_CRIME_HIST": (np.int16, 240),
    "MPF_MISC_HIST": (np.int16, 120), "MPF_POWER_STACK": (np.int16, 6000),
}

# Fields that make up "the engine state" for rollout parity.
PARITY_FIELDS = [n for n in FIELD_NAMES if n not in ("MPF_COUNT", "MPF_POWER_STACK")]

FN = {n: i for i, n in enumerate([
    "MAP_SCAN", "POWER_SCAN", "PTLV_SCAN", "CRIME_SCAN", "POP_DENSITY_SCAN",
    "FIRE_ANALYSIS", "DEC_TRAFFIC", "DEC_ROG", "SET_VALVES", "CLEAR_CENSUS",
    "TAKE10", "TAKE120", "COLLECT_TAX", "CITY_EVAL", "SEND_MESSAGES",
    "DO_DISASTERS", "MOVE_OBJECTS", "ANIMATE", "SIMULATE", "DO_SIM_INIT",
    "MAKE_FLOOD", "MAKE_TORNADO", "MAKE_EARTHQUAKE", "MAKE_MONSTER", "SET_FIRE",
    "MAKE_EXPLOSION", "GENERATE_SHIP", "GENERATE_PLANE", "GENERATE_COPTER",
    "GENERATE_TRAIN", "MAKE_MELTDOWN"])}


def ref_available(path=REF_LIB):
    return os.path.exists(path)


class EngineBase:
    """Common ctypes plumbing for any library exporting the `<prefix>_*` engine ABI."""

    def __init__(self, lib, prefix):
        self._lib = lib
        self._p = prefix
        f = lambda n: getattr(lib, prefix + n)
        f("new").restype = C.c_void_p
        for n, args in [("free", [C.c_void_p]), ("seed", [C.c_void_p, C.c_int]),
                        ("clear_map", [C.c_void_p]), ("generate", [C.c_void_p, C.c_int]),
                        ("set_funds", [C.c_void_p, C.c_int]), ("set_passes", [C.c_void_p, C.c_int]),
                        ("sim_tick", [C.c_void_p]), ("tick_engine", [C.c_void_p]),
                        ("control_sim_tick", [C.c_void_p]), ("sim_frames", [C.c_void_p, C.c_int]),
                        ("make_monster", [C.c_void_p]),
                        ("call", [C.c_void_p, C.c_int, C.c_int, C.c_int]),
                        ("get_scalars", [C.c_void_p, C.POINTER(Scalars)]),
                        ("set_scalars", [C.c_void_p, C.POINTER(Scalars)])]:
            f(n).argtypes = args
            f(n).restype = None
        f("tool").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        f("tool").restype = C.c_int
        f("get_field").argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        f("get_field").restype = C.c_int
        f("set_field").argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        f("set_field").restype = C.c_int
        f("get_sprites").argtypes = [C.c_void_p, C.POINTER(Sprite), C.c_int]
        f("get_sprites").restype = C.c_int
        self._h = C.c_void_p(f("new")())

    def _f(self, n):
        return getattr(self._lib, self._p + n)

    def close(self):
        if self._h:
            self._f("free")(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- engine methods the gym path calls (corecontrol.py / micropolisgtkengine.py)
    def seed(self, s): self._f("seed")(self._h, int(s))
    def clearMap(self): self._f("clear_map")(self._h)
    def generate(self, s): self._f("generate")(self._h, int(s))
    def toolDown(self, tool, x, y): return self._f("tool")(self._h, int(tool), int(x), int(y))
    def setFunds(self, v): self._f("set_funds")(self._h, int(v))

    def loadFileDir(self, path):
        fn = self._f("load_file_dir")
        fn.argtypes = [C.c_void_p, C.c_char_p]
        fn.restype = C.c_int
        return bool(fn(self._h, str(path).encode()))
    def setPasses(self, v): self._f("set_passes")(self._h, int(v))
    def simTick(self): self._f("sim_tick")(self._h)
    def tickEngine(self): self._f("tick_engine")(self._h)
    def controlSimTick(self): self._f("control_sim_tick")(self._h)
    def simFrames(self, n): self._f("sim_frames")(self._h, int(n))
    def makeMonster(self): self._f("make_monster")(self._h)
    def call(self, fn, a=0, b=0): self._f("call")(self._h, FN[fn], int(a), int(b))

    # --- state access
    def get(self, field):
        dt, n = FIELD_DTYPE[field]
        out = np.zeros(n, dtype=dt)
        r = self._f("get_field")(self._h, FIELD_ID[field], out.ctypes.data, out.nbytes)
        assert r == out.nbytes, (field, r)
        return out

    def set(self, field, arr):
        dt, n = FIELD_DTYPE[field]
        a = np.ascontiguousarray(arr, dtype=dt).reshape(-1)
        assert a.size == n, (field, a.size, n)
        r = self._f("set_field")(self._h, FIELD_ID[field], a.ctypes.data, a.nbytes)
        assert r == 0, (field, r)

    def scalars(self):
        s = Scalars()
        self._f("get_scalars")(self._h, C.byref(s))
        return s

    def set_scalars(self, s):
        self._f("set_scalars")(self._h, C.byref(s))

    def sprites(self):
        buf = (Sprite * 64)()
        n = self._f("get_sprites")(self._h, buf, 64)
        return [buf[i].as_tuple() for i in range(n)]

    def getTile(self, x, y):
        if not (0 <= x < WORLD_W and 0 <= y < WORLD_H):
            return 0
        return int(self.get("MPF_MAP")[x * WORLD_H + y])

    def snapshot(self):
        d = {f: self.get(f) for f in PARITY_FIELDS}
        d["scalars"] = self.scalars().as_dict()
        d["sprites"] = self.sprites()
        return d


def diff_snapshots(a, b, ignore_scalars=()):
    """Returns a list of human-readable differences between two snapshots (empty == equal)."""
    out = []
    for f in PARITY_FIELDS:
        if not np.array_equal(a[f], b[f]):
            idx = np.nonzero(a[f] != b[f])[0]
            i = int(idx[0])
            out.append("%s: %d cells differ, first at %d (%d vs %d)"
                       % (f, idx.size, i, int(a[f][i]), int(b[f][i])))
    for k, v in a["scalars"].items():
        if k in ignore_scalars:
            continue
        if v != b["scalars"][k]:
            out.append("scalar %s: %r vs %r" % (k, v, b["scalars"][k]))
    if a["sprites"] != b["sprites"]:
        out.append("sprites: %r vs %r" % (a["sprites"], b["sprites"]))
    return out


class RefEngine(EngineBase):
    """The reference C++ engine."""

    def __init__(self, path=REF_LIB):
        if not os.path.exists(path):
            raise FileNotFoundError(
                "%s missing: run oracle/build_ref.sh where /root/reference exists" % path)
        lib = C.CDLL(path)
        super().__init__(lib, "mpref_")
        lib.mpref_bench_rollout.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_int, C.c_uint, C.POINTER(C.c_long)]
        lib.mpref_bench_rollout.restype = C.c_double

    def bench_rollout(self, steps, win_w=16, win_h=16, xs=16, ys=8, act_seed=0):
        frames = C.c_long(0)
        secs = self._lib.mpref_bench_rollout(self._h, steps, win_w, win_h, xs, ys,
                                             act_seed, C.byref(frames))
        return secs, frames.value
