"""`.cty` city files as a wire format for env states (reference: MicropolisEngine/src/fileio.cpp:160-330).

A city file is 27 120 bytes of big-endian 16-bit words: resHist, comHist, indHist, crimeHist,
pollutionHist, moneyHist (240 words each), miscHist (120 words), then the 120 x 100 map (column-major,
x * 100 + y -- the engine's own order).  Scalars live in miscHist as the original game wrote them:
words 8-9 cityTime and 50-51 totalFunds (32-bit, high word first), 52 autoBulldoze, 53 autoBudget,
54 autoGoto, 55 enableSound, 56 cityTax, 57 simSpeed, 58-59 / 60-61 / 62-63 police / fire / road
funding as 16.16 fixed point.  (The reference's saveFile / loadFile move these through `long`, which is
8 bytes on LP64 and tramples the neighbouring words; this module writes and reads the 32-bit layout the
format was designed with, so files go to and from the original game.)

  save_env(engine, env, path)   device env -> file (for inspection in the original GUI)
  load_env(engine, env, path)   file -> device env: map, histories, funds, time, tax, auto flags
"""
import numpy as np

HIST = ("MPF_RES_HIST", "MPF_COM_HIST", "MPF_IND_HIST", "MPF_CRIME_HIST", "MPF_POLLUTION_HIST", "MPF_MONEY_HIST")
FILE_BYTES = 27120


def read_cty(path):
    raw = np.fromfile(path, dtype=">i2")
    if raw.size * 2 != FILE_BYTES:
        raise ValueError("%s: %d bytes, a city file has %d" % (path, raw.size * 2, FILE_BYTES))
    out = {name: raw[i * 240:(i + 1) * 240].astype(np.int16) for i, name in enumerate(HIST)}
    out["MPF_MISC_HIST"] = raw[1440:1560].astype(np.int16)
    out["MPF_MAP"] = raw[1560:].astype(np.int16).view(np.uint16)
    return out


def _get32(misc, i):
    return int((int(misc[i]) & 0xffff) << 16 | (int(misc[i + 1]) & 0xffff))


def _put32(misc, i, v):
    v = int(v) & 0xffffffff
    misc[i] = np.array(v >> 16, np.uint16).view(np.int16)
    misc[i + 1] = np.array(v & 0xffff, np.uint16).view(np.int16)


def decode_misc(misc):
    s32 = lambda v: v - (1 << 32) if v & 0x80000000 else v
    return {"cityTime": max(0, s32(_get32(misc, 8))), "totalFunds": s32(_get32(misc, 50)),
            "autoBulldoze": int(misc[52] != 0), "autoBudget": int(misc[53] != 0), "autoGoto": int(misc[54] != 0),
            "enableSound": int(misc[55] != 0), "cityTax": int(misc[56]), "simSpeed": int(misc[57]),
            "policePercent": s32(_get32(misc, 58)) / 65536.0, "firePercent": s32(_get32(misc, 60)) / 65536.0,
            "roadPercent": s32(_get32(misc, 62)) / 65536.0}


def write_cty(path, fields):
    words = np.concatenate([np.asarray(fields[n], np.int16).reshape(240) for n in HIST]
                           + [np.asarray(fields["MPF_MISC_HIST"], np.int16).reshape(120),
                              np.asarray(fields["MPF_MAP"]).astype(np.uint16).view(np.int16).reshape(12000)])
    words.astype(">i2").tofile(path)


def save_env(engine, env, path):
    """fileio.cpp:310-365 saveFile for env `env` of an EngineBatch."""
    f = {n: engine.get(env, n) for n in HIST + ("MPF_MISC_HIST", "MPF_MAP")}
    s = engine.scalars(env)
    misc = f["MPF_MISC_HIST"].copy()
    _put32(misc, 50, s.totalFunds)
    _put32(misc, 8, s.cityTime)
    misc[52], misc[53], misc[54], misc[55] = int(s.autoBulldoze), int(s.autoBudget), 0, 0
    misc[56], misc[57] = int(s.cityTax), int(s.simSpeed)
    _put32(misc, 58, int(s.policePercent * 65536))
    _put32(misc, 60, int(s.firePercent * 65536))
    _put32(misc, 62, int(s.roadPercent * 65536))
    f["MPF_MISC_HIST"] = misc
    write_cty(path, f)


def load_env(engine, env, path):
    """fileio.cpp:203-247 loadFileDir + the scalar decoding of loadFile :255-300 (tax and speed sanitised
    as there) for env `env`; the caller re-runs its reset tail (simTick) as after any map change."""
    f = read_cty(path)
    for n in HIST + ("MPF_MISC_HIST", "MPF_MAP"):
        engine.set(env, n, f[n])
    m = decode_misc(f["MPF_MISC_HIST"])
    s = engine.scalars(env)
    s.totalFunds, s.cityTime = m["totalFunds"], m["cityTime"]
    s.autoBulldoze, s.autoBudget = m["autoBulldoze"], m["autoBudget"]
    s.cityTax = m["cityTax"] if 0 <= m["cityTax"] <= 20 else 7
    engine.set_scalars(env, s)
    return m
