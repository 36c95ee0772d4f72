"""Fixture generator (run in the build container, where /root/reference exists): the tile maps of the 24 city files
the reference ships (gym_city/envs/micropolis/MicropolisCore/cities/*.cty), read through the REFERENCE loader
(Micropolis::loadFileDir, fileio.cpp:203-247, via oracle/_ref) -> tests/golden/cty_maps.npz.  BASELINE config 4's
dense variant starts its envs from these maps; the GPU box has no /root/reference, so the maps travel as a fixture.

  python tests/golden/make_cty_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_engine import RefEngine  # noqa: E402

CITIES = "/root/reference/gym_city/envs/micropolis/MicropolisCore/cities"

names = sorted(n for n in os.listdir(CITIES) if n.endswith(".cty"))
maps = []
for n in names:
    r = RefEngine()
    assert r.loadFileDir(os.path.join(CITIES, n)), n
    maps.append(np.asarray(r.get("MPF_MAP")).astype(np.uint16).copy())
out = os.path.join(ROOT, "tests", "golden", "cty_maps.npz")
np.savez_compressed(out, names=np.array(names), maps=np.stack(maps))
print(out, len(names), os.path.getsize(out))
