"""Regenerates the function list of gymcity_b200/csrc/mp_cov.h from the MP_COV(name) markers in the engine / gym
headers (every function definition there starts with one; add the marker to a new function, then run this)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "gymcity_b200", "csrc")
names = []
for f in ("mp_sim.h", "mp_sprites.h", "mp_tools.h", "mp_gym.h"):
    for m in re.finditer(r"^\s*MP_COV\((\w+)\);", open(os.path.join(CSRC, f)).read(), re.M):
        if m.group(1) not in names:
            names.append(m.group(1))
p = os.path.join(CSRC, "mp_cov.h")
s = open(p).read()
a = s.index("#define MP_COV_LIST(X) \\\n") + len("#define MP_COV_LIST(X) \\\n")
b = s.index("\nnamespace mp {")
rows = ["    " + " ".join("X(%s)" % n for n in names[k:k + 6]) for k in range(0, len(names), 6)]
open(p, "w").write(s[:a] + " \\\n".join(rows) + "\n" + s[b:])
print(len(names), "functions")
