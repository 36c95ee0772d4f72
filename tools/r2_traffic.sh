#!/bin/bash
# Per-kernel duration and DRAM traffic of the launches of aged env steps of the headline workload (ncu, serialised,
# cold caches): gpurun_out/<tag>_launches.csv, and <tag>_traffic.json = DRAM bytes per env-tick over one whole step.
tag=${1:-r02}; skip=${2:-400}; cnt=${3:-40}
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_launches.csv --launch-skip $skip --launch-count $cnt \
    python tools/r2_probe.py --age 40 --warm 2 --steps 3 --configs default= > gpurun_out/${tag}_launches.log 2>&1
python - "$tag" <<'PY'
import csv, json, sys
tag = sys.argv[1]
rows = [r for r in csv.reader(open("gpurun_out/%s_launches.csv" % tag)) if len(r) > 10 and r[0].isdigit()]
hdr = [r for r in csv.reader(open("gpurun_out/%s_launches.csv" % tag)) if r and r[0] == "ID"][0]
iu, im, iv, ik = hdr.index("Metric Unit"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Kernel Name")
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
launches = {}
for r in rows:
    d = launches.setdefault(int(r[0]), {"kernel": r[ik].split("(")[0].split("::")[-1]})
    d[r[im]] = float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
ids = sorted(launches)
# a step starts at the k_mp_sched_class launch that precedes a k_mp_env_pre (one-group steps rebuild the schedule a
# second time between the two simTicks)
pres = [i for i in ids if launches[i]["kernel"].startswith("k_mp_env_pre")]
def step_start(p):
    c = [i for i in ids if i < p and launches[i]["kernel"].startswith("k_mp_sched_class")]
    return max(c) if c else None
starts = [s for s in (step_start(p) for p in pres) if s is not None]
a = starts[0]
b = min(i for i in ids if i > a and launches[i]["kernel"].startswith("k_mp_env_post")) + 1      # ... through the step's k_mp_env_post
step = [launches[i] for i in ids if a <= i < b]
byk = {}
for l in step:
    k = byk.setdefault(l["kernel"], {"launches": 0, "ms": 0.0, "dram_bytes": 0.0})
    k["launches"] += 1
    k["ms"] += l.get("gpu__time_duration.sum", 0.0)
    k["dram_bytes"] += l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
tot_ms = sum(k["ms"] for k in byk.values()); tot_b = sum(k["dram_bytes"] for k in byk.values())
envs = 32768
out = {"envs": envs, "step_kernels": byk, "sum_kernel_ms_serialised": tot_ms, "dram_bytes_per_step": tot_b,
       "dram_bytes_per_env_tick": tot_b / envs,
       "share_of_step_time": {k: v["ms"] / tot_ms for k, v in byk.items()},
       "note": "ncu --metrics gpu__time_duration.sum,dram__bytes_* --clock-control none: one aged env step (from the schedule "
               "rebuild before k_mp_env_pre to the next one), kernels serialised and cold-cache, so shares are comparable, absolute times are not"}
json.dump(out, open("gpurun_out/%s_traffic.json" % tag, "w"), indent=1)
print(json.dumps({k: (v["launches"], round(v["ms"], 3), round(v["dram_bytes"] / 1e6, 1)) for k, v in byk.items()}))
print("dram bytes per env-tick %.0f, sum of kernel times %.2f ms" % (tot_b / envs, tot_ms))
PY
