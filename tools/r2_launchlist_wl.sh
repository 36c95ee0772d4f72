#!/bin/bash
# ncu launch list (durations, serialised) of two aged steps of a bench workload: tools/r2_launchlist_wl.sh mp120 <skip> <count>
wl=${1:-mp120}; skip=${2:-600}; cnt=${3:-60}
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_$wl.csv \
    --launch-skip $skip --launch-count $cnt python bench.py --workload $wl --steps 3 --warmup 3 --age-steps 30 --no-cpu-baseline --no-extra --rollout-steps 0 \
    > gpurun_out/r02_launches_$wl.log 2>&1
python - "$wl" <<'PY'
import csv, sys, collections
wl = sys.argv[1]
rows = [r for r in csv.reader(open("gpurun_out/r02_launches_%s.csv" % wl)) if len(r) > 10 and r[0].isdigit()]
names = [(r[4].split("(")[0].split("::")[-1], float(r[-1].replace(",", ""))) for r in rows]
idx = [i for i, (n, _) in enumerate(names) if n.startswith("k_mp_sched_class")]
a, b = idx[0], idx[1]
agg = collections.OrderedDict(); tot = 0
for n, t in names[a:b]:
    agg[n] = agg.get(n, [0, 0.0]); agg[n][0] += 1; agg[n][1] += t; tot += t
for n, (c, t) in agg.items():
    print("%-34s x%-3d %8.3f ms %5.1f %%" % (n[:34], c, t / 1e6, 100 * t / tot))
print("sum %.3f ms" % (tot / 1e6))
PY
