#!/bin/bash
# Frame launches of an aged 32768-env batch under ncu --set full, for a scheduler configuration; the report is
# turned into text (details page + raw csv) on the box and deleted (it is ~8 MB per launch).
#   tools/r2_ncu.sh <name> <config string for r2_probe.py> [launch-skip] [launch-count]
name=$1; cfg=$2; skip=${3:-250}; cnt=${4:-10}
ncu --set full --clock-control none --import-source off -k regex:k_mp_frames --launch-skip $skip --launch-count $cnt \
    -f -o /tmp/r2_ncu_$name python tools/r2_probe.py --age 30 --warm 1 --steps 1 --configs "$name=$cfg" \
    > gpurun_out/r2_ncu_$name.log 2>&1
ncu -i /tmp/r2_ncu_$name.ncu-rep --page details > gpurun_out/r2_ncu_$name.details.txt 2>/dev/null
ncu -i /tmp/r2_ncu_$name.ncu-rep --page raw --csv > gpurun_out/r2_ncu_$name.raw.csv 2>/dev/null
rm -f /tmp/r2_ncu_$name.ncu-rep
tail -1 gpurun_out/r2_ncu_$name.log | cut -c1-200
