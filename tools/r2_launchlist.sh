#!/bin/bash
# ncu launch list (gpu__time_duration only, serialised, cold caches) of two aged env steps of the headline workload
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv \
    --launch-skip ${1:-1150} --launch-count ${2:-60} python tools/r2_probe.py --age 40 --warm 2 --steps 3 --configs default= \
    > gpurun_out/r02_launches.log 2>&1
tail -2 gpurun_out/r02_launches.log | cut -c1-200
