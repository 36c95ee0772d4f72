#!/bin/bash
# k_mp_env_pre of an aged step under ncu --set full with source counters -> hottest lines / functions
tag=${1:-r02pre}; skip=${2:-42}
ncu --set full --clock-control none --import-source on -k regex:k_mp_env_pre --launch-skip $skip --launch-count 1 \
    -f -o /tmp/ncu_$tag python tools/r2_probe.py --age 40 --warm 2 --steps 1 --configs default= > gpurun_out/${tag}_ncu.log 2>&1
ncu -i /tmp/ncu_$tag.ncu-rep --page details > gpurun_out/${tag}_ncu_full.txt 2>/dev/null
ncu -i /tmp/ncu_$tag.ncu-rep --page source --csv > /tmp/ncu_${tag}_src.csv 2>/dev/null
mkdir -p /tmp/cub_$tag && (cd /tmp/cub_$tag && cuobjdump -xelf all /root/repo/gymcity_b200/libgymcity_b200.so > /dev/null 2>&1)
cubin=$(ls /tmp/cub_$tag/*mp_batch*.cubin | head -1)
nvdisasm -g -c $cubin > /tmp/ncu_${tag}_dis.txt 2>/dev/null
for col in "# Samples" "Instructions Executed"; do
    echo "==== top source lines by $col ====" >> gpurun_out/${tag}_by_line.txt
    python tools/ncu_by_line.py /tmp/ncu_${tag}_src.csv /tmp/ncu_${tag}_dis.txt k_mp_env_pre "$col" 60 >> gpurun_out/${tag}_by_line.txt 2>&1
done
python tools/ncu_by_function.py /tmp/ncu_${tag}_src.csv /tmp/ncu_${tag}_dis.txt k_mp_env_pre k_mp_env_pre > gpurun_out/${tag}_by_function.txt 2>&1
head -40 gpurun_out/${tag}_by_function.txt; grep -n "Instructions Executed" -A50 gpurun_out/${tag}_by_line.txt | head -64
