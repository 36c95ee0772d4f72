#!/bin/bash
# One k_mp_frames<32> launch (50 frames of ~16 000 aged cities) under ncu --set full with source counters; the report
# is reduced on the box to text: details page, raw metrics, and the hottest source lines by samples / instructions
# executed / selected stalls (tools/ncu_by_line.py over the product cubin's line table).
#   tools/r2_ncu_source.sh <tag> [launch-skip] [launch-count]
tag=${1:-r02}; skip=${2:-830}; cnt=${3:-2}
ncu --set full --clock-control none --import-source on -k regex:k_mp_frames --launch-skip $skip --launch-count $cnt \
    -f -o /tmp/ncu_$tag python tools/r2_probe.py --age 40 --warm 2 --steps 1 --configs default= > gpurun_out/${tag}_ncu.log 2>&1
ncu -i /tmp/ncu_$tag.ncu-rep --page details > gpurun_out/${tag}_k_mp_frames_ncu_full.txt 2>/dev/null
ncu -i /tmp/ncu_$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_k_mp_frames_ncu_raw.csv 2>/dev/null
ncu -i /tmp/ncu_$tag.ncu-rep --page source --csv > /tmp/ncu_${tag}_src.csv 2>/dev/null
mkdir -p /tmp/cub_$tag && (cd /tmp/cub_$tag && cuobjdump -xelf all /root/repo/gymcity_b200/libgymcity_b200.so > /dev/null 2>&1)
cubin=$(ls /tmp/cub_$tag/*mp_batch*.cubin | head -1)
nvdisasm -g -c $cubin > /tmp/ncu_${tag}_dis.txt 2>/dev/null
for col in "# Samples" "Instructions Executed" "stall_no_inst" "stall_long_sb" "stall_barrier"; do
    echo "==== top source lines by $col ====" >> gpurun_out/${tag}_k_mp_frames_by_line.txt
    python tools/ncu_by_line.py /tmp/ncu_${tag}_src.csv /tmp/ncu_${tag}_dis.txt k_mp_framesILi32E "$col" 90 >> gpurun_out/${tag}_k_mp_frames_by_line.txt 2>&1
done
python tools/ncu_by_function.py /tmp/ncu_${tag}_src.csv /tmp/ncu_${tag}_dis.txt k_mp_frames k_mp_framesILi32E > gpurun_out/${tag}_k_mp_frames_by_function.txt 2>&1
ls -la /tmp/ncu_$tag.ncu-rep /tmp/ncu_${tag}_src.csv | awk '{print $5, $9}'
head -30 gpurun_out/${tag}_k_mp_frames_by_line.txt
python tools/ncu_calls.py /tmp/ncu_${tag}_src.csv /root/repo/gymcity_b200/libgymcity_b200.so k_mp_framesILi32ELb0 60 > gpurun_out/${tag}_k_mp_frames_calls.txt 2>&1
head -40 gpurun_out/${tag}_k_mp_frames_calls.txt
