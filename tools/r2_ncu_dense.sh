#!/bin/bash
# by-function / by-line profile of k_mp_frames<8> on the dense-city workload
tag=r02dense
ncu --set full --clock-control none --import-source on -k regex:k_mp_frames --launch-skip ${1:-120} --launch-count 2 \
    -f -o /tmp/ncu_$tag python bench.py --workload mp120dense --steps 2 --warmup 3 --age-steps 8 --no-cpu-baseline --no-extra --rollout-steps 0 > gpurun_out/${tag}_ncu.log 2>&1
ncu -i /tmp/ncu_$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i /tmp/ncu_$tag.ncu-rep --page source --csv > /tmp/ncu_${tag}_src.csv 2>/dev/null
mkdir -p /tmp/cub_$tag && (cd /tmp/cub_$tag && cuobjdump -xelf all /root/repo/gymcity_b200/libgymcity_b200.so > /dev/null 2>&1)
cubin=$(ls /tmp/cub_$tag/*mp_batch*.cubin | head -1)
nvdisasm -g -c $cubin > /tmp/ncu_${tag}_dis.txt 2>/dev/null
python tools/ncu_by_function.py /tmp/ncu_${tag}_src.csv /tmp/ncu_${tag}_dis.txt k_mp_frames k_mp_framesILi8E > gpurun_out/${tag}_by_function.txt 2>&1
for col in "# Samples" "Instructions Executed"; do
    echo "==== top source lines by $col ====" >> gpurun_out/${tag}_by_line.txt
    python tools/ncu_by_line.py /tmp/ncu_${tag}_src.csv /tmp/ncu_${tag}_dis.txt k_mp_framesILi8E "$col" 50 >> gpurun_out/${tag}_by_line.txt 2>&1
done
python tools/r2_ncu_table.py gpurun_out/${tag}_raw.csv | cut -c1-420
head -32 gpurun_out/${tag}_by_function.txt
