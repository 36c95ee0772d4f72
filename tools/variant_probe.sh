# probe every gymcity_b200/libgymcity_b200_v_*.so with tools/r2_probe.py (one process each)
for so in gymcity_b200/libgymcity_b200_v_*.so; do n=$(basename $so .so); n=${n#libgymcity_b200_v_}; GYMCITY_B200_LIB=/root/repo/$so python tools/r2_probe.py --configs $n= 2>/dev/null | cut -c1-190; done
