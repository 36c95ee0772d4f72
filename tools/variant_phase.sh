for so in gymcity_b200/libgymcity_b200_v_*.so; do n=$(basename $so .so); n=${n#libgymcity_b200_v_}; echo "== $n"; GYMCITY_B200_LIB=/root/repo/$so python tools/phase_times.py 2>&1 | tail -n 1; done
