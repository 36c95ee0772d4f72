for cfg in "128 4" "64 4" "256 4" "128 3" "128 5" "128 6" "64 6" "256 3" "32 4"; do set -- $cfg; echo "T=$1 L=$2"; MPB_PRE_T=$1 MPB_PRE_L=$2 python tools/phase_times.py 2>&1 | tail -n 1; done
