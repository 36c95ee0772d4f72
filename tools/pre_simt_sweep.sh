for cfg in "0 0" "4 0" "4 1" "8 1" "16 1"; do set -- $cfg; echo "L=$1 zonekey=$2"; MPB_PRE_SIMT=$1 MPB_PRE_ZONE_KEY=$2 python tools/phase_times.py 2>&1 | tail -n 1; done
