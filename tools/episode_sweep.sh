run() { echo "== $*"; env "$@" python tools/episode_probe.py 2>&1 | grep "episodic steady"; }
run MPB_X=0
run MPB_FEPC=16 MPB_HEAVY_FEPC=16
run MPB_FEPC=16 MPB_HEAVY_FEPC=16 MPB_GROUPS=2
run MPB_FEPC=16 MPB_HEAVY_FEPC=16 MPB_GROUPS=1
run MPB_FRAMES_PER_LAUNCH=100
run MPB_GROUPS=2 MPB_FRAMES_PER_LAUNCH=100
