"""Which engine functions do the GPU workloads actually enter?  The -DMP_COVERAGE build of the library
(gymcity_b200/libgymcity_b200_cov.so, built by __graft_entry__.build(); csrc/mp_cov.h) sets one bit per engine / gym
function when an env enters it; tools/coverage_run.py drives the parity suite's workloads against that build in a
subprocess and reports the bitmap.  Every function of the device build must have been entered, except the ones listed
here with the reason."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COV_LIB = os.path.join(ROOT, "gymcity_b200", "libgymcity_b200_cov.so")

# Functions the DEVICE build never calls by construction: the serial terrain generator (the device runs
# generate_map_coop; the serial one is the CPU harness's and is diffed against the reference there), the serial
# earthquake (the device runs make_earthquake_coop), the per-cell traffic decay (the device decays four cells per
# 32-bit word), and the whole-env single-thread drivers of the CPU harness.
HOST_ONLY = {"put_on_map", "plop_b_river", "plop_s_river", "do_river", "smooth_river", "smooth_trees", "do_trees",
             "make_naked_island", "generate_map", "make_earthquake", "dec_traffic_cell", "gym_step", "gym_reset_finish",
             "sim_loop", "sim_tick", "control_sim_tick"}


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(COV_LIB), reason="coverage build not present (python -m gymcity_b200.build --coverage)")
def test_every_device_function_is_entered_by_the_gpu_workloads():
    env = dict(os.environ, GYMCITY_B200_LIB=COV_LIB)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "coverage_run.py")], env=env, cwd=ROOT,
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    rep = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert rep["count"] > 150, rep
    missed = set(rep["missed"]) - HOST_ONLY
    with open(os.path.join(ROOT, "gpurun_out", "r02_coverage.json") if os.path.isdir(os.path.join(ROOT, "gpurun_out"))
              else os.devnull, "w") as f:
        json.dump(rep, f)
    assert not missed, sorted(missed)
