#!/bin/bash
# AddressSanitizer + UndefinedBehaviorSanitizer over (1) the engine / gym headers the CUDA kernels are built from,
# compiled for one host thread (tests/hostbuild/host_engine.cpp), and (2) the reference engine + oracle shim, while
# the CPU parity tests drive both.  (compute-sanitizer is closed on the GPU pool, so the device build cannot be run
# under memcheck / racecheck; the serial-lane logic is the same source in both builds.)
#   tools/sanitize_host.sh > profiles/r02_asan_ubsan_host.txt 2>&1
set -u
cd "$(dirname "$0")/.."
SAN="-O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer"
mkdir -p tests/_hostbuild
g++ $SAN -fPIC -shared -std=c++17 -o tests/_hostbuild/libmph_asan.so tests/hostbuild/host_engine.cpp || exit 1
MPREF_OPT="$SAN" MPREF_SUFFIX=_asan bash oracle/build_ref.sh || exit 1
export MPH_LIB=$PWD/tests/_hostbuild/libmph_asan.so MPREF_LIB=$PWD/oracle/_ref/libmicropolis_ref_asan.so
export LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)"
export ASAN_OPTIONS=detect_leaks=0:abort_on_error=0:halt_on_error=0 UBSAN_OPTIONS=print_stacktrace=1
python -m pytest -q -x -p no:cacheprovider tests/test_mp_engine_host.py tests/test_gym_oracle.py tests/test_extinguisher.py \
    tests/test_paint_env.py tests/test_mp_golden.py -m "not gpu" > /tmp/mp_sanitize_host.log 2>&1
grep -v "^$" /tmp/mp_sanitize_host.log | tail -40
echo "sanitizer reports (runtime error / AddressSanitizer): $(grep -c 'runtime error\|ERROR: AddressSanitizer' /tmp/mp_sanitize_host.log)"
echo "libraries under test: $MPH_LIB $MPREF_LIB"
