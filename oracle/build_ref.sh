#!/usr/bin/env bash
# build_ref.sh -- TEST INFRASTRUCTURE ONLY.
#
# Compiles the reference Micropolis engine from the sources where they lie under
# /root/reference (nothing is copied into the repo) together with oracle/ref_shim.cpp into
# oracle/_ref/libmicropolis_ref.so (git-ignored, travels to the GPU box with gpurun).
#
# Oracle pinning: ENG/evaluate.cpp:104 declares `short problemTable[PROBNUM]`
# but voteProblems (evaluate.cpp:278-299) reads index PROBNUM, which makes -O0/-O2 builds
# diverge.  The one declaration is padded to `[PROBNUM + 2] = {0}` by a sed into a temporary
# file that is deleted after compilation; every other source is compiled in place.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${MICROPOLIS_REF_SRC:-/root/reference/gym_city/envs/micropolis/MicropolisCore/src/MicropolisEngine/src}"
OUT="$HERE/_ref"
OPT="${MPREF_OPT:--O2}"
if [ ! -d "$REF" ]; then
    echo "build_ref.sh: reference sources not found at $REF (prebuilt $OUT is used as is)" >&2
    exit 0
fi
mkdir -p "$OUT"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
CXXFLAGS="$OPT -fPIC -w -std=gnu++14 -Dprivate=public -I$REF"
objs=()
for f in "$REF"/*.cpp; do
    b="$(basename "$f" .cpp)"
    src="$f"
    if [ "$b" = "evaluate" ]; then
        sed 's/short problemTable\[PROBNUM\];/short problemTable[PROBNUM + 2] = {0};/' "$f" > "$TMP/evaluate.cpp"
        grep -q 'problemTable\[PROBNUM + 2\] = {0}' "$TMP/evaluate.cpp" || { echo "pin failed" >&2; exit 1; }
        src="$TMP/evaluate.cpp"
    fi
    g++ $CXXFLAGS -c "$src" -o "$TMP/$b.o" &
    objs+=("$TMP/$b.o")
done
g++ $CXXFLAGS -c "$HERE/ref_shim.cpp" -o "$TMP/ref_shim.o" &
objs+=("$TMP/ref_shim.o")
wait
g++ -shared -o "$OUT/libmicropolis_ref${MPREF_SUFFIX:-}.so" "${objs[@]}"
echo "built $OUT/libmicropolis_ref${MPREF_SUFFIX:-}.so"
