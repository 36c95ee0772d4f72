#!/bin/bash
# usage: tools/sweep.sh "MPB_GROUPS=8,MPB_FEPC=16" "MPB_FRAMES_PER_LAUNCH=20" ...   (one bench run per argument)
for cfg in "$@"; do
  r=$(env $(echo $cfg | tr ',' ' ') python bench.py --steps ${SWEEP_STEPS:-8} --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.0f e2e %.0f' % (d['value'], d['e2e']['value']))")
  echo "$cfg -> $r"
done
