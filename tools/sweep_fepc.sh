set -x
python -m pytest tests/test_vec_env.py -m gpu -x -q > gpurun_out/s5_vec.log 2>&1
: > gpurun_out/s5_sweep.jsonl
for wl in mp120 mp32puzzle; do
for f in default 1 8 16 32; do
  if [ $f = default ]; then unset MPB_FEPC MPB_HEAVY_FEPC; else export MPB_FEPC=$f MPB_HEAVY_FEPC=$f; fi
  echo "{\"wl\":\"$wl\",\"fepc\":\"$f\"}" >> gpurun_out/s5_sweep.jsonl
  python bench.py --workload $wl --no-extra --no-cpu-baseline --steps 20 --warmup 5 2>>gpurun_out/s5_err.log | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except: continue
    print(json.dumps({k:d.get(k) for k in ('value','ms_per_step','step_ms','reward_checksum','flags')}))
" >> gpurun_out/s5_sweep.jsonl
done; done
cat gpurun_out/s5_vec.log | tail -3; cat gpurun_out/s5_sweep.jsonl
