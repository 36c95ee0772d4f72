: > gpurun_out/s9.jsonl
for wl in mp120 mp32puzzle mp16; do
for f in 0 1 2; do
  export MPB_PRE_SIMT=$f
  echo "{\"wl\":\"$wl\",\"pre_simt\":\"$f\"}" >> gpurun_out/s9.jsonl
  python bench.py --workload $wl --no-extra --no-cpu-baseline --steps 20 --warmup 5 2>>gpurun_out/s9_err.log | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except: continue
    print(json.dumps({k:d.get(k) for k in ('value','ms_per_step','step_ms','reward_checksum','flags')}))
" >> gpurun_out/s9.jsonl
done; done
unset MPB_PRE_SIMT
for f in 0 1 2; do echo "episode pre_simt=$f" >> gpurun_out/s9.jsonl; MPB_PRE_SIMT=$f python tools/episode_probe.py 2>&1 | tail -n 3 >> gpurun_out/s9.jsonl; done
cat gpurun_out/s9.jsonl
