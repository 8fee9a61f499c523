for c in c1 c3 c5; do
  timeout 900 python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "$c rc=$?"; tail -2 gpurun_out/bench_$c.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_$c.json'))
print('$c','value',d['value'],'cand/s',d['candidates_per_sec'],'ms/step',d['ms_per_step'],'roof',d['roofline']['frac'],'pairs',d['pairs_per_step'],d['stage_ms_per_step'])
PY
done
