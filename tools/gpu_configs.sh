# one bench line per parity configuration (C1, C3, C5) for the record; C2 is the default bench
for cfg in c1 c3 c5; do
  timeout 400 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err; echo "$cfg rc=$?"
  python -c "import json; d=json.load(open('gpurun_out/bench_$cfg.json')); s=d['stage_ms_per_step']; print('$cfg', round(d['value']), 'spectra/s', round(d['ms_per_step'],2), 'ms', 'pairs/s', round(d['candidates_per_sec']), s)"
done
