# round 2, first GPU pass: new tests, the default bench (C2 + C4 section), attempt floor
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/t_full.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/t_full.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_r2a.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2a.json'))
print('value',d['value'],'e2e',d['e2e'],'ms/step',d['ms_per_step'],'roof',d['roofline']['frac'],d['stage_ms_per_step'],'crc',d['psm_crc'])
print('c4',d.get('c4_strong')); print('cpu',d.get('cpu_baseline')); print(d['one_time'])
PY
timeout 300 python tools/attempt_stats.py 2>&1 | tail -8
