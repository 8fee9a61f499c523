timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "decoy or identify" 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'],'roof',d['roofline']['frac'],d['stage_ms_per_step'])
PY
