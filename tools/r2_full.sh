set -x
mkdir -p gpurun_out
PER_TEST_TIMEOUT=420 bash tests/run_gpu_each.sh > gpurun_out/gpu_tests_summary.log 2>&1
grep -c "rc=0" gpurun_out/gpu_tests.log; grep -v "rc=0" gpurun_out/gpu_tests.log | head -40
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_r2_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_n1.json'))
print('value',d['value'],'e2e',d['e2e'],'ms/step',d['ms_per_step'],'roof',d['roofline']['frac'],d['roofline']['launch_ms'],d['stage_ms_per_step'],'crc',d['psm_crc'])
print('c4',d.get('c4_strong')); print('cpu',d.get('cpu_baseline')); print(d['k3'])
PY
python -c "import __graft_entry__ as g; g.smoke()"
