set -x
PER_TEST_TIMEOUT=420 bash tests/run_gpu_each.sh > gpurun_out/gpu_tests_summary.log 2>&1
grep -c "rc=0" gpurun_out/gpu_tests.log; grep -v "rc=0" gpurun_out/gpu_tests.log | head -40
for cfg in c2 c3 c1 c5; do
    timeout 600 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline --no-c4 > gpurun_out/b_${cfg}_pipe.json 2> gpurun_out/b_${cfg}_pipe.err; echo "$cfg rc=$?"
    python -c "import json; d=json.load(open('gpurun_out/b_${cfg}_pipe.json')); s=d['stage_ms_per_step']; print('$cfg pipe', 'step ms', round(d['ms_per_step'],2), 'lookup', round(s['lookup'],3), 'decoys', round(s['decoys'],2), 'score', round(s['score'],3), 'kscore', round(d['roofline']['launch_ms'],3), 'frac', round(d['roofline']['frac'],4), 'crc', d['psm_crc'], 'pairs/s', round(d['candidates_per_sec']))"
done
