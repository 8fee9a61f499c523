# pipelined score kernel: parity first (each test node under its own timeout), then bench A/B against the classic kernel
set -x
mkdir -p gpurun_out
PER_TEST_TIMEOUT=420 bash tests/run_gpu_each.sh > gpurun_out/gpu_tests_summary.log 2>&1
grep -c "rc=0" gpurun_out/gpu_tests.log; grep -v "rc=0" gpurun_out/gpu_tests.log | head -60
for mode in pipe classic; do
  if [ $mode = classic ]; then export MD_SCORE_CLASSIC=1; else unset MD_SCORE_CLASSIC; fi
  for cfg in c2 c3; do
    timeout 600 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline --no-c4 > gpurun_out/b_${cfg}_$mode.json 2> gpurun_out/b_${cfg}_$mode.err; echo "$cfg $mode rc=$?"
    python -c "import json; d=json.load(open('gpurun_out/b_${cfg}_$mode.json')); s=d['stage_ms_per_step']; print('$cfg $mode', 'step ms', round(d['ms_per_step'],2), 'kdecoy', round(s['kernel_decoy_attempts'],2), 'decoys', round(s['decoys'],2), 'score', round(s['score'],3), 'kscore', round(d['roofline']['launch_ms'],3), 'frac', round(d['roofline']['frac'],4), 'crc', d['psm_crc'], 'e2e', round(d['e2e']['value']))"
  done
done
unset MD_SCORE_CLASSIC
timeout 300 python bench.py --config c5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b_c5.json 2> gpurun_out/b_c5.err; echo "c5 rc=$?"
python -c "import json; d=json.load(open('gpurun_out/b_c5.json')); print('c5', d['ms_per_step'], d['candidates_per_sec'], d['roofline']['frac'])"
