set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "kernel_paths" 2>&1 | tail -3
for cfg in c1 c3 c5; do
  timeout 600 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_$cfg.json 2> gpurun_out/bench_r2_$cfg.err; echo "$cfg rc=$?"
  python -c "import json; d=json.load(open('gpurun_out/bench_r2_$cfg.json')); s=d['stage_ms_per_step']; print('$cfg', round(d['value']), 'spectra/s', round(d['ms_per_step'],2), 'ms', 'pairs/s', round(d['candidates_per_sec']), 'frac', round(d['roofline']['frac'],4), s)"
done
timeout 600 python bench.py --config c2 --decoy-mode permute --steps 10 --warmup 3 --no-cpu-baseline --no-c4 > gpurun_out/bench_r2_c2_permute.json 2> gpurun_out/bench_r2_c2_permute.err; echo "permute rc=$?"; tail -2 gpurun_out/bench_r2_c2_permute.err
python -c "import json; d=json.load(open('gpurun_out/bench_r2_c2_permute.json')); s=d['stage_ms_per_step']; print('permute', round(d['value']), 'spectra/s', round(d['ms_per_step'],2), 'ms', s)"
timeout 600 python bench.py --config c1 --decoy-mode exhaustive --decoys 100 --steps 5 --warmup 3 --no-cpu-baseline --no-c4 > gpurun_out/bench_r2_c1_exhaustive.json 2> gpurun_out/bench_r2_c1_exhaustive.err; echo "exhaustive rc=$?"; tail -2 gpurun_out/bench_r2_c1_exhaustive.err
python -c "import json; d=json.load(open('gpurun_out/bench_r2_c1_exhaustive.json')); s=d['stage_ms_per_step']; print('exhaustive', round(d['value']), 'spectra/s', round(d['ms_per_step'],2), 'ms', s)"
