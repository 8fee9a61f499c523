set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "c3" 2>&1 | tail -3
for cfg in c3 c2; do
  timeout 600 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --no-c4 > gpurun_out/b_$cfg.json 2> gpurun_out/b_$cfg.err; echo "$cfg rc=$?"
  python -c "import json; d=json.load(open('gpurun_out/b_$cfg.json')); s=d['stage_ms_per_step']; print('$cfg', round(d['value']), 'spectra/s', round(d['ms_per_step'],2), 'ms', d['psm_crc'], s)"
done
