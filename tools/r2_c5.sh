set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "kernel_paths or identify" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "c5" 2>&1 | tail -3
for cfg in c5; do
  timeout 600 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --no-c4 > gpurun_out/b_$cfg.json 2> gpurun_out/b_$cfg.err; echo "$cfg rc=$?"
  python -c "import json; d=json.load(open('gpurun_out/b_$cfg.json')); s=d['stage_ms_per_step']; print('$cfg', round(d['value']), 'spectra/s', round(d['ms_per_step'],2), 'ms', 'pairs/s', round(d['candidates_per_sec']), 'kscore', round(d['roofline']['launch_ms'],3), 'frac', round(d['roofline']['frac'],4), d['psm_crc'], s)"
done
MD_SCORE_SPLIT_CLASSIC=1 timeout 600 python bench.py --config c5 --steps 10 --warmup 3 --no-cpu-baseline --no-c4 > gpurun_out/b_c5_classic.json 2> gpurun_out/b_c5_classic.err
python -c "import json; d=json.load(open('gpurun_out/b_c5_classic.json')); s=d['stage_ms_per_step']; print('c5 classic', round(d['value']), 'spectra/s', round(d['ms_per_step'],2), 'ms', 'kscore', round(d['roofline']['launch_ms'],3), d['psm_crc'])"
