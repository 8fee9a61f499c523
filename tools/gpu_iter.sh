set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?" 
tail -15 gpurun_out/t_parity.log
for cfg in ${CFGS:-c2 c3}; do
MD_SCORE_TIMING=1 timeout 300 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_$cfg.json 2> gpurun_out/b_$cfg.err; echo "$cfg rc=$?"
grep md_score_timing gpurun_out/b_$cfg.err | tail -1
python -c "import json; d=json.load(open('gpurun_out/b_$cfg.json')); s=d['stage_ms_per_step']; print('$cfg', 'step ms', round(d['ms_per_step'],2), 'kdecoy', round(s['kernel_decoy_attempts'],2), 'decoys', round(s['decoys'],2), 'kscore', round(d['roofline']['launch_ms'],3), 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']))"
done
