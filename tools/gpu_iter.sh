set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?" 
tail -15 gpurun_out/t_parity.log
MD_SCORE_TIMING=1 timeout 300 python bench.py --config c2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_c2.json 2> gpurun_out/b_c2.err; echo "c2 rc=$?"
grep md_score_timing gpurun_out/b_c2.err | tail -1
timeout 300 python bench.py --config c3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_c3.json 2> gpurun_out/b_c3.err; echo "c3 rc=$?"
MD_DECOY_WIDE_ONLY=1 timeout 300 python bench.py --config c2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_c2_wide.json 2> gpurun_out/b_c2_wide.err; echo "c2w rc=$?"
python - <<'P'
import json
for f in ['b_c2','b_c3','b_c2_wide']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['stage_ms_per_step'], d['roofline']['frac'])
    except Exception as e: print(f, 'ERR', e)
P
