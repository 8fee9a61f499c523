set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for g in 4 1 2 8; do
  for cfg in c2 c3; do
    MD_DECOY_GROUPS=$g timeout 600 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline --no-c4 > gpurun_out/b_${cfg}_g$g.json 2> gpurun_out/b_${cfg}_g$g.err; echo "$cfg g=$g rc=$?"
    python -c "import json; d=json.load(open('gpurun_out/b_${cfg}_g$g.json')); s=d['stage_ms_per_step']; print('$cfg groups=$g', 'step ms', round(d['ms_per_step'],2), 'decoys', round(s['decoys'],2), 'krounds', round(s['kernel_decoy_attempts'],2), 'score', round(s['score'],3), 'attempts', s['decoy_attempts'], 'crc', d['psm_crc'], 'e2e', round(d['e2e']['value']))"
  done
done
