set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "identify or score or smoke or expanded or stored" 2>&1 | tail -2
for mode in pipe classic; do
  if [ $mode = classic ]; then export MD_SCORE_CLASSIC=1; else unset MD_SCORE_CLASSIC; fi
  for cfg in c2 c3; do
    timeout 600 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline --no-c4 > gpurun_out/b_${cfg}_$mode.json 2> gpurun_out/b_${cfg}_$mode.err; echo "$cfg $mode rc=$?"
    grep "md_score_timing" gpurun_out/b_${cfg}_$mode.err | tail -1
    python -c "import json; d=json.load(open('gpurun_out/b_${cfg}_$mode.json')); s=d['stage_ms_per_step']; print('$cfg $mode', 'step ms', round(d['ms_per_step'],2), 'score', round(s['score'],3), 'kscore', round(d['roofline']['launch_ms'],3), 'frac', round(d['roofline']['frac'],4), 'crc', d['psm_crc'])"
  done
done
