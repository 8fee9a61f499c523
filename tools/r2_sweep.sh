mkdir -p gpurun_out
for later in 105 120 140; do for w0 in 110 118; do
  for cfg in c2 c1; do
    MD_DECOY_LATER_PCT=$later MD_DECOY_WANT0_PCT=$w0 timeout 600 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --no-c4 > gpurun_out/sw.json 2> gpurun_out/sw.err
    python -c "import json; d=json.load(open('gpurun_out/sw.json')); s=d['stage_ms_per_step']; print('later=$later w0=$w0 $cfg', round(d['value']), 'spectra/s', round(d['ms_per_step'],2), 'ms attempts', s['decoy_attempts'], d['psm_crc'])"
  done
done; done
