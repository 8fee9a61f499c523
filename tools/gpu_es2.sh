mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "decoys_random or identify_bit_exact or stored or passes" > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?"
tail -3 gpurun_out/t_parity.log
MD_DECOY_GENEROUS_MIN=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "decoys_random or identify_bit_exact or stored or passes" > gpurun_out/t_gen.log 2>&1; echo "generous rc=$?"
tail -3 gpurun_out/t_gen.log
for pct in ${PCTS:-230}; do
for cfg in ${CFGS:-c2}; do
MD_DECOY_GENEROUS_PCT=$pct timeout 300 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline --no-c4 > gpurun_out/b_$cfg.json 2> gpurun_out/b_$cfg.err; echo "$cfg rc=$?"
python -c "import json; d=json.load(open('gpurun_out/b_$cfg.json')); s=d['stage_ms_per_step']; print('$cfg pct=$pct', 'step ms', round(d['ms_per_step'],2), 'kdecoy', round(s['kernel_decoy_attempts'],2), 'decoys', round(s['decoys'],2), 'attempts', s['decoy_attempts'], 'kscore', round(d['roofline']['launch_ms'],3), 'e2e', round(d['e2e']['value']), 'crc', d['psm_crc'])"
done
done
MD_TRACE=1 timeout 600 python bench.py --config c2 --steps 1 --warmup 3 --no-cpu-baseline --no-c4 2>&1 >/dev/null | grep "md_trace" | tail -9 > gpurun_out/trace_c2.txt
cat gpurun_out/trace_c2.txt
