for pct in 100 108 115; do
echo "== later $pct"
for i in 1 2; do MD_DECOY_LATER_PCT=$pct timeout 300 python bench.py --config c2 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err
python -c "import json; d=json.load(open('gpurun_out/bench_v.json')); s=d['stage_ms_per_step']; print('step', round(d['ms_per_step'],2), 'decoys', round(s['decoys'],2), 'kdecoy', round(s['kernel_decoy_attempts'],2), 'attempts', s['decoy_attempts'], 'launches', d['gpu_launches']//4)"; done
done
