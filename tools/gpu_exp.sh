for w in 1 2 4; do
echo "== wide ctas $w"
MD_DECOY_WIDE_CTAS=$w timeout 300 python bench.py --config c2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err
python -c "import json; d=json.load(open('gpurun_out/bench_v.json')); s=d['stage_ms_per_step']; print('c2', 'step ms', round(d['ms_per_step'],2), 'kdecoy', round(s['kernel_decoy_attempts'],2), 'decoys', round(s['decoys'],2))"
done
