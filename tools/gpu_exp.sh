timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "decoys or identify_bit or passes" 2>&1 | tail -2
for i in 1 2; do timeout 300 python bench.py --config c2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err
python -c "import json; d=json.load(open('gpurun_out/bench_v.json')); s=d['stage_ms_per_step']; print('step', round(d['ms_per_step'],2), 'decoys', round(s['decoys'],2), 'kdecoy', round(s['kernel_decoy_attempts'],2), 'overhead', round(s['decoys']-s['kernel_decoy_attempts'],2), 'e2e', round(d['e2e']['value']))"; done
