timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
MD_SCORE_TIMING=1 timeout 300 python bench.py --config c2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_c2.json 2> gpurun_out/b_c2.err
python -c "import json; d=json.load(open('gpurun_out/b_c2.json')); s=d['stage_ms_per_step']; print('c2', 'step ms', round(d['ms_per_step'],2), 'kdecoy', round(s['kernel_decoy_attempts'],2), 'decoys', round(s['decoys'],2), 'kscore', round(d['roofline']['launch_ms'],3), 'frac', round(d['roofline']['frac'],4))"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_score -s 1 -c 1 python bench.py --config c2 --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | grep -i "dram__\|gpu__time" 
