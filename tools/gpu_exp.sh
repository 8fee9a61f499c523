MD_SCORE_TIMING=1 timeout 300 python bench.py --config c2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_c2.json 2> gpurun_out/b_c2.err
grep md_score_timing gpurun_out/b_c2.err | tail -1
