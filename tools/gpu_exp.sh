MD_TRACE=1 timeout 300 python bench.py --config c2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_trace.json 2> gpurun_out/bench_trace.err
grep "decoy round" gpurun_out/bench_trace.err | tail -12
CFGS=c2 bash tools/gpu_variants.sh
