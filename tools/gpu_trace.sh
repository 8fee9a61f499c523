mkdir -p gpurun_out
MD_TRACE=1 timeout 600 python bench.py --config c2 --steps 1 --warmup 3 --no-cpu-baseline --no-c4 2>&1 >/dev/null | grep "md_trace" | tail -60 > gpurun_out/trace_c2.txt
tail -45 gpurun_out/trace_c2.txt
