B="python bench.py --config c2 --steps 1 --warmup 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_x.csv $B > gpurun_out/ncu_launch_x.log 2>&1
MD_TRACE=1 $B 2>&1 | grep "md_trace" | tail -3 | cut -c1-600
