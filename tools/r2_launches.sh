set -x
B="python bench.py --config ${CFG:-c2} --steps 1 --warmup 3 --no-cpu-baseline --no-c4"
$B > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/plain.log | cut -c1-200
