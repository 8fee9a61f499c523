# round 2 profiling pass: launch list + full captures of the two hot kernels on the bench command itself
set -x
B="python bench.py --config c2 --steps 1 --warmup 3 --no-cpu-baseline --no-c4"
$B > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_score_pipe -s 1 -c 1 -o gpurun_out/prof_pipe $B > gpurun_out/ncu_pipe.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_decoy_random -s 0 -c 1 -o gpurun_out/prof_decoy $B > gpurun_out/ncu_decoy.log 2>&1
ncu --set full --clock-control none -k regex:k_build_tables -s 1 -c 1 -o gpurun_out/prof_tables $B > gpurun_out/ncu_tables.log 2>&1
tail -2 gpurun_out/plain.log | cut -c1-300
