set -x
B="python bench.py --config c2 --steps 1 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_score -s 1 -c 1 -o gpurun_out/prof_score $B > gpurun_out/ncu_score.log 2>&1
tail -1 gpurun_out/ncu_score.log
