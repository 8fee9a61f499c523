set -x
B="python bench.py --config c2 --steps 1 --warmup 3 --no-cpu-baseline --no-c4"
MD_TRACE=1 $B > gpurun_out/plain.log 2> gpurun_out/plain.err
grep "score" gpurun_out/plain.err | tail -4
ncu --set full --clock-control none --import-source on -k regex:k_score_pipe -s 1 -c 1 -o gpurun_out/prof_pipe $B > gpurun_out/ncu_pipe.log 2>&1
tail -2 gpurun_out/ncu_pipe.log
