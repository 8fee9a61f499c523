set -x
B="python bench.py --config ${CFG:-c2} --spectra 2000 --steps 1 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_decoy_random -s ${NCU_SKIP:-0} -c 1 -o gpurun_out/prof_decoy $B > gpurun_out/ncu_decoy.log 2>&1
tail -2 gpurun_out/ncu_decoy.log
