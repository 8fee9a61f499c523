# round 2, final state: every GPU test, the default bench line, the reference arm, the other configs / decoy modes, ncu passes
set -x
mkdir -p gpurun_out
PER_TEST_TIMEOUT=420 bash tests/run_gpu_each.sh > gpurun_out/gpu_tests_summary.log 2>&1
grep -c "rc=0" gpurun_out/gpu_tests.log; grep -v "rc=0" gpurun_out/gpu_tests.log | head -40
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_r2_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref.json
python -c "import __graft_entry__ as g; g.smoke()"
bash tools/r2_breadth.sh
bash tools/r2_prof_final.sh
