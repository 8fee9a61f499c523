# GPU parity tests (each node under its own timeout) + one bench line
bash tests/run_gpu_each.sh > gpurun_out/gpu_tests_summary.log 2>&1
grep -c "rc=0" gpurun_out/gpu_tests.log; grep -v "rc=0" gpurun_out/gpu_tests.log | head -40
timeout 600 python bench.py --steps 3 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cut -c1-1800 gpurun_out/bench.json
