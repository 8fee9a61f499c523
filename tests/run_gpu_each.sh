#!/bin/bash
# Runs every GPU test node separately under a hard timeout so a hung kernel costs one test, not the call.
mkdir -p gpurun_out
: > gpurun_out/gpu_tests.log
python -m pytest tests -m gpu --collect-only -q 2>/dev/null | grep '::' > gpurun_out/nodes.txt
while read -r node; do
  start=$(date +%s)
  timeout -s KILL ${PER_TEST_TIMEOUT:-420} python -m pytest "$node" -x -q -m gpu > gpurun_out/one.log 2>&1
  rc=$?
  echo "rc=$rc t=$(( $(date +%s) - start ))s $node" >> gpurun_out/gpu_tests.log
  if [ $rc -ne 0 ]; then tail -30 gpurun_out/one.log >> gpurun_out/gpu_tests.log; fi
done < gpurun_out/nodes.txt
cat gpurun_out/gpu_tests.log | tail -80
