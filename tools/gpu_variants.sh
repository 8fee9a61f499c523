D=max-decoy_b200/csrc
cp $D/libmaxdecoy_cuda.so /tmp/orig.so
for v in $D/variants/lib_*.so; do
  cp $v $D/libmaxdecoy_cuda.so
  echo "== $v"
  timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "identify" 2>&1 | tail -1
  MD_SCORE_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err; tail -1 gpurun_out/bench_v.err
  python -c "import json; d=json.load(open('gpurun_out/bench_v.json')); print('kscore ms', d['roofline']['launch_ms'], 'frac', d['roofline']['frac'])"
done
cp /tmp/orig.so $D/libmaxdecoy_cuda.so
