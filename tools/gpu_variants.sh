# A/B of library variants (max-decoy_b200/csrc/variants/lib_*.so) on the decoy parity tests + a short C2 bench
D=max-decoy_b200/csrc
cp $D/libmaxdecoy_cuda.so /tmp/orig.so
run() {
  timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "${TESTK:-decoys_random}" 2>&1 | tail -1
  for cfg in ${CFGS:-c2 c3}; do
    timeout 300 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err
    python -c "import json; d=json.load(open('gpurun_out/bench_v.json')); s=d['stage_ms_per_step']; print('$cfg', 'step ms', round(d['ms_per_step'],2), 'kdecoy', round(s['kernel_decoy_attempts'],2), 'decoys', round(s['decoys'],2), 'kscore', round(d['roofline']['launch_ms'],3), 'frac', round(d['roofline']['frac'],4))"
  done
}
echo "== default"; run
for v in $D/variants/lib_*.so; do
  cp $v $D/libmaxdecoy_cuda.so
  echo "== $v"; run
done
cp /tmp/orig.so $D/libmaxdecoy_cuda.so
