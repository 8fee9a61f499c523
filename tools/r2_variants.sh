D=max-decoy_b200/csrc
cp $D/libmaxdecoy_cuda.so /tmp/orig.so
for v in $D/variants/lib_*.so; do
  cp $v $D/libmaxdecoy_cuda.so
  for cfg in ${CFGS:-c2 c3}; do
    MD_SCORE_TIMING=1 timeout 300 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline --no-c4 > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err
    grep md_score_timing gpurun_out/bench_v.err | tail -2
    timeout 300 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline --no-c4 > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err
    python -c "import json; d=json.load(open('gpurun_out/bench_v.json')); s=d['stage_ms_per_step']; print('$v $cfg', 'step ms', round(d['ms_per_step'],2), 'score', round(s['score'],3), 'kscore', round(d['roofline']['launch_ms'],3), 'frac', round(d['roofline']['frac'],4), 'crc', d['psm_crc'])"
  done
done
cp /tmp/orig.so $D/libmaxdecoy_cuda.so
