D=max-decoy_b200/csrc
cp $D/libmaxdecoy_cuda.so /tmp/orig.so
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
run() { for i in 1 2 3; do timeout 300 python bench.py --config c2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err
python -c "import json; d=json.load(open('gpurun_out/bench_v.json')); print('kscore', round(d['roofline']['launch_ms'],4), 'step', round(d['ms_per_step'],2))"; done; }
echo "== new"; run
cp $D/variants/lib_head.so $D/libmaxdecoy_cuda.so; echo "== head"; run
cp /tmp/orig.so $D/libmaxdecoy_cuda.so
