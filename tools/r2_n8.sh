set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
N=${N:-8}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err; echo "bench n$N rc=$?"
tail -4 gpurun_out/bench_r2_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2_n$N.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'],'crc',d['psm_crc'],'rows',d['psm_rows'], 'clocks', d['clocks'])
print('c4',d.get('c4_strong'))
PY
