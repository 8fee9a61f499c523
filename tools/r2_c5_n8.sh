# BASELINE configs[4] as stated: the +-500 Da open search on 8 GPUs (64 spectra x ~630k targets per rank), PSM tables gathered
set -x
mkdir -p gpurun_out
N=${N:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --config c5 --steps 20 --warmup 5 --no-cpu-baseline --no-c4 > gpurun_out/bench_r2_c5_n$N.json 2> gpurun_out/bench_r2_c5_n$N.err; echo "c5 n$N rc=$?"
tail -3 gpurun_out/bench_r2_c5_n$N.err
python -c "import json; d=json.load(open('gpurun_out/bench_r2_c5_n$N.json')); print(round(d['value']), 'spectra/s', round(d['ms_per_step'],2), 'ms', 'pairs/s', d['candidates_per_sec'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'crc', d['psm_crc'], d['stage_ms_per_step'])"
