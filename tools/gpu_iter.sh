# quick iteration: identify parity tests, one bench line, one ncu capture of k_score
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "identify or golden" 2>&1 | tail -5
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'],'roof',d['roofline']['frac'],'kscore ms',d['roofline']['launch_ms'],d['stage_ms_per_step'])
PY
B="python bench.py --config c2 --spectra 2000 --steps 1 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-k_score} -s ${NCU_SKIP:-1} -c 1 -o gpurun_out/prof_iter $B > gpurun_out/ncu_iter.log 2>&1
tail -1 gpurun_out/ncu_iter.log
