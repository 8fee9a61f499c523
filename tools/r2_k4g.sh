set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "identify or score or smoke or expanded or stored" 2>&1 | tail -2
for cfg in c2 c3 c1; do
    timeout 600 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline --no-c4 > gpurun_out/b_${cfg}_pipe.json 2> gpurun_out/b_${cfg}_pipe.err; echo "$cfg rc=$?"
    python -c "import json; d=json.load(open('gpurun_out/b_${cfg}_pipe.json')); s=d['stage_ms_per_step']; print('$cfg pipe', 'step ms', round(d['ms_per_step'],2), 'score', round(s['score'],3), 'kscore', round(d['roofline']['launch_ms'],3), 'frac', round(d['roofline']['frac'],4), 'crc', d['psm_crc'])"
done
MD_TRACE=1 timeout 600 python bench.py --config c2 --steps 2 --warmup 3 --no-cpu-baseline --no-c4 2>&1 >/dev/null | grep "score:" | tail -2
