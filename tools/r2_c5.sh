for i in 1 2; do
timeout 300 python bench.py --config c5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_c5.json 2> gpurun_out/b_c5.err; echo "c5 rc=$?"
python -c "import json; d=json.load(open('gpurun_out/b_c5.json')); s=d['stage_ms_per_step']; print('c5 step', d['ms_per_step'], 'lookup', s['lookup'], 'score', s['score'], 'kscore', s['kernel_score'], d['candidates_per_sec'], d['roofline']['frac'])"
done
