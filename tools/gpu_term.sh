set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_terminal_mods.py -x -q -m gpu > gpurun_out/t_term.log 2>&1; echo "term rc=$?"
tail -25 gpurun_out/t_term.log
bash tools/gpu_iter.sh
