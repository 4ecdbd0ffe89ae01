set -x
bash scripts/runs/profiles.sh
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_final.log
( time python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
