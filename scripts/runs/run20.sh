set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t21.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t21.log
python scripts/linkprof.py > gpurun_out/lp6.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/b23.log 2>&1
