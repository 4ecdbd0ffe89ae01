set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t19.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t19.log
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/b22.log 2>&1
