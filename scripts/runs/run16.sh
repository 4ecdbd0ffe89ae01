set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t17.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t17.log
python bench.py --frames 2368 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/b20.log 2>&1
