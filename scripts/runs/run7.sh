set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t8.log
python scripts/linkprof.py > gpurun_out/lp3.log 2>&1
python bench.py --frames 2048 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/b11_2048.log 2>&1
