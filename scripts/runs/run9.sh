set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t10.log
python scripts/linkprof.py > gpurun_out/lp4.log 2>&1
python bench.py --frames 2048 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/b13_2048.log 2>&1
python bench.py --frames 2368 --batch 592 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/b13_b592.log 2>&1
python bench.py --frames 2048 --batch 512 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/b13_b512.log 2>&1
