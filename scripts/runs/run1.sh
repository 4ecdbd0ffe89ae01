set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv
./scripts/ubench > gpurun_out/ubench.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t3.log
python bench.py --frames 2048 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/b6_2048.log 2>&1
python bench.py --frames 9000 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/b6_9000.log 2>&1
python bench.py --frames 9000 --steps 2 --warmup 3 --no-cpu --no-e2e --channels 1 > gpurun_out/b6_9000_c1.log 2>&1
