python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/b42.log 2>&1
