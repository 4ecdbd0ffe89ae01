set -x
timeout 900 python -m pytest tests/test_gpu_detect.py -m gpu -x -q > gpurun_out/t24.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t24.log
python bench.py --frames 2368 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/b26.log 2>&1
python bench.py --frames 2368 --steps 3 --warmup 3 --no-cpu --no-e2e --channels 1 > gpurun_out/b26_c1.log 2>&1
