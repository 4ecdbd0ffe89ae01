set -x
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_dropin.py -m gpu -x -q > gpurun_out/t40.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t40.log
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/b40.log 2>&1
YSMR_NO_POSTSPLIT=1 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/b40n.log 2>&1
