set -x
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_dropin.py tests/test_gpu_link.py -m gpu -x -q > gpurun_out/t37.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t37.log
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/b37.log 2>&1
