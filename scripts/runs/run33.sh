set -x
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_dropin.py -m gpu -x -q > gpurun_out/t33.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t33.log
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/b33.log 2>&1
YSMR_NO_SPLIT=1 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/b33n.log 2>&1
