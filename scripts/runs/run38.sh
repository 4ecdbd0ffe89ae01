set -x
timeout 600 python -m pytest tests/test_gpu_dropin.py -m gpu -x -q > gpurun_out/t38.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t38.log
