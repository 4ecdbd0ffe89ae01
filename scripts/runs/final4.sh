set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t_final4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_final4.log
