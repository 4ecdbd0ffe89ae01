set -x
timeout 900 python -m pytest tests/test_gpu_detect.py tests/test_gpu_e2e.py -m gpu -x -q > gpurun_out/t28.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t28.log
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/b30.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_v4.csv python bench.py --frames 1184 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_v4.log 2>&1
