set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t4.log
python bench.py --frames 2048 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/b7_2048.log 2>&1
YSMR_FRONTEND=strip python bench.py --frames 2048 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/b7_2048_strip.log 2>&1
python bench.py --frames 2048 --steps 2 --warmup 3 --no-cpu --no-e2e --channels 1 > gpurun_out/b7_2048_c1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_v3.csv python bench.py --frames 512 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_v3.log 2>&1
