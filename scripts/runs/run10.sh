set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t11.log
python bench.py --frames 2368 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/b14.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches_v3f.csv python bench.py --frames 1184 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_v3f.log 2>&1
