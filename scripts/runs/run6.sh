set -x
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dp_bench scripts/dp_bench.cu && /tmp/dp_bench > gpurun_out/dp_bench.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t7.log
python bench.py --frames 2048 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/b10_2048.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches_v3d.csv python bench.py --frames 512 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_v3d.log 2>&1
