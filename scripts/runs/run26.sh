set -x
timeout 900 python -m pytest tests/test_gpu_link.py tests/test_gpu_e2e.py -m gpu -x -q > gpurun_out/t27.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t27.log
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/b29.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"link_kernel" -s 4 -c 1 -o gpurun_out/prof_link_v4 python bench.py --frames 2368 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_link_v4.log 2>&1
