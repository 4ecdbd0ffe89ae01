set -x
timeout 600 python -m pytest tests/test_gpu_link.py tests/test_gpu_e2e.py -m gpu -x -q > gpurun_out/t20.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t20.log
ncu --set full --clock-control none --import-source on -k regex:"link_kernel" -s 4 -c 1 -o gpurun_out/prof_link_v3 python bench.py --frames 2368 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_link_v3.log 2>&1
