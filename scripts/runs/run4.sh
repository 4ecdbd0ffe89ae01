set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t5.log
python bench.py --frames 2048 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/b8_2048.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gauss_decide" -s 6 -c 1 -o gpurun_out/prof_v3b python bench.py --frames 512 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_v3b.log 2>&1
