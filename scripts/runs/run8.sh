set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t9.log
python bench.py --frames 2048 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/b12_2048.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"link_kernel" -s 4 -c 1 -o gpurun_out/prof_link_v2 python bench.py --frames 1024 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_link_v2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches_v3e.csv python bench.py --frames 512 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_v3e.log 2>&1
