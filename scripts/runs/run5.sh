set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t6.log
python bench.py --frames 2048 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/b9_2048.log 2>&1
python bench.py --frames 2048 --steps 2 --warmup 3 --no-cpu --no-e2e --channels 1 > gpurun_out/b9_2048_c1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"blur_prepass" -s 6 -c 1 -o gpurun_out/prof_v3c python bench.py --frames 512 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_v3c.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches_v3c.csv python bench.py --frames 512 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_v3c2.log 2>&1
