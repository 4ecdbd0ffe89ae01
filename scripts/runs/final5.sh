set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_final5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_final5.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
bash scripts/runs/profiles.sh
