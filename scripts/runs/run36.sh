set -x
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
