set -x
python scripts/linkprof.py > gpurun_out/linkprof5.log 2>&1
