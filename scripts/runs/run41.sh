set -x
for b in 1776 2368; do
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --batch $b > gpurun_out/b41_$b.log 2>&1
done
