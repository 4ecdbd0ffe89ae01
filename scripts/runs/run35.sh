set -x
for b in 592 888 1184 1776; do
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --batch $b > gpurun_out/b35_$b.log 2>&1
done
