set -x
for pad in 0 28000 40000; do
YSMR_K1B_PAD=$pad python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/b34_$pad.log 2>&1
done
YSMR_FLAT_PRIO=1 YSMR_K1B_PAD=40000 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/b34_flat.log 2>&1
