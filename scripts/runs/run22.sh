set -x
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -x -q > gpurun_out/t23.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t23.log
for v in 6_4 3_8 5_2; do
  YSMR_LIB=$PWD/variant_k1b_$v.so python bench.py --frames 2368 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/b25_$v.log 2>&1
done
python bench.py --frames 2368 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/b25_base.log 2>&1
