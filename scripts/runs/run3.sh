set -x
ncu --set full --clock-control none --import-source on -k regex:"gauss_decide|blur_prepass" -s 6 -c 2 -o gpurun_out/prof_v3 python bench.py --frames 512 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_v3full.log 2>&1
python scripts/linkprof.py > gpurun_out/lp2.log 2>&1
