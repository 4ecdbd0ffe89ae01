#!/bin/bash
python scripts/link_alone.py cfg2 2368 2>&1 | tail -2
for mode in cs ldg; do
  if [ $mode = ldg ]; then export YSMR_FUSED_LOADS=ldg; else unset YSMR_FUSED_LOADS; fi
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_${mode}.json 2> gpurun_out/bench_${mode}.err
  python -c "
import json
d=json.loads(open('gpurun_out/bench_${mode}.json').read().strip().splitlines()[-1]); print('${mode}', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
done
unset YSMR_FUSED_LOADS
B="python bench.py --steps 2 --warmup 3 --frames 2368 --no-e2e --no-cpu"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fused_front|label_kernel|geometry|link_|rows_|blur_prepass|gauss_decide|pack_masks|plane_margins" -c 400 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_ncu_launches.log 2>&1
LINKPROF_FRAMES=64 LINKPROF_PLAIN=1 LINKPROF_CFG=cfg3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:link_general_grid -c 1 -f -o gpurun_out/r2_grid_cfg3 python scripts/linkprof.py > gpurun_out/ncu_grid.log 2>&1; tail -2 gpurun_out/ncu_grid.log
ls -la gpurun_out/r2_grid_cfg3.ncu-rep gpurun_out/r2_launches.csv
